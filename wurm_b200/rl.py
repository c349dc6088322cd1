"""Advantage actor-critic losses with the return scan on the device (reference wurm/rl/a2c.py:9-79).

`A2C` keeps the reference's constructor and `loss(bootstrap_values, rewards, values, log_probs, dones)`.
The backward recurrence over the trajectory -- a Python loop of ~6 tensor ops per time step in the
reference (:49-63) -- is one kernel launch (`wurm_a2c_returns`, wurm_b200/csrc/a2c.cu), bit-identical in fp32;
the two closing reductions (:70-73) stay torch ops so that autograd sees `values` and `log_probs`.

Gradients: with n-step returns the targets carry no gradient in the reference either (the bootstrap values
are computed under no_grad and rewards have none), so the losses and their gradients are identical.  With GAE
the reference lets gradient flow into `values` *through the returns*; here the returns are constants
(the usual treatment of targets): same loss values, different value-loss gradient.
"""
import ctypes
from typing import Callable

import torch
import torch.nn.functional as F

from . import _lib

EPS = 1e-8


def returns(bootstrap_values: torch.Tensor, rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, gamma: float,
            gae_lambda: float = None) -> torch.Tensor:
    """Returns of shape rewards.shape ((T, N) or (T, N, 1)), no gradient."""
    T, N = rewards.shape[0], rewards.shape[1]
    dev = rewards.device
    if dev.type != 'cuda':
        raise RuntimeError('wurm_b200.rl runs on CUDA tensors only (there is no CPU fallback; the CPU implementation is the reference)')
    r = rewards.detach().reshape(T, N).to(torch.float32).contiguous()
    v = values.detach().reshape(T, N).to(torch.float32).contiguous()
    d = dones.reshape(T, N).to(torch.bool).contiguous()
    b = bootstrap_values.detach().reshape(N).to(torch.float32).contiguous()
    out = torch.empty((T, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().wurm_a2c_returns(T, N, float(gamma), -1.0 if gae_lambda is None else float(gae_lambda),
                                               b.data_ptr(), r.data_ptr(), v.data_ptr(), d.data_ptr(), out.data_ptr(),
                                               ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out.reshape(rewards.shape)


class A2C(object):
    """Advantage actor-critic (reference wurm/rl/a2c.py:9-31: same arguments)."""

    def __init__(self,
                 gamma: float,
                 value_loss_fn: Callable = F.smooth_l1_loss,
                 normalise_returns: bool = False,
                 use_gae: bool = False,
                 gae_lambda: float = None,
                 dtype: torch.dtype = torch.float):
        self.gamma = gamma
        self.normalise_returns = normalise_returns
        self.use_gae = use_gae
        self.gae_lambda = gae_lambda
        self.value_loss_fn = value_loss_fn
        self.dtype = dtype

    def loss(self, bootstrap_values, rewards, values, log_probs, dones, return_returns: bool = False):
        R = returns(bootstrap_values, rewards, values, dones, self.gamma, self.gae_lambda if self.use_gae else None)
        if self.normalise_returns:
            R = (R - R.mean()) / (R.std() + EPS)
        value_loss = self.value_loss_fn(values, R).mean()
        advantages = R - values
        policy_loss = - (advantages.detach() * log_probs).mean()
        if return_returns:
            return value_loss, policy_loss, R
        return value_loss, policy_loss
