"""GPU: the A2C return scan (wurm_a2c_returns) is bit-identical to the oracle (pinned against the reference's
A2C.loss in oracle/validate_vs_reference.py) and the A2C class gives the losses of the reference's formulas."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from golden_util import assert_same

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('T,N', [(20, 512), (5, 33), (64, 100000), (1, 7)])
@pytest.mark.parametrize('gae_lambda', [None, 0.95])
def test_returns_match_oracle(T, N, gae_lambda):
    from wurm_b200 import rl
    g = torch.Generator().manual_seed(T * 1000 + N)
    rewards = (torch.rand(T, N, 1, generator=g) < 0.1).float() - (torch.rand(T, N, 1, generator=g) < 0.05).float()
    values = torch.randn(T, N, 1, generator=g)
    dones = torch.rand(T, N, 1, generator=g) < 0.08
    bootstrap = torch.randn(N, 1, generator=g)
    got = rl.returns(bootstrap.to(DEV), rewards.to(DEV), values.to(DEV), dones.to(DEV), 0.99, gae_lambda)
    expect = orc.a2c_returns(bootstrap.numpy().reshape(N), rewards.numpy().reshape(T, N), values.numpy().reshape(T, N),
                             dones.numpy().reshape(T, N), 0.99, gae_lambda)
    assert got.shape == rewards.shape
    assert_same(got.cpu().numpy().reshape(T, N), expect, 'returns')


def test_a2c_loss_and_gradients_n_step():
    """The reference's loss (n-step returns, wurm/rl/a2c.py:58-73) written with torch ops, against A2C.loss."""
    from wurm_b200.rl import A2C
    T, N, gamma = 20, 512, 0.99
    g = torch.Generator().manual_seed(1)
    rewards = (torch.rand(T, N, 1, generator=g) < 0.1).float().to(DEV)
    dones = (torch.rand(T, N, 1, generator=g) < 0.08).to(DEV)
    bootstrap = torch.randn(N, 1, generator=g).to(DEV)
    values_a = torch.randn(T, N, 1, generator=g).to(DEV).requires_grad_(True)
    logp_a = (-torch.rand(T, N, 1, generator=g)).to(DEV).requires_grad_(True)
    values_b, logp_b = values_a.detach().clone().requires_grad_(True), logp_a.detach().clone().requires_grad_(True)

    value_loss, policy_loss = A2C(gamma=gamma).loss(bootstrap, rewards, values_a, logp_a, dones)
    (value_loss + policy_loss).backward()

    returns = []
    R = bootstrap * (~dones[-1]).float()
    for r, d in zip(reversed(rewards), reversed(dones)):
        R = r + gamma * R * (~d).float()
        returns.insert(0, R)
    returns = torch.stack(returns)
    vl = torch.nn.functional.smooth_l1_loss(values_b, returns).mean()
    pl = -((returns - values_b).detach() * logp_b).mean()
    (vl + pl).backward()
    assert_same(value_loss.detach().cpu().numpy(), vl.detach().cpu().numpy(), 'value loss')
    assert_same(policy_loss.detach().cpu().numpy(), pl.detach().cpu().numpy(), 'policy loss')
    assert_same(values_a.grad.cpu().numpy(), values_b.grad.cpu().numpy(), 'd loss / d values')
    assert_same(logp_a.grad.cpu().numpy(), logp_b.grad.cpu().numpy(), 'd loss / d log_probs')


def test_trajectory_ring_receives_observations_straight_from_the_step_kernel():
    """TrajectoryStore(capacity=T): the no-gradient fields live in preallocated (T, N, ...) rings; with state_slot() the
    step kernel renders the observation into the ring itself (obs_out=), and the properties are views, not stacks."""
    from wurm_b200.envs import SingleSnake
    from wurm_b200.trajectory_store import TrajectoryStore
    N, S, T = 300, 9, 6
    ringed = SingleSnake(num_envs=N, size=S, observation_mode='partial_2', device='cuda', seed=9)
    listed = SingleSnake(num_envs=N, size=S, observation_mode='partial_2', device='cuda', seed=9)
    ring, plain = TrajectoryStore(capacity=T), TrajectoryStore()
    g = torch.Generator().manual_seed(2)
    for update in range(2):
        for t in range(T):
            a = torch.randint(0, 4, (N,), generator=g).to('cuda')
            slot = ring.state_slot((N, 75))
            obs, reward, done, _ = ringed.step(a.clone(), auto_reset=True, obs_out=slot)
            assert obs.data_ptr() == slot.data_ptr()
            ring.append(state=obs, action=a, reward=reward, done=done)
            obs2, reward2, done2, _ = listed.step(a.clone(), auto_reset=True)
            plain.append(state=obs2, action=a, reward=reward2, done=done2)
        assert len(ring) == T and ring.states.data_ptr() == ring._ring['state'].data_ptr()      # a view of the ring
        for name in ('states', 'actions', 'rewards', 'dones'):
            assert torch.equal(getattr(ring, name), getattr(plain, name)), name
        ring.clear(); plain.clear()
    with pytest.raises(RuntimeError):
        for t in range(T + 1):
            ring.append(reward=reward)
