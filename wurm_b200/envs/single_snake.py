"""Batched classic-snake environment, B200-native.

Drop-in for the reference's `wurm.envs.SingleSnake` (wurm/envs/single_snake.py:17-428): same
constructor arguments, attributes (`envs`, `done`, `num_envs`, `size`, ...), `reset(done)`,
`step(actions) -> (observations, reward, done, info)`, exceptions and side effects (the caller's
`actions` tensor is sanitised in place, :222).  The state stays in the reference's own layout --
`envs` is a plain `(num_envs, 3, size, size)` fp32 CUDA tensor callers may read and write -- and each
call is ONE kernel launch through the C ABI of include/wurm_b200.h (wurm_b200/csrc/single_snake.cu).

Randomness (food respawn, spawn position/direction) is an input of the kernels: by default it is
derived on the device from Philox4x32-10 keyed by (`seed`, call counter, env); the keyword-only
`food_cell_replay` / `spawn_replay` arguments inject recorded draws instead, which is how bit-exact
parity with the reference is tested (the reference's own draws go through an unstable sort of a CPU
randperm, wurm/utils.py:188,224, and cannot be regenerated).
"""
from collections import namedtuple
from time import time
import ctypes
import os

import torch

from .. import _lib
from ..config import DEFAULT_DEVICE, BODY_CHANNEL, EPS, HEAD_CHANNEL, FOOD_CHANNEL  # noqa: F401

Spec = namedtuple('Spec', ['reward_threshold'])

_OBS_MODES = {'default': _lib.OBS_DEFAULT, 'raw': _lib.OBS_RAW, 'one_channel': _lib.OBS_ONE_CHANNEL,
              'positions': _lib.OBS_POSITIONS}
_ACTION_BYTES = {torch.uint8: 1, torch.short: 2, torch.int: 4, torch.long: 8}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class SingleSnake(object):
    """Batched snake environment (state layout and dynamics: reference single_snake.py:18-47).

    Channel 0 food (0/1), channel 1 head (0/1), channel 2 body (1 = tail ... L = head cell).
    """

    spec = Spec(float('inf'))
    metadata = {
        'render.modes': ['rgb_array'],
        'video.frames_per_second': 12
    }
    supports_fused_reset = True     # step(..., auto_reset=True): wurm_single_step_reset

    def __init__(self,
                 num_envs: int,
                 size: int,
                 max_timesteps: int = None,
                 initial_snake_length: int = 3,
                 on_death: str = 'restart',
                 observation_mode: str = 'one_channel',
                 device: str = DEFAULT_DEVICE,
                 manual_setup: bool = False,
                 verbose: int = 0,
                 render_args: dict = None,
                 seed: int = None,
                 state: str = None):
        self._lib = _lib.lib()      # raises if the CUDA library is not built: there is no fallback
        self._cfg_cache = {}
        if state is None:
            state = os.environ.get('WURM_B200_STATE', 'dense')     # process-wide default (how the reference's own tests are
                                                                   # run against the compact state without touching them)
        if state == 'dense_scan':                    # MultiSnake's name for "tensors only, no shadow records": SingleSnake's dense
            state = 'dense'                          # state never has any (tried and lost, profiles/r02_single_shadow_experiment.txt)
        if state not in ('dense', 'compact'):
            raise ValueError("state must be 'dense' (the reference's fp32 tensor is the state) or 'compact'")
        # state='compact' (an extension): between calls the env lives in HBM as one uint16 record per cell (include/
        # wurm_b200.h, wurm_single_compact_step) instead of the reference's (3,S,S) fp32 grids; `envs` is then materialised
        # on attribute access and folded back in if the caller wrote to it.
        self._compact = state == 'compact'
        self.num_envs = num_envs
        self.size = size
        self.max_timesteps = max_timesteps
        self.initial_snake_length = initial_snake_length
        self.on_death = on_death
        self.observation_mode = observation_mode
        self.device = device
        self.verbose = verbose
        if torch.device(device).type != 'cuda':
            raise RuntimeError(f"wurm_b200 envs run on CUDA devices only (got device={device!r}); "
                               "the CPU implementation of this path is the reference itself")

        if render_args is None:
            self.render_args = {'num_rows': 1, 'num_cols': 1, 'size': 256}
        else:
            self.render_args = render_args

        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, ()).item())    # follows torch.manual_seed
        self.seed = seed
        self._draws = 0          # Philox call counter: one tick per call that may draw
        # device-side addend of the call counter: stays 0 in normal use, bumped between CUDA-graph replays
        self._draws_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        # episode statistics accumulated by the step kernel (see `stats`)
        self._stats = torch.zeros((_lib.STATS_SLOTS, _lib.STATS_FIELDS), dtype=torch.int64, device=self.device)
        # (head cell, snake size, food cell, -) per env left by one kernel call for the next: verified hints that let the step
        # kernel skip two of its scans; never trusted, so editing `envs` behind the env's back stays safe
        self._hints = torch.full((num_envs, 4), -1, dtype=torch.short, device=self.device)

        self.t = 0
        self._hint_key = None
        self._dev = self._hints.device
        if self._compact:
            self._cells = torch.zeros((num_envs, (size * size + 7) & ~7), dtype=torch.int16, device=self.device)
            self._hints.zero_()
            self._hints[:, 0] = -1                   # aux vectors: no head, size 0, no neck, not canonical
            self._hints[:, 2] = -1
            self._dense = None                       # the fp32 tensor exists only while a caller looks at it
            self._dense_key = None
            if not manual_setup:
                if self.size <= 8:
                    raise NotImplementedError('Cannot make an env this small without making this code more clever')
                if self.initial_snake_length != 3:
                    raise NotImplementedError('Only initial snake length = 3 has been implemented.')
                self._reset_mask(None, torch.ones(num_envs, dtype=torch.bool, device=self.device))
        else:
            self._dense = torch.zeros((num_envs, 3, size, size), device=self.device)
            if not manual_setup:
                self._dense = self._create_envs(self.num_envs)
            self._adopt_state()

        self.done = torch.zeros(num_envs, dtype=torch.bool, device=self.device)

        self.viewer = None

        self.body_colour = torch.tensor((0, 127, 0), dtype=torch.short, device=self.device)
        self.head_colour = torch.tensor((0, 255, 0), dtype=torch.short, device=self.device)
        self.food_colour = torch.tensor((255, 0, 0), dtype=torch.short, device=self.device)
        self.edge_colour = torch.tensor((0, 0, 0), dtype=torch.short, device=self.device)

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def _cfg(self, observation_mode, num_envs=None):
        key = (observation_mode, self.observation_mode, num_envs, self.num_envs, self.size)
        cached = self._cfg_cache.get(key)
        if cached is None:
            cached = self._cfg_cache[key] = self._make_cfg(observation_mode, num_envs)
        return cached

    def _make_cfg(self, observation_mode, num_envs=None):
        if observation_mode is None:
            mode, n = _lib.OBS_NONE, 0
        elif observation_mode.startswith('partial_'):
            # the reference parses the window from self.observation_mode, not the argument (:167)
            mode, n = _lib.OBS_PARTIAL, int(self.observation_mode.split('_')[-1])
        elif observation_mode in _OBS_MODES:
            mode, n = _OBS_MODES[observation_mode], 0
        else:
            raise Exception(f'Unrecognised observation mode {observation_mode!r}')   # :195
        return _lib.WurmSingleCfg(self.num_envs if num_envs is None else num_envs, self.size, mode, n)

    def _obs_shape(self, cfg):
        n, s = cfg.num_envs, self.size
        w = 2 * cfg.obs_n + 1
        return {_lib.OBS_DEFAULT: (n, 3, s, s), _lib.OBS_RAW: (n, 3, s, s), _lib.OBS_ONE_CHANNEL: (n, 1, s, s),
                _lib.OBS_POSITIONS: (n, 4), _lib.OBS_PARTIAL: (n, 3 * w * w)}[cfg.obs_mode]

    # ---- `envs`: a plain tensor in dense mode, materialised on access in compact mode ----
    @property
    def envs(self):
        if self._dense is None:
            self._materialise()
        return self._dense

    @envs.setter
    def envs(self, value):
        self._dense = value
        self._dense_key = None               # compact mode: an assigned tensor is always folded back in before the next call

    def _materialise(self):
        """compact records -> the reference's (N,3,S,S) fp32 tensor (one launch), kept until the next state-changing call."""
        dense = torch.empty((self.num_envs, 3, self.size, self.size), dtype=torch.float32, device=self._dev)
        cfg = _lib.WurmSingleCfg(self.num_envs, self.size, _lib.OBS_NONE, 0)
        with _lib.device_guard(self._dev):
            _lib.check(self._lib.wurm_single_expand(ctypes.byref(cfg), _ptr(self._cells), _ptr(self._hints), _ptr(dense),
                                                    self._stream()))
        self._dense = dense
        self._dense_key = (dense.data_ptr(), dense._version)

    def _compact_state(self, mutates):
        """Compact mode: folds a caller-edited materialised tensor back into the records (raises if they cannot carry it)
        and drops the materialised tensor when the call about to be made changes the state."""
        if self._dense is not None:
            e = self._dense
            if e.dtype != torch.float32 or not e.is_contiguous() or e.device != self._dev:
                e = e.to(device=self._dev, dtype=torch.float32).contiguous()
                self._dense = e
            if tuple(e.shape) != (self.num_envs, 3, self.size, self.size):
                raise RuntimeError(f'envs has shape {tuple(e.shape)}, expected {(self.num_envs, 3, self.size, self.size)}')
            if (e.data_ptr(), e._version) != self._dense_key:
                cfg = _lib.WurmSingleCfg(self.num_envs, self.size, _lib.OBS_NONE, 0)
                with _lib.device_guard(self._dev):
                    _lib.check(self._lib.wurm_single_compact(ctypes.byref(cfg), _ptr(e), _ptr(self._cells), _ptr(self._hints),
                                                             _ptr(self._status), self._stream()))
                self._dense_key = (e.data_ptr(), e._version)
                st = int(self._status.item())
                if st & _lib.ST_NOT_COMPACT:
                    self._status.zero_()
                    raise RuntimeError("state='compact' cannot carry this state exactly (food / head values other than 0 or 1, "
                                       "non-integral, negative or oversized body values); use state='dense'")
            if mutates:
                self._dense = None

    def _before_replay(self):
        """GraphedStepper hook: what step() does before launching -- compact mode: fold caller edits of the materialised
        tensor in, then drop it; dense mode: drop the hints (on the device) if the caller wrote to `envs`."""
        if self._compact:
            self._compact_state(mutates=True)
        else:
            self._state()

    def _snapshot_names(self):
        """Attributes that make up the env's state (GraphedStepper snapshots them around its warm-up)."""
        return ('_cells' if self._compact else 'envs', 'done', '_hints', '_stats', '_status')

    def _state(self):
        """`envs` may have been replaced or sliced by the caller (tests assign it): normalise.  Also where the hints
        are dropped if the caller touched the tensor since this env's last own call: the kernels verify that the
        hinted head / food cells still hold a head / food, but not that they are the ONLY such cells, so a state
        edited behind the env's back (a second food cell, say -- which the reference's step handles like any other)
        must not be stepped on stale hints.  torch bumps `_version` on every in-place write through the tensor or
        any view of it, and an assignment changes `data_ptr()`; two integer compares per call."""
        e = self.envs
        if e.dtype != torch.float32 or not e.is_contiguous() or e.device.type != 'cuda':
            e = e.to(device=self.device, dtype=torch.float32).contiguous()
            self.envs = e
        if tuple(e.shape) != (self.num_envs, 3, self.size, self.size):
            raise RuntimeError(f'envs has shape {tuple(e.shape)}, expected {(self.num_envs, 3, self.size, self.size)}')
        if (e.data_ptr(), e._version) != self._hint_key:
            self.invalidate_hints()
        return e

    def _adopt_state(self):
        """Records the identity of `envs` as the state the hints describe (after this env's own kernels wrote it)."""
        if self._compact:
            if self._dense is not None:
                self._dense_key = (self._dense.data_ptr(), self._dense._version)
            return
        self._hint_key = (self.envs.data_ptr(), self.envs._version)

    def invalidate_hints(self):
        """Drops the per-env (head cell, size, food cell) hints: the next step re-derives everything from `envs`.
        Called automatically when `envs` was replaced or written through torch; call it by hand after writing the
        state through a raw pointer (a custom kernel, `.data_ptr()`), which torch's version counter cannot see."""
        if self._compact:                            # the aux vectors are derived state in this mode, not hints: what "forget
            self._dense_key = None                   # what you assumed" means here is "fold the materialised tensor back in"
            return
        self._hints.fill_(-1)
        self._adopt_state()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)

    def stats(self, reduce_group=None):
        """Episode statistics accumulated on the device since construction: a dict of int counters
        (env_steps, episodes, reward, self_collisions, edge_collisions).  With `reduce_group` (a
        torch.distributed process group, or True for the default group) the counters are summed
        over ranks with one small all-reduce -- the only collective this path has."""
        totals = self._stats.sum(dim=0)
        if reduce_group is not None:
            from ..distributed import all_reduce_stats
            all_reduce_stats(totals, None if reduce_group is True else reduce_group)
        return dict(zip(_lib.STAT_NAMES, totals.tolist()))

    def check_consistency(self, skip: torch.Tensor = None):
        """env_consistency (reference wurm/utils.py:167-178) of every env whose `skip` flag is not set, in one
        fused kernel -- what the reference driver does with `env_consistency(env.envs[~done])` (main.py:215),
        without the boolean-mask gather.  Raises the reference's RuntimeError messages."""
        if self._compact:                    # on the records: no materialisation of the fp32 tensor
            self._compact_state(mutates=False)
            report = torch.tensor([0, 0, 2 ** 31 - 1, 0], dtype=torch.int32, device=self._dev)
            cfg = _lib.WurmSingleCfg(self.num_envs, self.size, _lib.OBS_NONE, 0)
            if skip is not None:
                skip = skip.reshape(-1)
                if skip.dtype != torch.bool or skip.device != self._dev or not skip.is_contiguous():
                    skip = (skip != 0).to(self._dev).contiguous()
            with _lib.device_guard(self._dev):
                _lib.check(self._lib.wurm_single_compact_check(ctypes.byref(cfg), _ptr(self._cells), _ptr(skip), _ptr(report),
                                                               self._stream()))
            _lib.raise_on_report(report.tolist())
            self.check_status()
            return
        from ..utils import _fused_check
        _lib.raise_on_report(_fused_check(self._state(), skip))
        self.check_status()

    def check_status(self):
        """Raises if a kernel met a state outside the supported set since the last check (one sync)."""
        st = int(self._status.item())
        if st:
            self._status.zero_()
            msgs = []
            if st & _lib.ST_MULTI_HEAD:
                msgs.append('an environment holds more than one head cell')
            if st & _lib.ST_NO_HEAD_PARTIAL:
                msgs.append("partial observation of an environment without a head (the reference raises a view-shape "
                            "error here); zeros were written")
            raise RuntimeError('; '.join(msgs) or f'status {st}')

    # ------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------
    def _observe(self, observation_mode: str = 'default'):
        cfg = self._cfg(observation_mode)
        if self._compact:
            self._compact_state(mutates=False)
            obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=self._dev)
            with _lib.device_guard(self._dev):
                _lib.check(self._lib.wurm_single_compact_observe(ctypes.byref(cfg), _ptr(self._cells), _ptr(self._hints), _ptr(obs),
                                                                 _ptr(self._status), self._stream()))
            return obs
        envs = self._state()
        obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=envs.device)
        with _lib.device_guard(envs.device):
            _lib.check(self._lib.wurm_single_observe(ctypes.byref(cfg), _ptr(envs), _ptr(obs), _ptr(self._status),
                                                     self._stream()))
        return obs

    def _get_rgb(self):
        """int16 (N,3,S,S) image, as displayed by render() (reference :104-128)."""
        return (self._observe('default') * 255).round().short()

    def step(self, actions: torch.Tensor, *, food_cell_replay: torch.Tensor = None, auto_reset: bool = False,
             spawn_replay: torch.Tensor = None, obs_out: torch.Tensor = None, packed_out: torch.Tensor = None):
        """reference :197-304.  `auto_reset=True` (an extension) fuses the `reset(done)` the reference's driver
        issues right after every step (experiments/main.py:227) into the same launch: the returned observation,
        reward, done and info are the step's (terminal observation for envs that just ended, as in the
        reference's loop), `envs` holds the re-created environments; bit-identical to the two calls.
        `obs_out` (an extension): a contiguous fp32 CUDA tensor of the observation's shape -- e.g. one time slice of
        a trajectory buffer or the policy's input buffer -- that the kernel renders into directly.
        `packed_out` (an extension): a (num_envs,) uint8 CUDA tensor that additionally receives the step's results as
        one byte per env (bit 0 done, bit 1 self collision, bit 2 edge collision, bits 3-4 reward): what `HostStepper`
        brings back over PCIe.  uint8 actions are accepted besides the reference's three integer types."""
        if actions.dtype not in _ACTION_BYTES:
            raise TypeError('actions Tensor must be an integer type i.e. '
                            '{torch.ShortTensor, torch.IntTensor, torch.LongTensor}')

        if actions.shape[0] != self.num_envs:
            raise RuntimeError('Must have the same number of actions as environments.')

        t0 = time()
        if self._compact:
            self._compact_state(mutates=True)
            envs, dev = None, self._dev
        else:
            envs = self._state()
            dev = envs.device
        host_actions = None
        if actions.device != dev or not actions.is_contiguous():
            host_actions, actions = actions, actions.to(dev, non_blocking=True).contiguous()
        cfg = self._cfg(self.observation_mode)
        if obs_out is None:
            obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=dev)
        else:
            obs = obs_out
            if (tuple(obs.shape) != self._obs_shape(cfg) or obs.dtype != torch.float32 or obs.device != dev
                    or not obs.is_contiguous()):
                raise RuntimeError(f'obs_out must be a contiguous float32 tensor of shape {self._obs_shape(cfg)} on {dev}')
        reward = torch.empty(self.num_envs, dtype=torch.float32, device=dev)
        flags = torch.empty((3, self.num_envs), dtype=torch.bool, device=dev)        # one allocation for the three flag vectors
        done, self_collision, edge_collision = flags[0], flags[1], flags[2]
        if food_cell_replay is not None:
            food_cell_replay = food_cell_replay.to(device=dev, dtype=torch.int32).contiguous()
        if spawn_replay is not None:
            spawn_replay = spawn_replay.to(device=dev, dtype=torch.int32).contiguous()
        if packed_out is not None and (packed_out.dtype != torch.uint8 or packed_out.device != dev or
                                       packed_out.numel() != self.num_envs or not packed_out.is_contiguous()):
            raise RuntimeError(f'packed_out must be a contiguous uint8 tensor of {self.num_envs} elements on {dev}')
        self._draws += 1
        with _lib.device_guard(dev):
            if self._compact:
                _lib.check(self._lib.wurm_single_compact_step(
                    ctypes.byref(cfg), _ptr(self._cells), _ptr(self._hints), _ptr(actions), _ACTION_BYTES[actions.dtype],
                    _ptr(food_cell_replay), int(bool(auto_reset)), _ptr(spawn_replay), self.seed, self._draws,
                    _ptr(self._draws_dev), _ptr(obs), _ptr(reward), _ptr(done), _ptr(self_collision), _ptr(edge_collision),
                    _ptr(self._status), _ptr(self._stats), _ptr(packed_out), self._stream()))
                if auto_reset:
                    self._draws += 1        # the fused reset consumed the next counter value
            elif auto_reset:
                _lib.check(self._lib.wurm_single_step_reset(
                    ctypes.byref(cfg), _ptr(envs), _ptr(actions), _ACTION_BYTES[actions.dtype], _ptr(food_cell_replay),
                    _ptr(spawn_replay), self.seed, self._draws, _ptr(self._draws_dev), _ptr(obs), _ptr(reward), _ptr(done),
                    _ptr(self_collision), _ptr(edge_collision), _ptr(self._status), _ptr(self._stats), _ptr(self._hints),
                    _ptr(packed_out), self._stream()))
                self._draws += 1            # the fused reset consumed the next counter value
            else:
                _lib.check(self._lib.wurm_single_step(
                    ctypes.byref(cfg), _ptr(envs), _ptr(actions), _ACTION_BYTES[actions.dtype], _ptr(food_cell_replay),
                    self.seed, self._draws, _ptr(self._draws_dev), _ptr(obs), _ptr(reward), _ptr(done),
                    _ptr(self_collision), _ptr(edge_collision), _ptr(self._status), _ptr(self._stats), _ptr(self._hints),
                    _ptr(packed_out), self._stream()))
        if host_actions is not None:
            host_actions.copy_(actions, non_blocking=True)      # the sanitised actions (reference :222)
        info = {'self_collision': self_collision, 'edge_collision': edge_collision}
        self.done = done
        if self.verbose > 0:
            torch.cuda.synchronize(dev)
            print(f'step: {time() - t0}s')
        return obs, reward.unsqueeze(-1), done.unsqueeze(-1), info

    def reset(self, done: torch.Tensor = None, *, spawn_replay: torch.Tensor = None,
              return_observations: bool = True):
        """Re-creates the environments flagged in `done` (reference :322-342).

        `return_observations=False` (an extension, as MultiSnake.reset has in the reference) skips the
        full observation the reference recomputes and its drivers discard (experiments/main.py:227).
        """
        if done is None:
            done = self.done

        done = done.view((done.shape[0]))
        if done.shape[0] != self.num_envs:
            raise RuntimeError('Must have one done flag per environment.')

        t0 = time()
        if self._compact:
            self._compact_state(mutates=True)
            envs = None
        else:
            envs = self._state()
        if done.dtype == torch.bool and done.device == self._dev and done.is_contiguous():
            mask = done                     # the step's own flags: no conversion kernel on the hot loop
        else:
            mask = (done != 0).to(device=self._dev).contiguous()
        self._reset_mask(envs, mask, spawn_replay)

        if self.verbose:
            print(f'Resetting {done.sum().item()} envs: {time() - t0}s')

        if return_observations:
            return self._observe(self.observation_mode)

    def _reset_mask(self, envs, mask, spawn_replay=None):
        if spawn_replay is not None:
            spawn_replay = spawn_replay.to(device=self._dev, dtype=torch.int32).contiguous()
        self._draws += 1
        if envs is None:                    # compact resident state
            cfg = _lib.WurmSingleCfg(self.num_envs, self.size, _lib.OBS_NONE, 0)
            with _lib.device_guard(self._dev):
                _lib.check(self._lib.wurm_single_compact_reset(ctypes.byref(cfg), _ptr(self._cells), _ptr(self._hints), _ptr(mask),
                                                               _ptr(spawn_replay), self.seed, self._draws, _ptr(self._draws_dev),
                                                               self._stream()))
            return
        cfg = _lib.WurmSingleCfg(envs.shape[0], self.size, _lib.OBS_NONE, 0)
        with _lib.device_guard(envs.device):
            _lib.check(self._lib.wurm_single_reset(ctypes.byref(cfg), _ptr(envs), _ptr(mask), _ptr(spawn_replay),
                                                   self.seed, self._draws, _ptr(self._draws_dev),
                                                   _ptr(self._hints) if (envs.shape[0] == self.num_envs and not self._compact) else None,
                                                   self._stream()))

    def _create_envs(self, num_envs: int, *, spawn_replay: torch.Tensor = None):
        """Vectorised environment creation (reference :344-387)."""
        if self.size <= 8:
            raise NotImplementedError('Cannot make an env this small without making this code more clever')

        if self.initial_snake_length != 3:
            raise NotImplementedError('Only initial snake length = 3 has been implemented.')

        envs = torch.empty((num_envs, 3, self.size, self.size), device=self.device)
        self._reset_mask(envs, torch.ones(num_envs, dtype=torch.bool, device=self.device), spawn_replay)
        return envs

    def render(self, mode: str = 'human'):
        """Human display of `_get_rgb()` (reference :389-428); host-side convenience, outside the hot path."""
        from ._display import show
        return show(self, self._get_rgb(), mode)
