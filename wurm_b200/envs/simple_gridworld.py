"""Batched gridworld environment, B200-native.

Drop-in for the reference's `wurm.envs.SimpleGridworld` (wurm/envs/simple_gridworld.py:15-265): an agent
moves in the 4 cardinal directions, gets +1 for stepping on the food (which respawns on a free interior
cell) and dies when it walks onto the border.  `envs` is the reference's `(num_envs, 2, size, size)` fp32
tensor (channel 0 food, channel 1 agent); every call is one kernel launch through the C ABI
(wurm_b200/csrc/gridworld.cu).  Selectable from the reference's driver with `--env gridworld`
(experiments/main.py:166-168).
"""
from collections import namedtuple
from typing import Tuple
import ctypes

import torch

from .. import _lib
from ..config import DEFAULT_DEVICE

Spec = namedtuple('Spec', ['reward_threshold'])

_OBS_MODES = {'default': _lib.OBS_DEFAULT, 'raw': _lib.OBS_RAW, 'positions': _lib.OBS_POSITIONS}
_ACTION_BYTES = {torch.uint8: 1, torch.short: 2, torch.int: 4, torch.long: 8}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class SimpleGridworld(object):
    spec = Spec(float('inf'))

    def __init__(self,
                 num_envs: int,
                 size: int,
                 on_death: str = 'restart',
                 observation_mode: str = 'default',
                 device: str = DEFAULT_DEVICE,
                 start_location: Tuple[int, int] = None,
                 manual_setup: bool = False,
                 verbose: int = 0,
                 seed: int = None):
        self._lib = _lib.lib()
        self.num_envs = num_envs
        self.size = size
        self.on_death = on_death
        self.observation_mode = observation_mode
        self.start_location = start_location
        self.device = device
        self.verbose = verbose
        if torch.device(device).type != 'cuda':
            raise RuntimeError(f"wurm_b200 envs run on CUDA devices only (got device={device!r})")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, ()).item())
        self.seed = seed
        self._draws = 0
        self._draws_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._stats = torch.zeros((_lib.STATS_SLOTS, _lib.STATS_FIELDS), dtype=torch.int64, device=self.device)
        self.t = 0

        self.envs = torch.zeros((num_envs, 2, size, size), device=self.device)
        if not manual_setup:
            self.envs = self._create_envs(self.num_envs)

        self.done = torch.zeros(num_envs, dtype=torch.bool, device=self.device)
        self._dev = self.done.device
        self.viewer = None
        self.head_colour = torch.tensor((0, 255, 0), dtype=torch.short, device=self.device)
        self.food_colour = torch.tensor((255, 0, 0), dtype=torch.short, device=self.device)
        self.edge_colour = torch.tensor((0, 0, 0), dtype=torch.short, device=self.device)

    def _cfg(self, observation_mode, num_envs=None):
        if observation_mode is None:
            mode = _lib.OBS_NONE
        elif observation_mode in _OBS_MODES:
            mode = _OBS_MODES[observation_mode]
        else:
            raise Exception(f'Unrecognised observation mode {observation_mode!r}')       # reference :132-133
        sy, sx = self.start_location if self.start_location is not None else (0, 0)
        return _lib.WurmGridCfg(self.num_envs if num_envs is None else num_envs, self.size, mode, int(sy), int(sx))

    def _obs_shape(self, cfg):
        n, s = cfg.num_envs, self.size
        return {_lib.OBS_DEFAULT: (n, 3, s, s), _lib.OBS_RAW: (n, 2, s, s), _lib.OBS_POSITIONS: (n, 4)}[cfg.obs_mode]

    def _state(self):
        e = self.envs
        if e.dtype != torch.float32 or not e.is_contiguous() or e.device.type != 'cuda':
            e = e.to(device=self.device, dtype=torch.float32).contiguous()
            self.envs = e
        if tuple(e.shape) != (self.num_envs, 2, self.size, self.size):
            raise RuntimeError(f'envs has shape {tuple(e.shape)}, expected {(self.num_envs, 2, self.size, self.size)}')
        return e

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.envs.device).cuda_stream)

    def stats(self, reduce_group=None):
        """Episode statistics accumulated on the device (env_steps, episodes, reward, edge_collisions)."""
        totals = self._stats.sum(dim=0)
        if reduce_group is not None:
            from ..distributed import all_reduce_stats
            all_reduce_stats(totals, None if reduce_group is True else reduce_group)
        return dict(zip(_lib.STAT_NAMES, totals.tolist()))

    def check_status(self):
        """Raises if a kernel met a state outside the supported set since the last check (one sync)."""
        st = int(self._status.item())
        if st:
            self._status.zero_()
            raise RuntimeError('an environment holds more than one agent cell' if st & _lib.ST_MULTI_HEAD else f'status {st}')

    def _observe(self, observation_mode: str = 'default'):
        cfg = self._cfg(observation_mode)
        envs = self._state()
        obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=envs.device)
        with _lib.device_guard(envs.device):
            _lib.check(self._lib.wurm_grid_observe(ctypes.byref(cfg), _ptr(envs), _ptr(obs), self._stream()))
        return obs

    def _get_rgb(self):
        return (self._observe('default') * 255).round().short()

    def step(self, actions: torch.Tensor, *, food_cell_replay: torch.Tensor = None):
        if actions.dtype not in _ACTION_BYTES:     # the reference's three integer types, plus uint8 (an extension)
            raise TypeError('actions Tensor must be an integer type i.e. '
                            '{torch.ShortTensor, torch.IntTensor, torch.LongTensor}')

        if actions.shape[0] != self.num_envs:
            raise RuntimeError('Must have the same number of actions as environments.')

        envs = self._state()
        dev = envs.device
        if actions.device != dev or not actions.is_contiguous():
            actions = actions.to(dev, non_blocking=True).contiguous()
        cfg = self._cfg(self.observation_mode)
        obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=dev)
        reward = torch.empty(self.num_envs, dtype=torch.float32, device=dev)
        done = torch.empty(self.num_envs, dtype=torch.bool, device=dev)
        if food_cell_replay is not None:
            food_cell_replay = food_cell_replay.to(device=dev, dtype=torch.int32).contiguous()
        self._draws += 1
        with _lib.device_guard(dev):
            _lib.check(self._lib.wurm_grid_step(
                ctypes.byref(cfg), _ptr(envs), _ptr(actions), _ACTION_BYTES[actions.dtype], _ptr(food_cell_replay),
                self.seed, self._draws, _ptr(self._draws_dev), _ptr(obs), _ptr(reward), _ptr(done), _ptr(self._status),
                _ptr(self._stats), self._stream()))
        self.done = done
        return obs, reward.unsqueeze(-1), done.unsqueeze(-1), {'edge_collision': done.clone()}

    def _reset_mask(self, envs, mask, food_cell_replay=None):
        if self.start_location is None:
            raise NotImplementedError("Haven't implemented random starting locations")      # reference :249
        cfg = self._cfg(None, envs.shape[0])
        if food_cell_replay is not None:
            food_cell_replay = food_cell_replay.to(device=envs.device, dtype=torch.int32).contiguous()
        self._draws += 1
        with _lib.device_guard(envs.device):
            _lib.check(self._lib.wurm_grid_reset(ctypes.byref(cfg), _ptr(envs), _ptr(mask), _ptr(food_cell_replay), self.seed,
                                                 self._draws, _ptr(self._draws_dev), self._stream()))

    def reset(self, done: torch.Tensor = None, *, food_cell_replay: torch.Tensor = None, return_observations: bool = True):
        if done is None:
            done = self.done
        done = done.view((done.shape[0]))
        envs = self._state()
        if done.dtype == torch.bool and done.device == envs.device and done.is_contiguous():
            mask = done
        else:
            mask = (done != 0).to(device=envs.device).contiguous()
        self._reset_mask(envs, mask, food_cell_replay)
        if return_observations:
            return self._observe(self.observation_mode)

    def _create_envs(self, num_envs: int, *, food_cell_replay: torch.Tensor = None):
        if self.size <= 4:
            raise NotImplementedError('Environemnts smaller than this don\'t make sense.')
        envs = torch.empty((num_envs, 2, self.size, self.size), device=self.device)
        self._reset_mask(envs, torch.ones(num_envs, dtype=torch.bool, device=self.device), food_cell_replay)
        return envs
