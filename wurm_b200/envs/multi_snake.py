"""Batched slither-style multi-snake environment, B200-native.

Drop-in for the reference's `wurm.envs.MultiSnake` (wurm/envs/multi_snake.py:18-1019): same
constructor arguments and defaults, the same state attributes in the same layouts (`foods`, `heads`,
`bodies`, `dones`, `orientations`, `boost_this_step`, `agent_colours`, all readable and writable),
the same mutable rule attributes (`food_rate`, `food_on_death_prob`, `boost`, `boost_cost_prob`,
`food_mode`, `respawn_mode`, `reward_on_death` -- re-read at every call because the reference's
drivers anneal them, experiments/multiagent.py:337-345), and the same `step(dict) -> (obs dict,
rewards dict, dones dict incl. '__all__', info dict)`, `reset(done, return_observations)`,
`check_consistency()`, `_observe()`.  Every call is ONE kernel launch through the C ABI of
include/wurm_b200.h (wurm_b200/csrc/multi_snake.cu).

Randomness is an input of the kernels: by default derived on the device from Philox4x32-10 keyed by
(`seed`, call counter, env); the keyword-only `draws` arguments inject recorded draws instead, which is
how bit-exact parity with the reference is tested.
"""
from collections import namedtuple, OrderedDict
from time import time
from typing import Dict, Optional
import ctypes
import os

import numpy as np
import torch

from .. import _lib
from ..config import DEFAULT_DEVICE, EPS  # noqa: F401

Spec = namedtuple('Spec', ['reward_threshold'])

_ACTION_BYTES = {torch.uint8: 1, torch.short: 2, torch.int: 4, torch.long: 8}


def _ptr(t):
    return None if t is None else t.data_ptr()


class MultiSnake(object):
    """Batched multi-snake environment (state layout and dynamics: reference multi_snake.py:18-48)."""

    spec = Spec(float('inf'))
    metadata = {
        'render.modes': ['rgb_array'],
        'video.frames_per_second': 12
    }
    supports_fused_reset = True     # step(..., auto_reset=True): wurm_multi_step_reset

    def __init__(self,
                 num_envs: int,
                 num_snakes: int,
                 size: int,
                 initial_snake_length: int = 3,
                 on_death: str = 'restart',
                 observation_mode: str = 'full',
                 device: str = DEFAULT_DEVICE,
                 dtype: torch.dtype = torch.float,
                 manual_setup: bool = False,
                 food_on_death_prob: float = 0.5,
                 boost: bool = True,
                 boost_cost_prob: float = 0.5,
                 food_mode: str = 'only_one',
                 food_rate: float = 5e-4,
                 respawn_mode: str = 'all',
                 reward_on_death: int = -1,
                 verbose: int = 0,
                 render_args: dict = None,
                 agent_colours: str = 'random',
                 seed: int = None,
                 state: str = None):
        self._lib = _lib.lib()      # raises if the CUDA library is not built: there is no fallback
        self._cfg_cache = {}
        self.num_envs = num_envs
        self.num_snakes = num_snakes
        self.size = size
        self.initial_snake_length = initial_snake_length
        self.on_death = on_death
        self.device = device
        self.verbose = verbose
        self.dtype = dtype
        self.observation_mode = observation_mode
        if torch.device(device).type != 'cuda':
            raise RuntimeError(f"wurm_b200 envs run on CUDA devices only (got device={device!r}); "
                               "the CPU implementation of this path is the reference itself")
        if dtype != torch.float:
            raise NotImplementedError('wurm_b200.MultiSnake keeps the state in float32 only')
        if state is None:
            state = os.environ.get('WURM_B200_STATE', 'dense')     # process-wide default (how the reference's own tests are
                                                                   # run against the compact state without touching them)
        if state == 'dense' and os.environ.get('WURM_B200_SHADOW', '1') == '0':
            state = 'dense_scan'
        if state not in ('dense', 'dense_scan', 'compact'):
            raise ValueError("state must be 'dense' (the reference's fp32 tensors are the state; the library shadows them with "
                             "its own records), 'dense_scan' (the same without the shadow: every call streams the tensors) or 'compact'")
        # state='compact' (an extension): between calls the env lives in HBM as one 32-bit record per cell (include/
        # wurm_b200.h, WurmMultiState.cells) instead of the reference's (1+2K) fp32 grids that are ~99 % zeros; `foods`,
        # `heads` and `bodies` are then materialised on attribute access and folded back in if the caller wrote to them.
        self._compact = state == 'compact'
        if num_snakes > _lib.MULTI_MAX_SNAKES:
            raise NotImplementedError(f'at most {_lib.MULTI_MAX_SNAKES} snakes per environment')
        if observation_mode.startswith('partial_'):
            self.observation_width = int(observation_mode.split('_')[1])
            self.observation_size = 2 * int(observation_mode.split('_')[1]) + 1

        if render_args is None:
            self.render_args = {'num_rows': 1, 'num_cols': 1, 'size': 256}
        else:
            self.render_args = render_args

        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, ()).item())    # follows torch.manual_seed
        self.seed = seed
        self._draws = 0
        # device-side addend of the call counter: stays 0 in normal use, bumped between CUDA-graph replays
        self._draws_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        # head cell per snake left by one kernel call for the next (-1 dead, -2 unknown): verified hints that let a call
        # skip streaming the heads tensor (include/wurm_b200.h, WurmMultiState.head_hints)
        self._head_hints = torch.full((num_envs * num_snakes,), -2, dtype=torch.short, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._stats = torch.zeros((_lib.STATS_SLOTS, _lib.STATS_FIELDS), dtype=torch.int64, device=self.device)

        E, K, S = num_envs, num_snakes, size
        self.dones = torch.zeros(E * K, dtype=torch.bool, device=self.device)
        self._dev = self.dones.device
        self._dense_key = None
        if self._compact:
            self._cells = torch.zeros((E, (S * S + 3) & ~3), dtype=torch.int32, device=self.device)
            self._head_hints.fill_(-1)               # authoritative in this mode: no heads yet
            self._dense = None                       # fp32 tensors exist only while a caller looks at them
        else:
            # The reference's tensors stay the state; the library keeps the kernel's own records BESIDE them (4 B per cell, 1 / (1+2K)
            # of the tensors' size) so that a step need not stream ~99 % zeros to find the snakes: it loads the records, verifies
            # them against the tensors and writes its changes to both (include/wurm_b200.h, WurmMultiState.cells_valid).  The
            # records are only trusted while torch's version counters say nobody wrote to the tensors (_state()).
            # state='dense_scan' (or WURM_B200_SHADOW=0) switches this off: every call then streams the tensors.
            shadow = state == 'dense'
            self._cells = torch.zeros((E, (S * S + 3) & ~3), dtype=torch.int32, device=self.device) if shadow else None
            self._dense = {'foods': torch.zeros((E, 1, S, S), dtype=self.dtype, device=self.device),
                           'heads': torch.zeros((E * K, 1, S, S), dtype=self.dtype, device=self.device),
                           'bodies': torch.zeros((E * K, 1, S, S), dtype=self.dtype, device=self.device)}
        self.boost_this_step = torch.zeros(E * K, dtype=torch.bool, device=self.device)
        self.rewards = torch.zeros(E * K, dtype=torch.float, device=self.device)
        self.env_lifetimes = torch.zeros(E, dtype=torch.long, device=self.device)
        self.snake_lifetimes = torch.zeros((E, K), dtype=torch.long, device=self.device)
        self.orientations = torch.zeros(E * K, dtype=torch.long, device=self.device)
        self.viewer = None

        # Environment dynamics parameters (mutable; read at every call)
        self.respawn_mode = respawn_mode
        self.food_on_death_prob = food_on_death_prob
        self.boost = boost
        self.boost_cost_prob = boost_cost_prob
        self.food_mode = food_mode
        self.food_rate = food_rate
        self.max_food = self.num_snakes * 8
        self.max_env_lifetime = 5000
        self.reward_on_death = reward_on_death

        # Rendering parameters
        self.self_colour = torch.tensor((0, 192, 0), dtype=torch.short, device=self.device)
        self.self_boost_colour = torch.tensor((0, 255, 0), dtype=torch.short, device=self.device)
        self.other_colour = torch.tensor((0, 0, 192), dtype=torch.short, device=self.device)
        self.other_boost_colour = torch.tensor((0, 0, 255), dtype=torch.short, device=self.device)
        self.food_colour = torch.tensor((255, 0, 0), dtype=torch.short, device=self.device)
        self.edge_colour = torch.tensor((0, 0, 0), dtype=torch.short, device=self.device)

        if agent_colours == 'random':
            self.colour_mode = 'random'
            self.agent_colours = self.get_n_colours(E * K)
        elif agent_colours == 'fixed':
            self.colour_mode = 'fixed'
            self.agent_colours = self.get_n_colours(K).repeat(E, 1)
        else:
            raise ValueError('agent_colours must in {random, fixed}')
        self.num_colours = self.agent_colours.shape[0]

        self.info = {}

        self.edge_locations_mask = torch.zeros((1, 1, S, S), dtype=self.dtype, device=self.device)
        self.edge_locations_mask[:, :, :1, :] = 1
        self.edge_locations_mask[:, :, :, :1] = 1
        self.edge_locations_mask[:, :, -1:, :] = 1
        self.edge_locations_mask[:, :, :, -1:] = 1

        self._names = {kind: tuple(f'{kind}_{i}' for i in range(K))
                       for kind in ('agent', 'boost', 'snake_collision', 'edge_collision', 'food', 'size')}
        self._hint_key = None
        self._shadow_ok = False                      # dense mode: the records describe the tensors' current content
        if not manual_setup:
            self._create_all()
        self._adopt_state()

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def get_n_colours(self, n: int) -> torch.Tensor:
        """Random agent colours (reference :163-169); construction-time helper, not on the hot path."""
        colours = torch.rand((n, 3), device=self.device)
        colours[:, 0] /= 1.5
        colours /= colours.norm(2, dim=1, keepdim=True)
        colours *= 192
        return colours.short()

    def _log(self, msg: str):
        if self.verbose > 0:
            print(msg)

    def _cfg(self, mode=None):
        # (the rule attributes are mutable and re-read at every call; the struct is rebuilt only when one of them moved)
        key = (mode, self.food_on_death_prob, self.boost, self.boost_cost_prob, self.food_mode, self.food_rate,
               self.reward_on_death, self.respawn_mode, self.colour_mode, self.num_envs, self.num_snakes, self.size)
        hit = self._cfg_cache.get(key)
        if hit is None:
            if len(self._cfg_cache) > 64:
                self._cfg_cache.clear()
            hit = self._cfg_cache[key] = self._make_cfg(mode)
        return hit

    def _make_cfg(self, mode=None):
        if mode is None:
            obs_mode, n = _lib.MOBS_NONE, 0
        elif mode == 'full':
            obs_mode, n = _lib.MOBS_FULL, 0
        elif mode.startswith('partial_'):
            obs_mode, n = _lib.MOBS_PARTIAL, int(mode.split('_')[1])
        else:
            raise ValueError('Unrecognised observation mode.')
        if self.food_mode not in ('only_one', 'random_rate'):
            raise ValueError('food_mechanics not recognised')
        f32 = np.float32
        return _lib.WurmMultiCfg(
            self.num_envs, self.num_snakes, self.size, obs_mode, n, int(bool(self.boost)),
            int(self.food_on_death_prob > 0), float(f32(1 - self.food_on_death_prob)), float(f32(self.boost_cost_prob)),
            0 if self.food_mode == 'only_one' else 1, float(f32(self.food_rate)), float(self.reward_on_death),
            int(self.respawn_mode == 'any'), int(self.colour_mode == 'random'))

    def _norm(self, name, dtype, shape):
        t = getattr(self, name)
        if t.dtype != dtype or not t.is_contiguous() or t.device.type != 'cuda':
            t = t.to(device=self.device, dtype=dtype).contiguous()
            setattr(self, name, t)
        if tuple(t.shape) != shape:
            raise RuntimeError(f'{name} has shape {tuple(t.shape)}, expected {shape}')
        return t

    # ---- the reference's state tensors: plain tensors in dense mode, materialised on access in compact mode ----
    def _get_dense(self, name):
        if self._dense is None:
            self._materialise()
        return self._dense[name]

    def _set_dense(self, name, value):
        if self._dense is None:
            self._materialise()
        self._dense[name] = value
        self._dense_key = None               # compact mode: an assigned tensor is always folded back in before the next call

    foods = property(lambda self: self._get_dense('foods'), lambda self, v: self._set_dense('foods', v))
    heads = property(lambda self: self._get_dense('heads'), lambda self, v: self._set_dense('heads', v))
    bodies = property(lambda self: self._get_dense('bodies'), lambda self, v: self._set_dense('bodies', v))

    def _dense_identity(self):
        return tuple((t.data_ptr(), t._version) for t in self._dense.values())

    def _compact_struct(self, dense):
        return _lib.WurmMultiState(
            _ptr(dense['foods']) if dense else None, _ptr(dense['heads']) if dense else None,
            _ptr(dense['bodies']) if dense else None, _ptr(self.dones), _ptr(self.orientations), _ptr(self.boost_this_step),
            _ptr(self.agent_colours), _ptr(self._head_hints), _ptr(self._cells), 0)

    def _materialise(self):
        """compact records -> the reference's three fp32 tensors (one launch), remembered until the next state-changing
        call so that repeated reads cost nothing."""
        E, K, S = self.num_envs, self.num_snakes, self.size
        dense = {'foods': torch.empty((E, 1, S, S), dtype=torch.float32, device=self._dev),
                 'heads': torch.empty((E * K, 1, S, S), dtype=torch.float32, device=self._dev),
                 'bodies': torch.empty((E * K, 1, S, S), dtype=torch.float32, device=self._dev)}
        cfg = self._cfg(None)
        st = self._compact_struct(dense)
        with _lib.device_guard(self._dev):
            _lib.check(self._lib.wurm_multi_expand(ctypes.byref(cfg), ctypes.byref(st), self._stream()))
        self._dense = dense
        self._dense_key = self._dense_identity()

    def _compress(self):
        """the (caller-edited) fp32 tensors -> compact records; raises if the records cannot carry the state."""
        cfg = self._cfg(None)
        st = self._compact_struct(self._dense)
        with _lib.device_guard(self._dev):
            _lib.check(self._lib.wurm_multi_compact(ctypes.byref(cfg), ctypes.byref(st), _ptr(self._status), self._stream()))
        self._dense_key = self._dense_identity()
        st_word = int(self._status.item())
        if st_word & _lib.ST_NOT_COMPACT:
            self._status.zero_()
            raise RuntimeError("state='compact' cannot carry this state exactly (food / head values other than 1, non-integral "
                               "body values, two bodies on one cell, or a head off its own body); use state='dense'")

    def _before_replay(self):
        """GraphedStepper hook: what step() does before launching -- compact mode: fold caller edits of the materialised
        tensors in, then drop them; dense mode: if the caller wrote to a state tensor, mark every hint 'unknown' on the device,
        which is what makes the captured launch re-load those envs from the tensors."""
        self._state(mutates=True)
        if not self._compact:
            self._shadow_ok = self._cells is not None

    def _snapshot_names(self):
        """Attributes that make up the env's state (GraphedStepper snapshots them around its warm-up)."""
        small = ('dones', 'orientations', 'boost_this_step', 'agent_colours', 'rewards', '_head_hints', '_stats', '_status')
        if self._compact:
            return ('_cells',) + small
        return ('foods', 'heads', 'bodies') + small       # (the shadow records are re-derived: _sync_shadow)

    def _sync_shadow(self):
        """dense mode: derives the shadow records from the tensors now (one conversion launch and one host sync) instead of
        letting the next step emit them.  GraphedStepper calls it before capturing, so that the captured launch is the
        record-loading one.  A state the records cannot carry leaves the shadow off until the next step."""
        if self._compact or self._cells is None:
            return
        st = self._state()
        if self._shadow_ok:
            return
        cfg = self._cfg(None)
        with _lib.device_guard(self._dev):
            _lib.check(self._lib.wurm_multi_compact(ctypes.byref(cfg), ctypes.byref(st), _ptr(self._status), self._stream()))
        st_word = int(self._status.item())
        if st_word & _lib.ST_NOT_COMPACT:
            self._status.fill_(st_word & ~_lib.ST_NOT_COMPACT)
            self._head_hints.fill_(-2)
        else:
            self._shadow_ok = True

    def _hint_identity(self):
        return tuple((t.data_ptr(), t._version) for t in (self.heads, self.bodies, self.foods, self.dones))

    def _adopt_state(self):
        """Records the identity of the state tensors as what the head hints describe (after this env's own kernels)."""
        if self._compact:
            if self._dense is not None:
                self._dense_key = self._dense_identity()
            return
        self._hint_key = self._hint_identity()

    def invalidate_hints(self):
        """Drops the per-snake head-cell hints: the next call streams the heads tensor again.  Called automatically
        when a state tensor was replaced or written through torch since this env's last own call (the kernels verify
        that a hinted cell still holds a head, not that it is the snake's ONLY head cell); call it by hand after
        writing the state through a raw pointer, which torch's version counter cannot see."""
        if self._compact:                            # the head cells are state in this mode, not hints: "forget what you
            self._dense_key = None                   # assumed" means "fold the materialised tensors back in"
            return
        self._head_hints.fill_(-2)                   # (also what tells a captured record-loading launch to re-load from the tensors)
        self._shadow_ok = False
        self._adopt_state()

    def _state(self, mutates=False):
        """The state attributes may have been replaced by the caller (tests assign them): normalise.  Also where the
        head hints are dropped if the caller touched a state tensor since this env's last own call.  `mutates`: the
        call about to be made changes the state (compact mode: materialised fp32 tensors go stale)."""
        E, K, S = self.num_envs, self.num_snakes, self.size
        dones = self._norm('dones', torch.bool, (E * K,))
        small = (_ptr(dones), _ptr(self._norm('orientations', torch.long, (E * K,))),
                 _ptr(self._norm('boost_this_step', torch.bool, (E * K,))), _ptr(self._norm('agent_colours', torch.short, (E * K, 3))))
        if self._compact:
            if self._dense is not None:
                for name, shape in (('foods', (E, 1, S, S)), ('heads', (E * K, 1, S, S)), ('bodies', (E * K, 1, S, S))):
                    self._norm(name, torch.float32, shape)
                if self._dense_identity() != self._dense_key:
                    self._compress()                 # the caller wrote to (or replaced) a materialised tensor
                if mutates:
                    self._dense = None
            return _lib.WurmMultiState(None, None, None, *small, _ptr(self._head_hints), _ptr(self._cells), 1)
        foods = self._norm('foods', torch.float32, (E, 1, S, S))
        heads = self._norm('heads', torch.float32, (E * K, 1, S, S))
        bodies = self._norm('bodies', torch.float32, (E * K, 1, S, S))
        if self._hint_identity() != self._hint_key:
            self.invalidate_hints()
        return _lib.WurmMultiState(_ptr(foods), _ptr(heads), _ptr(bodies), *small, _ptr(self._head_hints), _ptr(self._cells),
                                   int(self._shadow_ok))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)

    def _obs_shape(self, cfg):
        E, K, S = self.num_envs, self.num_snakes, self.size
        w = 2 * cfg.obs_n + 1
        return (K, E, 3, S, S) if cfg.obs_mode == _lib.MOBS_FULL else (K, E, 3, w, w)

    def stats(self, reduce_group=None):
        """Episode statistics accumulated on the device since construction (env_steps, episodes = envs in
        which every snake died, reward = food eaten, self_collisions = snake collisions, edge_collisions);
        with `reduce_group` summed over ranks by one small all-reduce."""
        totals = self._stats.sum(dim=0)
        if reduce_group is not None:
            from ..distributed import all_reduce_stats
            all_reduce_stats(totals, None if reduce_group is True else reduce_group)
        return dict(zip(_lib.STAT_NAMES, totals.tolist()))

    def check_status(self):
        """Raises if a kernel met a condition it reports through the status word since the last check."""
        st = int(self._status.item())
        if st:
            self._status.zero_()
            if st & _lib.ST_NO_SPAWN:
                raise RuntimeError('There is no available locations to create snake!')     # reference :865,947
            msgs = []
            if st & _lib.ST_OVERLAP:
                msgs.append('an environment contains overlapping snakes at the start of a step')
            if st & _lib.ST_MULTI_HEAD:
                msgs.append('a snake has more than one head cell')
            raise RuntimeError('; '.join(msgs) or f'status {st}')

    # ------------------------------------------------------------------------------------------
    # observations
    # ------------------------------------------------------------------------------------------
    def _observe_tensor(self, mode):
        cfg = self._cfg(mode)
        st = self._state()
        obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=self._dev)
        with _lib.device_guard(self._dev):
            _lib.check(self._lib.wurm_multi_observe(ctypes.byref(cfg), ctypes.byref(st), _ptr(obs), _ptr(self._status),
                                                    self._stream()))
        return obs

    def _observe(self, mode: str = None) -> Dict[str, torch.Tensor]:
        if mode is None:
            mode = self.observation_mode
        obs = self._observe_tensor(mode)
        return OrderedDict([(f'agent_{i}', obs[i]) for i in range(self.num_snakes)])

    def _observe_agent(self, agent: int) -> torch.Tensor:
        return self._observe_tensor('full')[agent]

    def _get_env_images(self):
        """int16 (E,3,S,S) image of every env (reference :194-227)."""
        cfg = self._cfg(None)
        st = self._state()
        img = torch.empty((self.num_envs, 3, self.size, self.size), dtype=torch.short, device=self._dev)
        with _lib.device_guard(self._dev):
            _lib.check(self._lib.wurm_multi_env_images(ctypes.byref(cfg), ctypes.byref(st), _ptr(img), _ptr(self._status),
                                                       self._stream()))
        return img

    # ------------------------------------------------------------------------------------------
    # step
    # ------------------------------------------------------------------------------------------
    def step(self, actions: Dict[str, torch.Tensor], *, draws: dict = None, auto_reset: bool = False,
             reset_draws: dict = None):
        """reference :462-731.  `auto_reset=True` (an extension) fuses the `reset(dones['__all__'],
        return_observations=False)` the reference's driver issues right after every step
        (experiments/multiagent.py:377) into the same launch; outputs and observations are the step's, the
        state afterwards is the one after the reset; bit-identical to the two calls."""
        if len(actions) != self.num_snakes:
            raise RuntimeError('Must have a Tensor of actions for each snake')

        for agent, act in actions.items():
            if act.dtype not in _ACTION_BYTES:     # the reference's three integer types, plus uint8 (an extension)
                raise TypeError('actions Tensor must be an integer type i.e. '
                                '{torch.ShortTensor, torch.IntTensor, torch.LongTensor}')

            if act.shape[0] != self.num_envs:
                raise RuntimeError('Must have the same number of actions as environments.')

        t0 = time()
        E, K = self.num_envs, self.num_snakes
        st = self._state(mutates=True)
        dev = self._dev
        acts = list(actions.values())                     # dict order, key names are never parsed (reference :482)
        dtype = acts[0].dtype
        if any(a.dtype != dtype for a in acts):
            dtype = torch.long
        acts = [a if (a.device == dev and a.dtype == dtype and a.is_contiguous())
                else a.to(device=dev, dtype=dtype, non_blocking=True).contiguous() for a in acts]
        act_ptrs = (ctypes.c_void_p * K)(*[a.data_ptr() for a in acts])

        cfg = self._cfg(self.observation_mode)
        obs = torch.empty(self._obs_shape(cfg), dtype=torch.float32, device=dev)
        floats = torch.empty((3, E, K), dtype=torch.float32, device=dev)  # one allocation each for the float and the flag outputs
        rewards, food, size = floats[0], floats[1], floats[2]
        flags = torch.empty((4, E, K), dtype=torch.bool, device=dev)     # snake_collision, edge_collision, dones, boost
        all_done = torch.empty(E, dtype=torch.bool, device=dev)
        out = _lib.WurmMultiStepOut(_ptr(rewards), _ptr(flags[0]), _ptr(flags[1]), _ptr(food), _ptr(size), _ptr(flags[2]),
                                    _ptr(flags[3]), _ptr(all_done), _ptr(obs))
        dr, keep = None, []
        if draws is not None:
            def dev_t(key, dt):
                t = draws.get(key)
                if t is None:
                    return None
                t = torch.as_tensor(t).to(device=dev, dtype=dt).contiguous()
                keep.append(t)
                return t.data_ptr()
            dr = _lib.WurmMultiStepDraws(int(bool(draws['boost_phase_ran'])), dev_t('u_boost', torch.float32),
                                         dev_t('u_cost', torch.float32), dev_t('u_reg', torch.float32),
                                         dev_t('food_cell', torch.int32), dev_t('u_rate', torch.float32))
        rdr = None
        if auto_reset and reset_draws is not None:
            def dev_r(key, dt):
                t = torch.as_tensor(reset_draws[key]).to(device=dev, dtype=dt).contiguous()
                keep.append(t)
                return t.data_ptr()
            rdr = _lib.WurmMultiResetDraws(dev_r('create', torch.int32), dev_r('respawn', torch.int32),
                                           dev_r('colours', torch.short))
        self._draws += 1
        with _lib.device_guard(dev):
            if auto_reset:
                _lib.check(self._lib.wurm_multi_step_reset(
                    ctypes.byref(cfg), ctypes.byref(st), act_ptrs, _ACTION_BYTES[dtype],
                    ctypes.byref(dr) if dr is not None else None, ctypes.byref(rdr) if rdr is not None else None,
                    self.seed, self._draws, _ptr(self._draws_dev), ctypes.byref(out), _ptr(self._status), _ptr(self._stats),
                    self._stream()))
                self._draws += 1            # the fused reset consumed the next counter value
            else:
                _lib.check(self._lib.wurm_multi_step(
                    ctypes.byref(cfg), ctypes.byref(st), act_ptrs, _ACTION_BYTES[dtype],
                    ctypes.byref(dr) if dr is not None else None, self.seed, self._draws, _ptr(self._draws_dev),
                    ctypes.byref(out), _ptr(self._status), _ptr(self._stats), self._stream()))

        if not self._compact:
            self._shadow_ok = self._cells is not None      # every step leaves the records describing the new state
        self.rewards = rewards.view(E * K)
        self._step_dones = flags[2]          # (E,K) copy of the done flags owned by this step's outputs
        # The reference's per-agent dicts (:686-731) are views of the (K,E,..) / (E,K) outputs.  One unbind() per tensor makes
        # the K views in a single call: 8 K Python-level slices took three times as long, of the order of the kernel's own
        # 0.3 ms at 16 snakes (profiles/r02_host_overhead.txt).
        names = self._names
        observations = OrderedDict(zip(names['agent'], obs.unbind(0)))
        dones = dict(zip(names['agent'], flags[2].unbind(1)))
        dones['__all__'] = all_done
        rewards_dict = dict(zip(names['agent'], rewards.unbind(1)))
        self.info = info = {}
        for nb, ns, ne, b, sc, ec in zip(names['boost'], names['snake_collision'], names['edge_collision'],
                                         flags[3].unbind(1), flags[0].unbind(1), flags[1].unbind(1)):
            info[nb] = b
            info[ns] = sc
            info[ne] = ec
        info.update(zip(names['food'], food.unbind(1)))
        info.update(zip(names['size'], size.unbind(1)))
        if self.verbose > 0:
            torch.cuda.synchronize(dev)
            self._log(f'step: {1000 * (time() - t0)}ms')
        return observations, rewards_dict, dones, self.info

    # ------------------------------------------------------------------------------------------
    # reset
    # ------------------------------------------------------------------------------------------
    def _reset_mask(self, env_done, draws=None):
        cfg = self._cfg(None)
        st = self._state(mutates=True)
        dev = self._dev
        dr, keep = None, []
        if draws is not None:
            def dev_t(key, dt):
                t = torch.as_tensor(draws[key]).to(device=dev, dtype=dt).contiguous()
                keep.append(t)
                return t.data_ptr()
            dr = _lib.WurmMultiResetDraws(dev_t('create', torch.int32), dev_t('respawn', torch.int32),
                                          dev_t('colours', torch.short))
        self._draws += 1
        with _lib.device_guard(dev):
            _lib.check(self._lib.wurm_multi_reset(ctypes.byref(cfg), ctypes.byref(st), _ptr(env_done),
                                                  ctypes.byref(dr) if dr is not None else None, self.seed, self._draws,
                                                  _ptr(self._draws_dev), _ptr(self._status), self._stream()))

    def _create_all(self, draws=None):
        """__init__'s _create_envs(num_envs) (reference :113, :996-1019)."""
        self._reset_mask(torch.ones(self.num_envs, dtype=torch.bool, device=self.device), draws)
        self.check_status()      # RuntimeError('There is no available locations to create snake!') like :947

    def reset(self, done: torch.Tensor = None, return_observations: bool = True, *,
              draws: dict = None) -> Optional[Dict[str, torch.Tensor]]:
        """Re-creates the envs flagged in `done`, re-colours dead snakes, respawns one dead snake per
        env in respawn_mode 'any' (reference :771-836)."""
        t0 = time()
        if done is None:
            done = self.dones.view(self.num_envs, self.num_snakes).all(dim=1)

        done = done.view((done.shape[0]))
        if done.shape[0] != self.num_envs:
            raise RuntimeError('Must have one done flag per environment.')
        dev = self._dev
        if not (done.dtype == torch.bool and done.device == dev and done.is_contiguous()):
            done = (done != 0).to(device=dev).contiguous()
        self._reset_mask(done, draws)
        # env_lifetimes is never incremented by the reference (:106,128,705,797), so there is nothing to clear
        if self.verbose > 0:
            torch.cuda.synchronize(dev)
            self._log(f'reset: {1000 * (time() - t0)}ms')

        if return_observations:
            return self._observe()

    # ------------------------------------------------------------------------------------------
    # invariants and display (off the hot path)
    # ------------------------------------------------------------------------------------------
    def check_consistency(self):
        """The reference's invariants (reference :733-769 + wurm/utils.py:113-164 snake_consistency on the
        living snakes), as one fused kernel and one host sync."""
        cfg = self._cfg(None)
        st = self._state()
        dev = self._dev
        report = torch.tensor([0, 0, 2 ** 31 - 1, 0], dtype=torch.int32, device=dev)
        with _lib.device_guard(dev):
            _lib.check(self._lib.wurm_multi_check(ctypes.byref(cfg), ctypes.byref(st), _ptr(report), self._stream()))
        _lib.raise_on_report(report.tolist())
        self.check_status()

    def render(self, mode: str = 'human', env: int = None):
        """Human display of `_get_env_images()` (reference :229-266); host-side convenience, outside the hot path."""
        from ._display import show
        return show(self, self._get_env_images(), mode, env)
