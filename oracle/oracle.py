"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/wurm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (wurm_b200/) never does; it fails loudly without its CUDA library.

All arrays are numpy, in the reference's layouts.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'wurm_oracle.c')
_SO = os.path.join(_HERE, '_build', 'libwurm_oracle.so')

OBS_MODES = {'default': 0, 'raw': 1, 'one_channel': 2, 'positions': 3, 'partial': 4}


def build(force=False):
    """gcc -O2 -fopenmp -shared oracle/wurm_oracle.c -> oracle/_build/libwurm_oracle.so"""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(['gcc', '-O2', '-std=c99', '-fopenmp', '-fPIC', '-shared', '-Wall', '-o', _SO, _SRC, '-lm'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.wurm_oracle_single_observe.restype = ctypes.c_int
        _lib.wurm_oracle_multi_step.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def philox(ctr, key):
    ctr = _c(ctr, np.uint32); key = _c(key, np.uint32); out = np.zeros(4, np.uint32)
    lib().wurm_oracle_philox(_p(ctr), _p(key), _p(out))
    return out


def parse_obs_mode(mode):
    """'partial_3' -> (4, 3); 'default' -> (0, 0)"""
    if mode.startswith('partial_'):
        return OBS_MODES['partial'], int(mode.split('_')[-1])
    return OBS_MODES[mode], 0


def single_obs_shape(N, S, mode):
    m, n = parse_obs_mode(mode)
    return {0: (N, 3, S, S), 1: (N, 3, S, S), 2: (N, 1, S, S), 3: (N, 4), 4: (N, 3 * (2 * n + 1) ** 2)}[m]


def single_step(envs, actions, food_cell=None, seed=0, step=0):
    """In place on `envs` (N,3,S,S) f32 and `actions` (N,) i64.  Returns reward, done, self_col, edge_col."""
    assert envs.dtype == np.float32 and envs.flags.c_contiguous
    assert actions.dtype == np.int64 and actions.flags.c_contiguous
    N, _, S, _ = envs.shape
    reward = np.zeros(N, np.float32); done = np.zeros(N, np.uint8)
    sc = np.zeros(N, np.uint8); ec = np.zeros(N, np.uint8)
    if food_cell is not None:
        food_cell = _c(food_cell, np.int32)
    lib().wurm_oracle_single_step(N, S, _p(envs), _p(actions), _p(food_cell), ctypes.c_uint64(seed),
                                  ctypes.c_uint64(step), _p(reward), _p(done), _p(sc), _p(ec))
    return reward, done, sc, ec


def single_reset(envs, done, spawn=None, seed=0, step=0):
    """In place on `envs`.  `spawn` (N,4) i32 rows (y, x, dir, food_cell), read for done envs only."""
    assert envs.dtype == np.float32 and envs.flags.c_contiguous
    N, _, S, _ = envs.shape
    done = _c(done, np.uint8)
    if spawn is not None:
        spawn = _c(spawn, np.int32)
    lib().wurm_oracle_single_reset(N, S, _p(envs), _p(done), _p(spawn), ctypes.c_uint64(seed), ctypes.c_uint64(step))


def single_observe(envs, mode):
    assert envs.dtype == np.float32 and envs.flags.c_contiguous
    N, _, S, _ = envs.shape
    m, n = parse_obs_mode(mode)
    obs = np.zeros(single_obs_shape(N, S, mode), np.float32)
    bad = lib().wurm_oracle_single_observe(N, S, _p(envs), m, n, _p(obs))
    return obs, bad


# ------------------------------------------------------------------------------------------------
# MultiSnake
# ------------------------------------------------------------------------------------------------
class MultiCfg(ctypes.Structure):
    _fields_ = [('num_envs', ctypes.c_int32), ('num_snakes', ctypes.c_int32), ('size', ctypes.c_int32),
                ('boost', ctypes.c_int32), ('food_on_death', ctypes.c_int32), ('death_threshold', ctypes.c_float),
                ('boost_cost_prob', ctypes.c_float), ('food_mode', ctypes.c_int32), ('food_rate', ctypes.c_float),
                ('reward_on_death', ctypes.c_float), ('respawn_any', ctypes.c_int32), ('colour_random', ctypes.c_int32)]


class MultiStepDraws(ctypes.Structure):
    _fields_ = [('replay', ctypes.c_int32), ('boost_phase_ran', ctypes.c_int32), ('u_boost', ctypes.c_void_p),
                ('u_cost', ctypes.c_void_p), ('u_reg', ctypes.c_void_p), ('food_cell', ctypes.c_void_p),
                ('u_rate', ctypes.c_void_p), ('n_rate_rows', ctypes.c_int32)]


class MultiResetDraws(ctypes.Structure):
    _fields_ = [('replay', ctypes.c_int32), ('create', ctypes.c_void_p), ('respawn', ctypes.c_void_p),
                ('colours', ctypes.c_void_p)]


def multi_cfg(E, K, S, food_on_death_prob=0.5, boost=True, boost_cost_prob=0.5, food_mode='only_one', food_rate=5e-4,
              respawn_mode='all', reward_on_death=-1, colour_mode='random'):
    """Same parameter names and defaults as the reference constructor (multi_snake.py:56-75)."""
    f32 = lambda v: float(np.float32(v))
    return MultiCfg(E, K, S, int(bool(boost)), int(food_on_death_prob > 0), f32(1 - food_on_death_prob),
                    f32(boost_cost_prob), {'only_one': 0, 'random_rate': 1}[food_mode], f32(food_rate),
                    float(reward_on_death), int(respawn_mode == 'any'), int(colour_mode == 'random'))


class MultiState(object):
    """The MultiSnake state tensors as numpy arrays in the reference's layouts."""

    def __init__(self, E, K, S):
        self.E, self.K, self.S = E, K, S
        self.foods = np.zeros((E, 1, S, S), np.float32)
        self.heads = np.zeros((E * K, 1, S, S), np.float32)
        self.bodies = np.zeros((E * K, 1, S, S), np.float32)
        self.dones = np.zeros(E * K, np.uint8)
        self.orientations = np.zeros(E * K, np.int64)
        self.boost_this_step = np.zeros(E * K, np.uint8)
        self.agent_colours = np.zeros((E * K, 3), np.int16)

    def copy(self):
        out = MultiState.__new__(MultiState)
        out.E, out.K, out.S = self.E, self.K, self.S
        for name in ('foods', 'heads', 'bodies', 'dones', 'orientations', 'boost_this_step', 'agent_colours'):
            setattr(out, name, getattr(self, name).copy())
        return out


def _keep(*arrays):
    return [a for a in arrays if a is not None]


def multi_step(cfg, st, actions, draws=None, seed=0, step=0):
    """In place on `st`.  actions (E,K) int64.  draws: None (Philox) or a dict with keys boost_phase_ran, u_boost,
    u_cost, u_reg, food_cell, u_rate (compact rows).  Returns a dict of (E,K) arrays + all_done (E) + rate_selected."""
    E, K = cfg.num_envs, cfg.num_snakes
    actions = _c(actions, np.int64)
    assert actions.shape == (E, K)
    out = dict(rewards=np.zeros((E, K), np.float32), snake_collision=np.zeros((E, K), np.uint8),
               edge_collision=np.zeros((E, K), np.uint8), food=np.zeros((E, K), np.float32),
               size=np.zeros((E, K), np.float32), all_done=np.zeros(E, np.uint8), rate_selected=np.zeros(E, np.uint8))
    d = MultiStepDraws()
    hold = []
    if draws is not None:
        d.replay = 1
        d.boost_phase_ran = int(draws['boost_phase_ran'])
        for key, dt in [('u_boost', np.float32), ('u_cost', np.float32), ('u_reg', np.float32), ('food_cell', np.int32),
                        ('u_rate', np.float32)]:
            a = draws.get(key)
            if a is not None:
                a = _c(a, dt); hold.append(a)
                setattr(d, key, a.ctypes.data)
        d.n_rate_rows = 0 if draws.get('u_rate') is None else int(np.asarray(draws['u_rate']).shape[0])
    rc = lib().wurm_oracle_multi_step(ctypes.byref(cfg), _p(st.foods), _p(st.heads), _p(st.bodies), _p(st.dones),
                                      _p(st.orientations), _p(st.boost_this_step), _p(actions), ctypes.byref(d),
                                      ctypes.c_uint64(seed), ctypes.c_uint64(step), _p(out['rewards']),
                                      _p(out['snake_collision']), _p(out['edge_collision']), _p(out['food']),
                                      _p(out['size']), _p(out['all_done']), _p(out['rate_selected']))
    if rc != 0:
        raise RuntimeError(f'wurm_oracle_multi_step: replayed tape inconsistent with the state (rc={rc})')
    return out


def multi_observe(cfg, st, mode):
    """mode 'full' -> (K,E,3,S,S); 'partial_n' -> (K,E,3,W,W); obs[k] is the reference's 'agent_k' tensor."""
    E, K, S = cfg.num_envs, cfg.num_snakes, cfg.size
    if mode == 'full':
        m, n, shape = 0, 0, (K, E, 3, S, S)
    else:
        n = int(mode.split('_')[1]); m = 1; shape = (K, E, 3, 2 * n + 1, 2 * n + 1)
    obs = np.zeros(shape, np.float32)
    lib().wurm_oracle_multi_observe.restype = ctypes.c_int
    bad = lib().wurm_oracle_multi_observe(ctypes.byref(cfg), _p(st.foods), _p(st.heads), _p(st.bodies), _p(st.dones),
                                          _p(st.boost_this_step), _p(st.agent_colours), m, n, _p(obs))
    return obs, bad


def multi_env_images(cfg, st):
    img = np.zeros((cfg.num_envs, 3, cfg.size, cfg.size), np.int16)
    lib().wurm_oracle_multi_env_images(ctypes.byref(cfg), _p(st.foods), _p(st.heads), _p(st.bodies),
                                       _p(st.boost_this_step), _p(st.agent_colours), _p(img))
    return img


def multi_reset(cfg, st, env_done, draws=None, seed=0, step=0):
    """In place on `st`.  draws: None (Philox) or dict(create (E,K+1,2) i32, respawn (E,2) i32, colours (E*K,3) i16).
    Returns the number of snakes that could not be created."""
    env_done = _c(env_done, np.uint8)
    d = MultiResetDraws()
    hold = []
    if draws is not None:
        d.replay = 1
        for key, dt in [('create', np.int32), ('respawn', np.int32), ('colours', np.int16)]:
            a = draws.get(key)
            if a is not None:
                a = _c(a, dt); hold.append(a)
                setattr(d, key, a.ctypes.data)
    lib().wurm_oracle_multi_reset.restype = ctypes.c_int
    return lib().wurm_oracle_multi_reset(ctypes.byref(cfg), _p(st.foods), _p(st.heads), _p(st.bodies), _p(st.dones),
                                         _p(st.orientations), _p(st.agent_colours), _p(env_done), ctypes.byref(d),
                                         ctypes.c_uint64(seed), ctypes.c_uint64(step))


# ------------------------------------------------------------------------------------------------
# SimpleGridworld
# ------------------------------------------------------------------------------------------------
def grid_step(envs, actions, food_cell=None, seed=0, step=0):
    """In place on `envs` (N,2,S,S) f32.  Returns reward, done (= edge_collision)."""
    assert envs.dtype == np.float32 and envs.flags.c_contiguous
    N, _, S, _ = envs.shape
    actions = _c(actions, np.int64)
    reward = np.zeros(N, np.float32); done = np.zeros(N, np.uint8)
    if food_cell is not None:
        food_cell = _c(food_cell, np.int32)
    lib().wurm_oracle_grid_step(N, S, _p(envs), _p(actions), _p(food_cell), ctypes.c_uint64(seed), ctypes.c_uint64(step),
                                _p(reward), _p(done))
    return reward, done


def grid_reset(envs, done, start_location, food_cell=None, seed=0, step=0):
    assert envs.dtype == np.float32 and envs.flags.c_contiguous
    N, _, S, _ = envs.shape
    done = _c(done, np.uint8)
    if food_cell is not None:
        food_cell = _c(food_cell, np.int32)
    lib().wurm_oracle_grid_reset(N, S, _p(envs), _p(done), int(start_location[0]), int(start_location[1]), _p(food_cell),
                                 ctypes.c_uint64(seed), ctypes.c_uint64(step))


def grid_observe(envs, mode):
    assert envs.dtype == np.float32 and envs.flags.c_contiguous
    N, _, S, _ = envs.shape
    m = OBS_MODES[mode]
    shape = {0: (N, 3, S, S), 1: (N, 2, S, S), 3: (N, 4)}[m]
    obs = np.zeros(shape, np.float32)
    lib().wurm_oracle_grid_observe(N, S, _p(envs), m, _p(obs))
    return obs


# ------------------------------------------------------------------------------------------------
# A2C returns (wurm/rl/a2c.py:49-63)
# ------------------------------------------------------------------------------------------------
def a2c_returns(bootstrap, rewards, values, dones, gamma, gae_lambda=None):
    """(T,N) fp32 returns; gae_lambda None -> n-step returns."""
    rewards = _c(rewards, np.float32); values = _c(values, np.float32); dones = _c(dones, np.uint8)
    bootstrap = _c(bootstrap, np.float32)
    T, N = rewards.shape
    out = np.zeros((T, N), np.float32)
    lib().wurm_oracle_a2c_returns(T, N, ctypes.c_double(gamma), ctypes.c_double(-1.0 if gae_lambda is None else gae_lambda),
                                  _p(bootstrap), _p(rewards), _p(values), _p(dones), _p(out))
    return out
