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
