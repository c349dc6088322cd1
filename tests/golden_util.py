"""Access to tests/golden/*.npz (generated from the reference by oracle/gen_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def bits(a):
    """fp32 arrays are compared as int32 bit patterns: -0.0 != +0.0, NaN-safe."""
    a = np.ascontiguousarray(a)
    return a.view(np.int32) if a.dtype == np.float32 else a


def assert_same(a, b, what=''):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f'{what}: shape {a.shape} != {b.shape}'
    if not np.array_equal(bits(a), bits(b)):
        bad = np.flatnonzero(bits(a).reshape(-1) != bits(b).reshape(-1))
        raise AssertionError(f'{what}: {len(bad)} of {a.size} elements differ, first at flat index {bad[0]}: '
                             f'{a.reshape(-1)[bad[0]]!r} vs {b.reshape(-1)[bad[0]]!r}')


class Trajectory(object):
    def __init__(self, z, i):
        self._z, self._i = z, i

    def __getitem__(self, key):
        return self._z[f'{self._i}/{key}']

    def __contains__(self, key):
        return f'{self._i}/{key}' in self._z.files

    @property
    def N(self): return int(self['N'])

    @property
    def S(self): return int(self['S'])

    @property
    def mode(self): return str(self['mode'])

    @property
    def steps(self): return int(self['steps'])


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return [Trajectory(z, i) for i in range(int(z['count']))]


# ---- MultiSnake trajectories (tests/golden/multi.npz) ----
RULE_DEFAULTS = dict(food_on_death_prob=0.5, boost=True, boost_cost_prob=0.5, food_mode='only_one', food_rate=5e-4,
                     respawn_mode='all', reward_on_death=-1, agent_colours='random')
STATE_FIELDS = ('foods', 'heads', 'bodies', 'dones', 'orientations', 'boost_this_step', 'agent_colours')


def multi_rules(tr):
    """Constructor kwargs of the reference env the trajectory was recorded from."""
    rules = dict(RULE_DEFAULTS)
    for k in rules:
        if f'rule/{k}' in tr:
            v = tr[f'rule/{k}']
            rules[k] = v.item() if v.dtype.kind in 'fib' else str(v)
    return rules


def multi_group(tr, prefix, keys):
    return {k: (tr[f'{prefix}/{k}'] if f'{prefix}/{k}' in tr else None) for k in keys}


def multi_step_draws(tr, t, dense_rate):
    """The recorded draws of step t; `dense_rate`: scatter the random_rate rows to (E,S,S) for the CUDA library."""
    d = multi_group(tr, f'{t}/draws', ('boost_phase_ran', 'u_boost', 'u_cost', 'u_reg', 'food_cell', 'u_rate', 'selected'))
    d['boost_phase_ran'] = bool(d['boost_phase_ran'])
    if dense_rate and d['u_rate'] is not None:
        E, S = int(tr['E']), int(tr['S'])
        dense = np.ones((E, S, S), np.float32)
        dense[np.flatnonzero(d['selected'])] = d['u_rate']
        d['u_rate'] = dense
    return d


def multi_state_arrays(tr, prefix):
    """State arrays in the dtypes of the live tensors."""
    st = multi_group(tr, prefix, STATE_FIELDS)
    for k in ('foods', 'heads', 'bodies'):
        st[k] = st[k].astype(np.float32)
    return st
