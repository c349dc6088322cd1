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
