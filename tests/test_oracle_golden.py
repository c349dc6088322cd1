"""CPU: the oracle (oracle/wurm_oracle.c) reproduces the reference's golden vectors bit for bit.

The vectors in tests/golden/ were produced by the unmodified reference (oracle/gen_golden.py); this
is what pins the oracle on machines without the reference tree.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from golden_util import load, assert_same

SINGLE = load('single.npz')


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, out in kat:
        assert list(orc.philox(ctr, key)) == out


@pytest.mark.parametrize('i', range(len(SINGLE)))
def test_single_golden(i):
    tr = SINGLE[i]
    N, S, mode = tr.N, tr.S, tr.mode
    if 'init_spawn' in tr:
        state = np.zeros((N, 3, S, S), np.float32)
        orc.single_reset(state, np.ones(N, np.uint8), tr['init_spawn'])
        assert_same(state, tr['init_envs'].astype(np.float32), 'created envs')
    else:
        state = tr['init_envs'].astype(np.float32)
    for t in range(tr.steps):
        a = tr[f'{t}/actions_in'].copy()
        r, d, sc, ec = orc.single_step(state, a, tr[f'{t}/food_cell'])
        tag = f'trajectory {i} ({mode}, S={S}) step {t}: '
        assert_same(state, tr[f'{t}/envs'].astype(np.float32), tag + 'envs')
        assert_same(a, tr[f'{t}/actions_out'], tag + 'sanitised actions')
        assert_same(r, tr[f'{t}/reward'], tag + 'reward')
        assert_same(d, tr[f'{t}/done'], tag + 'done')
        assert_same(sc, tr[f'{t}/self_collision'], tag + 'self_collision')
        assert_same(ec, tr[f'{t}/edge_collision'], tag + 'edge_collision')
        o, bad = orc.single_observe(state, mode)
        assert bad == 0
        assert_same(o, tr[f'{t}/obs'], tag + 'observation')
        if bool(tr[f'{t}/did_reset']):
            orc.single_reset(state, d, tr[f'{t}/spawn'])
            assert_same(state, tr[f'{t}/reset_envs'].astype(np.float32), tag + 'envs after reset')
            o, bad = orc.single_observe(state, mode)
            assert_same(o, tr[f'{t}/reset_obs'], tag + 'observation after reset')


def test_reference_scenarios_known_answers():
    """The head tracks asserted by the reference's own tests (tests/test_single_snake_env.py:56-63,
    175-180) hold for the oracle on the reference's fixture (wurm/utils.py:68-110, 'up', size 12)."""
    S = 12
    env = np.zeros((1, 3, S, S), np.float32)
    for (y, x), v in {(3, 3): 1, (3, 4): 2, (4, 4): 3, (5, 4): 4}.items():
        env[0, 2, y, x] = v
    env[0, 1, 5, 4] = 1
    env[0, 0, 6, 6] = 1
    for actions, track in [([0, 0, 3, 0, 0, 1], [(6, 4), (7, 4), (7, 5), (8, 5), (9, 5), (9, 4)]),
                           ([2, 2, 2, 3], [(6, 4), (7, 4), (8, 4), (8, 5)])]:
        state = env.copy()
        for a, (hy, hx) in zip(actions, track):
            _, d, _, _ = orc.single_step(state, np.array([a], np.int64), np.array([-1], np.int32))
            assert not d[0]
            assert np.argmax(state[0, 1]) == hy * S + hx
    # [1]*10 hits the boundary (:119-141); [0,3,3,2,1,0,0,0] hits itself (:143-169)
    state = env.copy()
    assert any(orc.single_step(state, np.array([1], np.int64), np.array([-1], np.int32))[1][0] for _ in range(10))
    state = env.copy()
    dones = []
    for a in [0, 3, 3, 2, 1, 0, 0, 0]:
        r, d, sc, ec = orc.single_step(state, np.array([a], np.int64))
        dones.append((d[0], sc[0]))
        if d[0]:
            break
    assert dones[-1] == (1, 1)
    assert state[0, 0].sum() == 1
