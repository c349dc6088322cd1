"""CPU: the oracle (oracle/wurm_oracle.c) reproduces the reference's golden vectors bit for bit.

The vectors in tests/golden/ were produced by the unmodified reference (oracle/gen_golden.py); this
is what pins the oracle on machines without the reference tree.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from golden_util import (load, assert_same, multi_rules, multi_group, multi_step_draws, multi_state_arrays,
                         STATE_FIELDS)

SINGLE = load('single.npz') + load('single_baseline.npz')      # round-1 fixtures + the BASELINE.json geometries
MULTI = load('multi.npz') + load('multi_baseline.npz')        # round-1 fixtures + the BASELINE.json geometries
GRID = load('gridworld.npz')


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


def check_multi_state(st, expect, tag):
    for name in STATE_FIELDS:
        assert_same(getattr(st, name), expect[name], f'{tag}: {name}')


@pytest.mark.parametrize('i', range(len(MULTI)))
def test_multi_golden(i):
    """MultiSnake: creation, step (all outputs, both observation modes, env images) and reset (re-creation,
    re-colouring, respawn) of the oracle equal the reference's recorded trajectories."""
    tr = MULTI[i]
    E, K, S, mode = int(tr['E']), int(tr['K']), int(tr['S']), str(tr['mode'])
    rules = multi_rules(tr)
    colour_mode = rules.pop('agent_colours')
    cfg = orc.multi_cfg(E, K, S, colour_mode=colour_mode, **rules)
    st = orc.MultiState(E, K, S)
    init = multi_state_arrays(tr, 'init')
    st.agent_colours[:] = init['agent_colours']
    assert orc.multi_reset(cfg, st, np.ones(E, np.uint8), dict(create=tr['init/create'], respawn=np.full((E, 2), -1, np.int32),
                                                               colours=init['agent_colours'])) == 0
    check_multi_state(st, init, f'trajectory {i} creation')
    for t in range(int(tr['steps'])):
        tag = f'trajectory {i} ({mode}, K={K}, S={S}, {tr["rules"]}) step {t}'
        out = orc.multi_step(cfg, st, tr[f'{t}/actions'], multi_step_draws(tr, t, dense_rate=False))
        check_multi_state(st, multi_state_arrays(tr, f'{t}/state'), tag)
        assert_same(out['rewards'], tr[f'{t}/rewards'], tag + ': rewards')
        assert_same(st.dones.reshape(E, K), tr[f'{t}/dones'], tag + ': dones')
        assert_same(out['all_done'], tr[f'{t}/all_done'], tag + ': __all__')
        assert_same(out['snake_collision'], tr[f'{t}/snake_collision'], tag + ': snake_collision')
        assert_same(out['edge_collision'], tr[f'{t}/edge_collision'], tag + ': edge_collision')
        assert_same(out['food'], tr[f'{t}/food'], tag + ': food')
        assert_same(out['size'], tr[f'{t}/size'], tag + ': size')
        assert_same(st.boost_this_step.reshape(E, K), tr[f'{t}/boost'], tag + ': boost')
        obs, bad = orc.multi_observe(cfg, st, mode)
        assert bad == 0
        assert_same(obs, tr[f'{t}/obs'], tag + ': observations')
        if f'{t}/env_images' in tr:
            assert_same(orc.multi_env_images(cfg, st), tr[f'{t}/env_images'], tag + ': env images')
        orc.multi_reset(cfg, st, tr[f'{t}/all_done'], multi_group(tr, f'{t}/reset_draws', ('create', 'respawn', 'colours')))
        check_multi_state(st, multi_state_arrays(tr, f'{t}/reset_state'), tag + ' after reset')
        if f'{t}/reset_obs' in tr:
            assert_same(orc.multi_observe(cfg, st, mode)[0], tr[f'{t}/reset_obs'], tag + ': observations after reset')


def multi_test_env(E=1):
    """The reference's two-snake fixture (tests/test_multi_snake_env.py:21-47), size 12."""
    st = orc.MultiState(E, 2, 12)
    h, b = st.heads.reshape(E, 2, 12, 12), st.bodies.reshape(E, 2, 12, 12)
    h[:, 0, 5, 5] = 1; b[:, 0, 5, 5] = 4; b[:, 0, 4, 5] = 3; b[:, 0, 4, 4] = 2; b[:, 0, 4, 3] = 1
    h[:, 1, 8, 7] = 1; b[:, 1, 8, 7] = 4; b[:, 1, 8, 8] = 3; b[:, 1, 8, 9] = 2; b[:, 1, 9, 9] = 1
    st.orientations[:] = np.tile([2, 3], E)          # determine_orientations of the fixture
    return st


def run_multi(st, cfg, agent_0, agent_1):
    E = cfg.num_envs
    for a0, a1 in zip(agent_0, agent_1):
        out = orc.multi_step(cfg, st, np.tile([[a0, a1]], (E, 1)), None, seed=3, step=1)
        yield out


def test_reference_multi_scenarios_known_answers():
    """Known answers asserted by the reference's own tests (tests/test_multi_snake_env.py)."""
    # test_basic_movement :126-176
    cfg = orc.multi_cfg(1, 2, 12)
    st = multi_test_env(); st.foods[0, 0, 1, 1] = 1
    tracks = ([(5, 4), (4, 4), (4, 3), (4, 2), (5, 2), (5, 3)], [(9, 7), (9, 6), (9, 5), (8, 5), (8, 4), (9, 4)])
    for t, out in enumerate(run_multi(st, cfg, [1, 2, 1, 1, 0, 3], [0, 1, 3, 2, 1, 0])):
        for k in range(2):
            assert np.argmax(st.heads[k, 0]) == tracks[k][t][0] * 12 + tracks[k][t][1]
        assert not st.dones.any()
    # test_edge_collision :178-220 (food_on_death_prob = 1; action 6 = boost by a dead snake)
    cfg = orc.multi_cfg(1, 2, 12, food_on_death_prob=1)
    st = multi_test_env(); st.foods[0, 0, 1, 1] = 1
    for t, out in enumerate(run_multi(st, cfg, [1, 1, 1, 1, 1], [0, 2, 2, 6, 2])):
        assert st.dones[0] == (t >= 4) and st.dones[1] == (t >= 2)
        if t == 4:
            assert out['rewards'][0, 0] == -1
        if t == 2:
            assert out['rewards'][0, 1] == -1
    # test_eat_food :285-336: reward at step 0 only, sizes [4,5], food moved
    cfg = orc.multi_cfg(1, 2, 12)
    st = multi_test_env(); st.foods[0, 0, 9, 7] = 1
    for t, out in enumerate(run_multi(st, cfg, [1, 2, 1, 1, 0, 3], [0, 1, 3, 2, 1, 0])):
        if t == 0:
            assert out['rewards'][0, 1] == 1
        assert not st.dones.any()
    assert st.bodies.reshape(2, -1).max(axis=1).tolist() == [4, 5]
    assert st.foods[0, 0, 9, 7] == 0 and st.foods.sum() == 1
    # test_boost_cost :524-555: boost_cost_prob = 1 -> reward -1, sizes [3,4], tail became food at (4,3)
    cfg = orc.multi_cfg(1, 2, 12, boost_cost_prob=1)
    st = multi_test_env(); st.foods[0, 0, 1, 1] = 1
    for t, out in enumerate(run_multi(st, cfg, [4, 1, 2], [0, 1, 3])):
        if t == 0:
            assert out['rewards'][0, 0] == -1
        assert st.bodies.reshape(2, -1).max(axis=1).tolist() == [3, 4]
    assert st.foods[0, 0, 4, 4] == 1                  # test_boost_leaves_food :458
    # test_boost_through_food :398-426: boost_cost_prob = 0, food two cells ahead is eaten in the boost phase
    cfg = orc.multi_cfg(1, 2, 12, boost_cost_prob=0)
    st = multi_test_env(); st.foods[0, 0, 6, 5] = 1
    out = next(run_multi(st, cfg, [4], [0]))
    assert out['rewards'][0, 0] == 1


@pytest.mark.parametrize('i', range(len(GRID)))
def test_gridworld_golden(i):
    """SimpleGridworld oracle == the reference's recorded trajectories (incl. its test scenarios)."""
    tr = GRID[i]
    N, S, mode, start = tr.N, tr.S, tr.mode, tuple(tr['start'])
    if 'init_food' in tr:
        state = np.zeros((N, 2, S, S), np.float32)
        orc.grid_reset(state, np.ones(N, np.uint8), start, tr['init_food'])
        assert_same(state, tr['init_envs'].astype(np.float32), 'created envs')
    else:
        state = tr['init_envs'].astype(np.float32)
    for t in range(tr.steps):
        r, d = orc.grid_step(state, tr[f'{t}/actions'], tr[f'{t}/food_cell'])
        tag = f'gridworld trajectory {i} ({mode}, S={S}) step {t}: '
        assert_same(state, tr[f'{t}/envs'].astype(np.float32), tag + 'envs')
        assert_same(r, tr[f'{t}/reward'], tag + 'reward')
        assert_same(d, tr[f'{t}/done'], tag + 'done')
        assert_same(orc.grid_observe(state, mode), tr[f'{t}/obs'], tag + 'observation')
        orc.grid_reset(state, d, start, tr[f'{t}/reset_food'])
        assert_same(state, tr[f'{t}/reset_envs'].astype(np.float32), tag + 'envs after reset')


def test_a2c_returns_known_answers():
    """Hand-checked recurrences (reference wurm/rl/a2c.py:49-63): n-step R_t = r_t + g R_{t+1} (1 - d_t), and GAE."""
    rewards = np.array([[1.0], [0.0], [2.0]], np.float32)
    dones = np.array([[0], [1], [0]], np.uint8)
    values = np.array([[0.5], [0.25], [1.0]], np.float32)
    boot = np.array([4.0], np.float32)
    g = np.float32(0.5)
    got = orc.a2c_returns(boot, rewards, values, dones, 0.5)
    R2 = np.float32(2.0) + g * np.float32(4.0)          # 4.0
    R1 = np.float32(0.0)                                 # done at t=1 cuts the bootstrap
    R0 = np.float32(1.0) + g * R1
    assert_same(got.reshape(-1), np.array([R0, R1, R2], np.float32), 'n-step returns')
    lam = np.float32(0.5)
    d2 = np.float32(2.0) + g * np.float32(4.0) - np.float32(1.0); gae2 = d2
    d1 = np.float32(0.0) - np.float32(0.25); gae1 = d1   # done: no bootstrap, no carried advantage
    d0 = np.float32(1.0) + g * np.float32(0.25) - np.float32(0.5); gae0 = d0 + g * lam * gae1
    got = orc.a2c_returns(boot, rewards, values, dones, 0.5, 0.5)
    assert_same(got.reshape(-1), np.array([gae0 + np.float32(0.5), gae1 + np.float32(0.25), gae2 + np.float32(1.0)], np.float32),
                'GAE returns')
