"""GPU: MultiSnake through the C ABI (wurm_b200.envs.MultiSnake) against
  (1) the reference's golden vectors, replaying the reference's own random draws,
  (2) the CPU oracle on seeded random rollouts with Philox-derived draws (same seed on both sides),
  (3) the reference's own scenario tests (tests/test_multi_snake_env.py in the reference tree),
  (4) the reference's invariants (check_consistency) at BASELINE.json config-4 geometry.
Everything is compared bit for bit (fp32 as int32 patterns).
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from golden_util import (load, assert_same, multi_rules, multi_group, multi_step_draws, multi_state_arrays,
                         STATE_FIELDS)

pytestmark = pytest.mark.gpu

MULTI = load('multi.npz') + load('multi_baseline.npz')        # round-1 fixtures + the BASELINE.json geometries
DEV = 'cuda'


def np_(t):
    return t.detach().cpu().numpy()


def make_env(E, K, S, mode, **kw):
    from wurm_b200.envs import MultiSnake
    return MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode=mode, device=DEV, **kw)


def env_state(env):
    return dict(foods=np_(env.foods), heads=np_(env.heads), bodies=np_(env.bodies), dones=np_(env.dones).astype(np.uint8),
                orientations=np_(env.orientations), boost_this_step=np_(env.boost_this_step).astype(np.uint8),
                agent_colours=np_(env.agent_colours))


def check_state(env, expect, tag):
    got = env_state(env)
    for name in STATE_FIELDS:
        e = expect[name] if isinstance(expect, dict) else getattr(expect, name)
        assert_same(got[name], np.asarray(e).reshape(got[name].shape), f'{tag}: {name}')


def stack_dict(d, K, prefix='agent_'):
    return np.stack([np_(d[f'{prefix}{k}']) for k in range(K)], axis=1)


def check_step_outputs(env, K, obs, rewards, dones, info, expect, tag):
    """expect: dict with rewards/dones/all_done/snake_collision/edge_collision/food/boost/size (E,K) and obs (K,E,...)."""
    assert_same(stack_dict(rewards, K), expect['rewards'], tag + ': rewards')
    assert_same(stack_dict(dones, K).astype(np.uint8), expect['dones'], tag + ': dones')
    assert_same(np_(dones['__all__']).astype(np.uint8), expect['all_done'], tag + ': __all__')
    for name in ('snake_collision', 'edge_collision', 'boost'):
        assert_same(stack_dict(info, K, name + '_').astype(np.uint8), expect[name], f'{tag}: {name}')
    for name in ('food', 'size'):
        assert_same(stack_dict(info, K, name + '_'), expect[name], f'{tag}: {name}')
    assert_same(np.stack([np_(obs[f'agent_{k}']) for k in range(K)]), expect['obs'], tag + ': observations')


# state='dense': the reference's tensors, shadowed by the kernels' records (loaded instead of the tensors while nobody else wrote
# to them); 'dense_scan': the tensors alone, streamed by every call; 'compact': records resident in HBM, fp32 tensors
# materialised for the comparison
STATES = ['dense', 'dense_scan', 'compact']


@pytest.mark.parametrize('state', STATES)
@pytest.mark.parametrize('i', range(len(MULTI)))
def test_golden_replay(i, state):
    """CUDA path == reference on the recorded trajectories."""
    tr = MULTI[i]
    E, K, S, mode = int(tr['E']), int(tr['K']), int(tr['S']), str(tr['mode'])
    env = make_env(E, K, S, mode, manual_setup=True, state=state, **multi_rules(tr))
    init = multi_state_arrays(tr, 'init')
    env.agent_colours = torch.from_numpy(init['agent_colours']).to(DEV)
    env._create_all(draws=dict(create=tr['init/create'], respawn=np.full((E, 2), -1, np.int32), colours=init['agent_colours']))
    check_state(env, init, f'trajectory {i} creation')
    for t in range(int(tr['steps'])):
        tag = f'trajectory {i} ({mode}, K={K}, S={S}, {tr["rules"]}) step {t}'
        acts = torch.from_numpy(tr[f'{t}/actions']).to(DEV)
        obs, rewards, dones, info = env.step({f'agent_{k}': acts[:, k].contiguous() for k in range(K)},
                                             draws=multi_step_draws(tr, t, dense_rate=True))
        check_state(env, multi_state_arrays(tr, f'{t}/state'), tag)
        expect = {k: tr[f'{t}/{k}'] for k in ('rewards', 'dones', 'all_done', 'snake_collision', 'edge_collision', 'food',
                                              'boost', 'size', 'obs')}
        check_step_outputs(env, K, obs, rewards, dones, info, expect, tag)
        if f'{t}/env_images' in tr:
            assert_same(np_(env._get_env_images()), tr[f'{t}/env_images'], tag + ': env images')
        obs2 = env.reset(dones['__all__'], draws=multi_group(tr, f'{t}/reset_draws', ('create', 'respawn', 'colours')))
        check_state(env, multi_state_arrays(tr, f'{t}/reset_state'), tag + ' after reset')
        if f'{t}/reset_obs' in tr:
            assert_same(np.stack([np_(obs2[f'agent_{k}']) for k in range(K)]), tr[f'{t}/reset_obs'], tag + ': obs after reset')
    env.check_status()


DEFAULT_RULES = dict()
ANNEAL_RULES = dict(food_mode='random_rate', food_rate=3e-3, respawn_mode='any', food_on_death_prob=0.33, boost_cost_prob=0.25)
ROLLOUTS = [
    # E, K, S, mode, steps, rules, action dtype
    (200, 2, 12, 'full', 40, DEFAULT_RULES, torch.long),
    (150, 4, 25, 'partial_4', 60, DEFAULT_RULES, torch.long),          # BASELINE config 4 geometry
    (150, 4, 25, 'partial_5', 60, ANNEAL_RULES, torch.int),            # experiments/multiagent.py-style rules
    (64, 4, 12, 'partial_3', 60, dict(respawn_mode='any', food_on_death_prob=1.0, boost_cost_prob=1.0), torch.short),
    (64, 3, 13, 'full', 40, dict(boost=False, food_on_death_prob=0.0, reward_on_death=-2), torch.long),
    (12, 16, 64, 'partial_4', 40, DEFAULT_RULES, torch.long),          # BASELINE config 5 geometry
    (6, 16, 64, 'full', 12, ANNEAL_RULES, torch.long),
    (40, 10, 36, 'partial_2', 40, dict(respawn_mode='any'), torch.long),   # experiments/speeds.py geometry
    (3, 32, 40, 'partial_1', 20, DEFAULT_RULES, torch.long),           # the maximum number of snakes
    (1, 1, 7, 'full', 20, DEFAULT_RULES, torch.long),                  # the smallest env
    (20, 12, 20, 'partial_2', 30, dict(respawn_mode='any'), torch.long),   # one warp per env with 3K colours > 32 lanes
    (4, 6, 30, 'partial_3', 60, dict(food_on_death_prob=1.0), torch.long),  # long snakes: the load scan's hit queue overflows
]


@pytest.mark.parametrize('state', STATES)
@pytest.mark.parametrize('E,K,S,mode,steps,rules,adtype', ROLLOUTS)
def test_rollout_matches_oracle(E, K, S, mode, steps, rules, adtype, state):
    """Seeded random rollout (actions in [0,8): moves and boosts), Philox draws on both sides."""
    seed = 4321 + E + K + S
    env = make_env(E, K, S, mode, seed=seed, state=state, **rules)
    cfg = orc.multi_cfg(E, K, S, **rules)
    st = orc.MultiState(E, K, S)
    assert orc.multi_reset(cfg, st, np.ones(E, np.uint8), None, seed=seed, step=env._draws) == 0
    st.agent_colours[:] = np_(env.agent_colours)       # construction-time colours come from torch.rand
    check_state(env, st, 'creation')
    g = torch.Generator().manual_seed(seed)
    for t in range(steps):
        acts = torch.randint(0, 8, (E, K), generator=g)
        dev_acts = acts.to(device=DEV, dtype=adtype)
        obs, rewards, dones, info = env.step({f'agent_{k}': dev_acts[:, k].contiguous() for k in range(K)})
        out = orc.multi_step(cfg, st, acts.numpy(), None, seed=seed, step=env._draws)
        tag = f'step {t}'
        check_state(env, st, tag)
        o, bad = orc.multi_observe(cfg, st, mode)
        assert bad == 0
        expect = dict(rewards=out['rewards'], dones=st.dones.reshape(E, K), all_done=out['all_done'],
                      snake_collision=out['snake_collision'], edge_collision=out['edge_collision'], food=out['food'],
                      boost=st.boost_this_step.reshape(E, K), size=out['size'], obs=o)
        check_step_outputs(env, K, obs, rewards, dones, info, expect, tag)
        assert_same(np_(env._get_env_images()), orc.multi_env_images(cfg, st), tag + ': env images')
        obs2 = env.reset(dones['__all__'])
        orc.multi_reset(cfg, st, out['all_done'], None, seed=seed, step=env._draws)
        check_state(env, st, tag + ' after reset')
        o2, _ = orc.multi_observe(cfg, st, mode)
        assert_same(np.stack([np_(obs2[f'agent_{k}']) for k in range(K)]), o2, tag + ': obs after reset')
        if t % 10 == 9:
            env.check_consistency()
    env.check_status()


# ---- the reference's scenario tests on the CUDA path (reference tests/test_multi_snake_env.py) ----
size = 12


def get_test_env(num_envs=1, **kw):
    """The reference's two-snake fixture (:21-47)."""
    from wurm_b200.utils import determine_orientations
    kw.setdefault('seed', 20)                    # scenarios assert exact sizes: the food respawns must not depend on luck
    env = make_env(num_envs, 2, size, 'full', manual_setup=True, **kw)
    for i in range(num_envs):
        env.heads[2 * i, 0, 5, 5] = 1
        env.bodies[2 * i, 0, 5, 5] = 4
        env.bodies[2 * i, 0, 4, 5] = 3
        env.bodies[2 * i, 0, 4, 4] = 2
        env.bodies[2 * i, 0, 4, 3] = 1
        env.heads[2 * i + 1, 0, 8, 7] = 1
        env.bodies[2 * i + 1, 0, 8, 7] = 4
        env.bodies[2 * i + 1, 0, 8, 8] = 3
        env.bodies[2 * i + 1, 0, 8, 9] = 2
        env.bodies[2 * i + 1, 0, 9, 9] = 1
    _envs = torch.cat([env.foods.repeat_interleave(env.num_snakes, dim=0), env.heads, env.bodies], dim=1)
    env.orientations = determine_orientations(_envs)
    assert env.orientations.tolist() == [2, 3] * num_envs
    return env


def actions_at(all_actions, i):
    return {agent: torch.tensor([a[i]], device=DEV) for agent, a in all_actions.items()}


def head_of(env, agent):
    idx = env.heads[agent, 0].flatten().argmax().item()
    return [idx // size, idx % size]


def test_basic_movement():
    env = get_test_env()
    env.foods[0, 0, 1, 1] = 1
    all_actions = {'agent_0': [1, 2, 1, 1, 0, 3], 'agent_1': [0, 1, 3, 2, 1, 0]}
    expected = [[[5, 4], [4, 4], [4, 3], [4, 2], [5, 2], [5, 3]], [[9, 7], [9, 6], [9, 5], [8, 5], [8, 4], [9, 4]]]
    for i in range(6):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.check_consistency()
        for k in range(2):
            assert head_of(env, k) == expected[k][i]
        assert not any(d.item() for d in dones.values())


def test_edge_collision():
    env = get_test_env()
    env.food_on_death_prob = 1
    env.foods[0, 0, 1, 1] = 1
    all_actions = {'agent_0': [1, 1, 1, 1, 1], 'agent_1': [0, 2, 2, 6, 2]}
    for i in range(5):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.check_consistency()
        if i == 4:
            assert rewards['agent_0'].item() == env.reward_on_death
        if i == 2:
            assert rewards['agent_1'].item() == env.reward_on_death
        assert dones['agent_0'].item() == (i >= 4)
        assert dones['agent_1'].item() == (i >= 2)


def test_self_collision():
    env = get_test_env()
    env.food_on_death_prob = 1
    env.foods[0, 0, 4, 3] = 1
    all_actions = {'agent_0': [1, 2, 1, 1, 0, 3, 2, 0], 'agent_1': [0, 1, 3, 2, 1, 0, 0, 1]}
    for i in range(8):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.check_consistency()
        assert dones['agent_0'].item() == (i >= 6)


def test_other_snake_collision():
    env = get_test_env()
    env.foods[0, 0, 1, 1] = 1
    env.food_on_death_prob = 1
    all_actions = {'agent_0': [1, 2, 3, 3, 3, 3, 3, 2], 'agent_1': [1, 2, 2, 2, 2, 2, 2, 2]}
    for i in range(8):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.check_consistency()
        assert dones['agent_1'].item() == (i >= 4)
    assert env.foods[:, 0].sum().item() >= 2


def test_eat_food():
    env = get_test_env()
    env.foods[:, 0, 9, 7] = 1
    all_actions = {'agent_0': [1, 2, 1, 1, 0, 3], 'agent_1': [0, 1, 3, 2, 1, 0]}
    eaten = [0, 0]
    for i in range(6):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.check_consistency()
        assert not any(d.item() for d in dones.values())
        if i == 0:
            assert rewards['agent_1'].item() == 1
            assert env.foods[0, 0, 9, 7].item() == 0
        # (the respawned food lands on a random free cell and may be eaten again on a later step)
        eaten = [eaten[k] + int(rewards[f'agent_{k}'].item()) for k in range(2)]
    assert eaten[1] >= 1
    assert env.bodies.view(1, 2, -1).max(dim=2)[0].long().tolist() == [[4 + eaten[0], 4 + eaten[1]]]
    assert env.foods.sum().item() == 1


def test_create_envs():
    from wurm_b200.utils import determine_orientations
    env = make_env(512, 2, size, 'full')
    env.check_consistency()
    _envs = torch.cat([env.foods.repeat_interleave(env.num_snakes, dim=0), env.heads, env.bodies], dim=1)
    assert torch.equal(env.orientations, determine_orientations(_envs))


def test_reset():
    env = get_test_env()
    env.foods[:, 0, 1, 1] = 1
    all_actions = {'agent_0': [1, 2, 3, 3, 3, 3, 3, 3, 3], 'agent_1': [0, 1, 2, 2, 2, 2, 2, 2, 2]}
    for i in range(9):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.reset(dones['__all__'])
        env.check_consistency()
    assert torch.all(env.bodies.view(1, 2, -1).max(dim=-1)[0] == env.initial_snake_length)


def test_agent_observations():
    env = get_test_env()
    env.foods[:, 0, 1, 1] = 1
    obs_0, obs_1 = env._observe_agent(0), env._observe_agent(1)
    assert torch.allclose(obs_0[0, :, 4, 5] * 255, env.self_colour.float() / 2)
    assert torch.allclose(obs_0[0, :, 8, 8] * 255, env.other_colour.float() / 2)
    assert torch.allclose(obs_1[0, :, 4, 5] * 255, env.other_colour.float() / 2)
    assert torch.allclose(obs_1[0, :, 8, 8] * 255, env.self_colour.float() / 2)


def test_boost_through_food():
    env = get_test_env()
    env.boost = True
    env.foods[:, 0, 6, 5] = 1
    env.boost_cost_prob = 0
    all_actions = {'agent_0': [4, 1, 2], 'agent_1': [0, 1, 3]}
    for i in range(3):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.reset(dones['__all__'])
        env.check_consistency()
        if i == 0:
            assert rewards['agent_0'].item() == 1


def test_boost_cost_and_boost_leaves_food():
    env = get_test_env()
    env.boost = True
    env.boost_cost_prob = 1
    env.foods[:, 0, 1, 1] = 1
    all_actions = {'agent_0': [4, 1, 2], 'agent_1': [0, 1, 3]}
    for i in range(3):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.reset(dones['__all__'])
        env.check_consistency()
        assert env.bodies.view(1, 2, -1).max(dim=2)[0].long().tolist() == [[3, 4]]
        if i == 0:
            assert rewards['agent_0'].item() == -1
    assert env.foods[0, 0, 4, 4].item() == 1


def test_cant_boost_until_size_4():
    from wurm_b200.utils import determine_orientations
    env = make_env(1, 2, size, 'full', manual_setup=True, boost=True)
    env.foods[:, 0, 1, 1] = 1
    env.heads[0, 0, 5, 5] = 1
    env.bodies[0, 0, 5, 5] = 3
    env.bodies[0, 0, 4, 5] = 2
    env.bodies[0, 0, 4, 4] = 1
    env.heads[1, 0, 8, 7] = 1
    env.bodies[1, 0, 8, 7] = 3
    env.bodies[1, 0, 8, 8] = 2
    env.bodies[1, 0, 8, 9] = 1
    _envs = torch.cat([env.foods.repeat_interleave(env.num_snakes, dim=0), env.heads, env.bodies], dim=1)
    env.orientations = determine_orientations(_envs)
    expected = [[6, 5], [6, 4], [5, 4]]
    all_actions = {'agent_0': [4, 1, 2], 'agent_1': [0, 1, 3]}
    for i in range(3):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.reset(dones['__all__'])
        env.check_consistency()
        assert head_of(env, 0) == expected[i]


def test_respawn_mode_any_without_room():
    env = get_test_env()
    env.respawn_mode = 'any'
    for i in range(2, 9, 2):
        for j in range(2, 9, 2):
            env.foods[0, 0, i, j] = 1
    all_actions = {'agent_0': [1, 1, 1, 1, 2, 2, 2, 3], 'agent_1': [0, 1, 0, 0, 0, 0, 0, 1]}
    for i in range(8):
        observations, rewards, dones, info = env.step(actions_at(all_actions, i))
        env.reset(dones['__all__'])
        env.check_consistency()


def test_step_argument_errors():
    env = make_env(4, 2, size, 'full')
    good = torch.zeros(4, dtype=torch.long, device=DEV)
    with pytest.raises(RuntimeError):
        env.step({'agent_0': good})
    with pytest.raises(TypeError):
        env.step({'agent_0': good, 'agent_1': good.float()})
    with pytest.raises(RuntimeError):
        env.step({'agent_0': good, 'agent_1': torch.zeros(5, dtype=torch.long, device=DEV)})


def test_overlapping_input_is_reported():
    env = get_test_env()
    env.bodies[1, 0, 4, 4] = 7          # two bodies on one cell: outside the supported states
    env.step(actions_at({'agent_0': [1], 'agent_1': [0]}, 0))
    with pytest.raises(RuntimeError):
        env.check_status()


def test_config4_invariants_with_boost_and_respawn():
    """reference test_random_actions_with_boost (:94-124) at BASELINE config-4 geometry, scaled up."""
    E, K = 4096, 4
    env = make_env(E, K, 25, 'partial_4', respawn_mode='any', food_mode='random_rate', boost_cost_prob=0.25,
                   food_on_death_prob=0.33, food_rate=2.5e-4, seed=11)
    env.check_consistency()
    for t in range(60):
        actions = {f'agent_{k}': torch.randint(8, size=(E,), device=DEV) for k in range(K)}
        observations, reward, done, info = env.step(actions)
        assert observations['agent_0'].shape == (E, 3, 9, 9)
        env.reset(done['__all__'], return_observations=False)
        if t % 6 == 0:
            env.check_consistency()
    stats = env.stats()
    assert stats['env_steps'] == 60 * E and stats['reward'] > 0 and stats['edge_collisions'] > 0
    env.check_status()


def test_host_stepper_matches_direct_stepping():
    from wurm_b200 import HostStepper
    E, K, S, steps = 500, 4, 25, 10
    acts = torch.randint(0, 8, (steps, K, E), generator=torch.Generator().manual_seed(5))
    direct = make_env(E, K, S, 'partial_4', seed=21)
    piped = make_env(E, K, S, 'partial_4', seed=21)
    piped.agent_colours = direct.agent_colours.clone()
    stepper = HostStepper(piped, depth=2)
    expect = []
    for t in range(steps):
        obs, rewards, dones, info = direct.step({f'agent_{k}': acts[t, k].to(DEV) for k in range(K)})
        expect.append((stack_dict(rewards, K), stack_dict(dones, K), np_(dones['__all__'])))
        direct.reset(dones['__all__'], return_observations=False)
    for t in range(steps):
        ticket = stepper.submit({f'agent_{k}': acts[t, k].clone().pin_memory() for k in range(K)}).wait()
        assert_same(ticket.reward.numpy(), expect[t][0], f'step {t}: rewards')
        assert_same(ticket.done.numpy(), expect[t][1], f'step {t}: dones')
        assert_same(ticket.all_done.numpy(), expect[t][2], f'step {t}: __all__')
    check_state(piped, env_state(direct), 'final state')


def test_graphed_stepper_is_bit_identical_to_call_by_call_stepping():
    from wurm_b200 import GraphedStepper
    E, K, S, steps = 128, 4, 25, 25
    rules = dict(respawn_mode='any', food_mode='random_rate', food_rate=2e-3)
    plain = make_env(E, K, S, 'partial_4', seed=31, **rules)
    graphed = make_env(E, K, S, 'partial_4', seed=31, **rules)
    graphed.agent_colours = plain.agent_colours.clone()
    acts = torch.randint(0, 8, (steps + 1, K, E), generator=torch.Generator().manual_seed(2)).to(DEV)
    static = {f'agent_{k}': acts[0, k].clone() for k in range(K)}
    stepper = GraphedStepper(graphed, static, warmup=2)       # warms up with real steps, then restores the env
    check_state(graphed, env_state(plain), 'state after construction')
    assert graphed._draws == plain._draws
    for t in range(1, steps + 1):
        if t % 6 == 2:                                          # a direct call interleaved with the replays
            obs, rewards, dones, info = graphed.step({f'agent_{k}': acts[t, k].clone() for k in range(K)})
            graphed.reset(dones['__all__'], return_observations=False)
        else:
            for k in range(K):
                static[f'agent_{k}'].copy_(acts[t, k])
            obs, rewards, dones, info = stepper.step()
        obs2, rewards2, dones2, info2 = plain.step({f'agent_{k}': acts[t, k] for k in range(K)})
        for k in range(K):
            assert_same(np_(obs[f'agent_{k}']), np_(obs2[f'agent_{k}']), f'step {t}: obs {k}')
        assert_same(stack_dict(rewards, K), stack_dict(rewards2, K), f'step {t}: rewards')
        assert_same(stack_dict(dones, K), stack_dict(dones2, K), f'step {t}: dones')
        plain.reset(dones2['__all__'], return_observations=False)
        check_state(graphed, env_state(plain), f'step {t}: state after reset')


def test_only_one_food_on_a_nearly_full_board():
    """MultiSnake only_one food respawn with 0..5 free interior cells: rejection misses, ranking fallback, none free."""
    S, K = 9, 2
    cells = []
    for r in range(1, S - 1):
        cols = range(1, S - 1) if r % 2 == 1 else range(S - 2, 0, -1)
        cells += [(r, c) for c in cols]
    lengths = [40, 41, 42, 43, 44, 45] * 30                          # snake 0 covers path[0:L], snake 1 the last 3 cells
    E = len(lengths)
    env = make_env(E, K, S, 'partial_2', manual_setup=True, seed=77)
    st = orc.MultiState(E, K, S)
    acts = np.zeros((E, K), np.int64)
    delta = {(1, 0): 0, (0, -1): 1, (-1, 0): 2, (0, 1): 3}
    for e, L in enumerate(lengths):
        for v, (y, x) in enumerate(cells[:L], 1):
            st.bodies[e * K, 0, y, x] = v
        hy, hx = cells[L - 1]
        st.heads[e * K, 0, hy, hx] = 1
        ny, nx = cells[L]
        acts[e, 0] = delta[(ny - hy, nx - hx)]
        st.orientations[e * K] = (acts[e, 0] + 2) % 4                 # facing where it is about to go
        for v, (y, x) in enumerate(reversed(cells[46:49]), 1):       # snake 1: cells 48,47,46 hold 1,2,3; head on cell 46
            st.bodies[e * K + 1, 0, y, x] = v
        y, x = cells[46]
        st.heads[e * K + 1, 0, y, x] = 1
        acts[e, 1] = 7                                                # arbitrary: it will collide or move, both fine
    for name in ('foods', 'heads', 'bodies'):
        setattr(env, name, torch.from_numpy(getattr(st, name).copy()).to(DEV))
    env.orientations = torch.from_numpy(st.orientations.copy()).to(DEV)
    st.agent_colours[:] = np_(env.agent_colours)
    cfg = orc.multi_cfg(E, K, S)
    dev_acts = torch.from_numpy(acts).to(DEV)
    obs, rewards, dones, info = env.step({f'agent_{k}': dev_acts[:, k].contiguous() for k in range(K)})
    out = orc.multi_step(cfg, st, acts, None, seed=77, step=env._draws)
    check_state(env, st, 'state after a step on a nearly full board')
    assert_same(stack_dict(rewards, K), out['rewards'], 'rewards')
    env.check_status()


FUSED_CASES = [
    (300, 4, 25, 'partial_4', dict()),
    (200, 4, 25, 'partial_3', dict(food_mode='random_rate', food_rate=3e-3, respawn_mode='any', food_on_death_prob=0.33,
                                   boost_cost_prob=0.25)),
    (96, 3, 12, 'full', dict(respawn_mode='any', food_on_death_prob=1.0, agent_colours='fixed')),
    (24, 16, 64, 'partial_4', dict(respawn_mode='any')),
]


@pytest.mark.parametrize('state', STATES)
@pytest.mark.parametrize('E,K,S,mode,rules', FUSED_CASES)
def test_fused_step_reset_equals_step_then_reset(E, K, S, mode, rules, state):
    """step(a, auto_reset=True) == step(a); reset(dones['__all__'], return_observations=False): same outputs, same
    state afterwards (re-created envs, recoloured and respawned snakes), same draws."""
    two_calls = make_env(E, K, S, mode, seed=91, **rules)
    fused = make_env(E, K, S, mode, seed=91, state=state, **rules)
    fused.agent_colours = two_calls.agent_colours.clone()
    g = torch.Generator().manual_seed(3)
    for t in range(50):
        acts = torch.randint(0, 8, (K, E), generator=g).to(DEV)
        obs, rewards, dones, info = two_calls.step({f'agent_{k}': acts[k] for k in range(K)})
        two_calls.reset(dones['__all__'], return_observations=False)
        obs_f, rewards_f, dones_f, info_f = fused.step({f'agent_{k}': acts[k] for k in range(K)}, auto_reset=True)
        tag = f'step {t}: '
        for k in range(K):
            assert_same(np_(obs_f[f'agent_{k}']), np_(obs[f'agent_{k}']), tag + f'observation {k}')
        assert_same(stack_dict(rewards_f, K), stack_dict(rewards, K), tag + 'rewards')
        assert_same(stack_dict(dones_f, K), stack_dict(dones, K), tag + 'dones')
        assert_same(np_(dones_f['__all__']), np_(dones['__all__']), tag + '__all__')
        check_state(fused, env_state(two_calls), tag + 'state after reset')
    fused.check_consistency()
    assert fused._draws == two_calls._draws


@pytest.mark.parametrize('state', STATES)
@pytest.mark.parametrize('i', range(len(MULTI)))
def test_fused_step_reset_golden_replay(i, state):
    """The fused launch against the reference's recorded step+reset trajectories (both tapes replayed)."""
    tr = MULTI[i]
    E, K, S, mode = int(tr['E']), int(tr['K']), int(tr['S']), str(tr['mode'])
    env = make_env(E, K, S, mode, manual_setup=True, state=state, **multi_rules(tr))
    init = multi_state_arrays(tr, 'init')
    env.agent_colours = torch.from_numpy(init['agent_colours']).to(DEV)
    env._create_all(draws=dict(create=tr['init/create'], respawn=np.full((E, 2), -1, np.int32), colours=init['agent_colours']))
    for t in range(int(tr['steps'])):
        tag = f'trajectory {i} step {t}'
        acts = torch.from_numpy(tr[f'{t}/actions']).to(DEV)
        obs, rewards, dones, info = env.step({f'agent_{k}': acts[:, k].contiguous() for k in range(K)}, auto_reset=True,
                                             draws=multi_step_draws(tr, t, dense_rate=True),
                                             reset_draws=multi_group(tr, f'{t}/reset_draws', ('create', 'respawn', 'colours')))
        expect = {k: tr[f'{t}/{k}'] for k in ('rewards', 'dones', 'all_done', 'snake_collision', 'edge_collision', 'food',
                                              'boost', 'size', 'obs')}
        check_step_outputs(env, K, obs, rewards, dones, info, expect, tag)
        check_state(env, multi_state_arrays(tr, f'{t}/reset_state'), tag + ' after the fused reset')
    env.check_status()


@pytest.mark.parametrize('state', ['dense', 'dense_scan'])
@pytest.mark.parametrize('E,K,S,mode', [(96, 4, 25, 'partial_4'), (24, 16, 64, 'partial_4'), (48, 3, 12, 'full')])
def test_state_edited_between_calls_makes_head_hints_stale_not_wrong(E, K, S, mode, state):
    """The kernels keep each snake's head cell from one call to the next and skip streaming the heads tensor when every
    hint still verifies.  Rolling the envs along the batch (every env now sits under its neighbour's hints), replacing
    the tensors wholesale and killing snakes by hand must only cost the scan back, never the result.  The Python class
    drops the hints by itself when it sees such edits (tensor identity / version); `_adopt_state()` after every edit
    hides them from it -- as a raw-pointer writer would -- so that what is tested is the KERNEL's own verification.
    state='dense': the same for the shadow records (every cell they name must still hold that value in the tensors)."""
    seed = 99 + E
    rules = dict(respawn_mode='any')
    env = make_env(E, K, S, mode, seed=seed, state=state, **rules)
    cfg = orc.multi_cfg(E, K, S, **rules)
    st = orc.MultiState(E, K, S)
    assert orc.multi_reset(cfg, st, np.ones(E, np.uint8), None, seed=seed, step=env._draws) == 0
    st.agent_colours[:] = np_(env.agent_colours)
    g = torch.Generator().manual_seed(seed)

    def run(steps, tag):
        for t in range(steps):
            acts = torch.randint(0, 8, (E, K), generator=g)
            obs, rewards, dones, info = env.step({f'agent_{k}': acts[:, k].contiguous().to(DEV) for k in range(K)})
            out = orc.multi_step(cfg, st, acts.numpy(), None, seed=seed, step=env._draws)
            check_state(env, st, f'{tag} step {t}')
            o, _ = orc.multi_observe(cfg, st, mode)
            assert_same(np.stack([np_(obs[f'agent_{k}']) for k in range(K)]), o, f'{tag} step {t}: obs')
            assert_same(stack_dict(rewards, K), out['rewards'], f'{tag} step {t}: rewards')
            env.reset(dones['__all__'], return_observations=False)
            orc.multi_reset(cfg, st, out['all_done'], None, seed=seed, step=env._draws)
            check_state(env, st, f'{tag} step {t} after reset')

    run(8, 'warm')                                         # hints populated
    # 1. roll every tensor by one env: consistent states, foreign hints
    for name, per_env in (('foods', 1), ('heads', K), ('bodies', K), ('dones', K), ('orientations', K), ('boost_this_step', K),
                          ('agent_colours', K)):
        t = getattr(env, name)
        t.copy_(t.roll(per_env, dims=0))
        a = getattr(st, name)
        a[...] = np.roll(a, per_env, axis=0)
    env._adopt_state()
    run(4, 'rolled')
    # 2. the caller replaces the tensors by new objects holding another env's state (hints now describe the old tensors)
    env.heads = env.heads.view(E, K, 1, S, S).flip(0).reshape(E * K, 1, S, S).contiguous()      # env order reversed
    env.bodies = env.bodies.view(E, K, 1, S, S).flip(0).reshape(E * K, 1, S, S).contiguous()
    env.foods = env.foods.flip(0).contiguous()
    for name in ('dones', 'orientations', 'boost_this_step'):
        setattr(env, name, getattr(env, name).view(E, K).flip(0).reshape(E * K).contiguous())
    env.agent_colours = env.agent_colours.view(E, K, 3).flip(0).reshape(E * K, 3).contiguous()
    st.heads[...] = st.heads.reshape(E, K, 1, S, S)[::-1].reshape(E * K, 1, S, S)
    st.bodies[...] = st.bodies.reshape(E, K, 1, S, S)[::-1].reshape(E * K, 1, S, S)
    st.foods[...] = st.foods[::-1]
    for name in ('dones', 'orientations', 'boost_this_step'):
        a = getattr(st, name)
        a[...] = a.reshape(E, K)[::-1].reshape(E * K)
    st.agent_colours[...] = st.agent_colours.reshape(E, K, 3)[::-1].reshape(E * K, 3)
    env._adopt_state()
    run(4, 'replaced')
    # 3. kill one live snake per env by hand (tensors zeroed, flag set): its hint must not resurrect a head
    alive = (~env.dones.view(E, K)).float()
    victim = alive.argmax(dim=1)                           # first live snake of each env (all have one after a reset)
    rows = torch.arange(E, device=DEV) * K + victim
    has_live = alive.sum(dim=1) > 0
    rows = rows[has_live]
    env.heads[rows] = 0; env.bodies[rows] = 0; env.dones[rows] = True
    r = rows.cpu().numpy()
    st.heads[r] = 0; st.bodies[r] = 0; st.dones[r] = 1
    env._adopt_state()
    run(4, 'killed')


def test_second_head_written_by_the_caller_is_reported_not_ignored():
    """A second head cell of a snake is outside the supported set (the reference's check_consistency rejects it too).
    In steady state the kernels run on verified head hints and do not stream the heads tensor, so they would never
    SEE it: a write through `env.heads` must drop the hints so that the next call scans the tensor and raises
    WURM_ST_MULTI_HEAD instead of silently stepping a different state."""
    E, K, S = 32, 4, 25
    env = make_env(E, K, S, 'partial_4', seed=5)
    g = torch.Generator().manual_seed(5)
    for t in range(3):
        acts = torch.randint(0, 4, (E, K), generator=g)
        env.step({f'agent_{k}': acts[:, k].contiguous().to(DEV) for k in range(K)}, auto_reset=True)
    env.check_status()
    assert int((env._head_hints >= 0).sum()) > 0
    live = int((~env.dones).nonzero()[0])
    free = (env.heads[live, 0] == 0).nonzero()[0]
    env.heads[live, 0, free[0], free[1]] = 1.0
    acts = torch.randint(0, 4, (E, K), generator=g)
    env.step({f'agent_{k}': acts[:, k].contiguous().to(DEV) for k in range(K)})
    with pytest.raises(RuntimeError, match='more than one head'):
        env.check_status()
    # by hand, for raw-pointer writers
    env.invalidate_hints()
    assert int((env._head_hints != -2).sum()) == 0


def test_compact_state_takes_caller_edits_and_refuses_what_it_cannot_carry():
    """state='compact': the fp32 tensors are materialised on access; in-place writes and replaced tensors are folded back
    into the records before the next call (the reference's tests build their fixtures exactly like this), and a state the
    records cannot carry exactly raises instead of being rounded."""
    E, K, S = 8, 2, 12
    twin = make_env(E, K, S, 'full', manual_setup=True, seed=3)
    env = make_env(E, K, S, 'full', manual_setup=True, seed=3, state='compact')
    env.agent_colours = twin.agent_colours.clone()
    for e_ in (twin, env):
        for i in range(E):
            e_.heads[2 * i, 0, 5, 5] = 1
            e_.bodies[2 * i, 0, 5, 5] = 4; e_.bodies[2 * i, 0, 4, 5] = 3; e_.bodies[2 * i, 0, 4, 4] = 2; e_.bodies[2 * i, 0, 4, 3] = 1
            e_.heads[2 * i + 1, 0, 8, 7] = 1
            e_.bodies[2 * i + 1, 0, 8, 7] = 4; e_.bodies[2 * i + 1, 0, 8, 8] = 3; e_.bodies[2 * i + 1, 0, 8, 9] = 2; e_.bodies[2 * i + 1, 0, 9, 9] = 1
        e_.foods[:, 0, 1, 1] = 1
        e_.orientations = torch.tensor([1, 3] * E, device=DEV)
    g = torch.Generator().manual_seed(1)
    for t in range(12):
        acts = torch.randint(0, 8, (K, E), generator=g).to(DEV)
        o1, r1, d1, _ = twin.step({f'agent_{k}': acts[k] for k in range(K)}, auto_reset=True)
        o2, r2, d2, _ = env.step({f'agent_{k}': acts[k] for k in range(K)}, auto_reset=True)
        for k in range(K):
            assert_same(np_(o2[f'agent_{k}']), np_(o1[f'agent_{k}']), f'step {t}: obs {k}')
        assert_same(stack_dict(r2, K), stack_dict(r1, K), f'step {t}: rewards')
        check_state(env, env_state(twin), f'step {t}')
        if t == 5:                                       # replace a tensor wholesale, and write into another
            new_foods = twin.foods.clone(); new_foods[:, 0, 6, 6] = 1
            twin.foods = new_foods.clone(); env.foods = new_foods
    env.check_consistency()
    assert env._dense is not None                        # materialised by the checks above ...
    acts = torch.zeros((K, E), dtype=torch.long, device=DEV)
    env.step({f'agent_{k}': acts[k] for k in range(K)})
    assert env._dense is None                            # ... and dropped by the state-changing call
    env.bodies[0, 0, 2, 2] = 0.5                         # a non-integral body value: not a record
    with pytest.raises(RuntimeError, match='compact'):
        env.step({f'agent_{k}': acts[k] for k in range(K)})


@pytest.mark.parametrize('E,K,S,mode', [(24, 4, 25, 'partial_4'), (6, 8, 40, 'partial_2'), (4, 16, 64, 'partial_4')])
def test_compact_state_with_more_live_cells_than_the_live_list_holds(E, K, S, mode):
    """state='compact' keeps its live list only as large as the fused reset's scratch (shared memory per CTA is what
    limits residency); a board flooded with food overflows it and the per-cell passes walk the whole grid instead.
    Results must not change: compared with the dense twin, whose list can hold every cell."""
    rules = dict(food_mode='random_rate', food_rate=0.97, respawn_mode='any', food_on_death_prob=1.0)
    dense = make_env(E, K, S, mode, seed=7, **rules)
    compact = make_env(E, K, S, mode, seed=7, state='compact', **rules)
    compact.agent_colours = dense.agent_colours.clone()
    g = torch.Generator().manual_seed(4)
    for t in range(16):
        acts = torch.randint(0, 8, (K, E), generator=g).to(DEV)
        o1, r1, d1, _ = dense.step({f'agent_{k}': acts[k] for k in range(K)}, auto_reset=True)
        o2, r2, d2, _ = compact.step({f'agent_{k}': acts[k] for k in range(K)}, auto_reset=True)
        for k in range(K):
            assert_same(np_(o2[f'agent_{k}']), np_(o1[f'agent_{k}']), f'step {t}: obs {k}')
        assert_same(stack_dict(r2, K), stack_dict(r1, K), f'step {t}: rewards')
        check_state(compact, env_state(dense), f'step {t}')
        if t == 0:
            assert float(dense.foods.sum()) > 0.6 * E * (S - 2) ** 2     # the flood happened
        if t == 8:                                      # back to sparse boards: the list is used again
            dense.food_rate = compact.food_rate = 1e-3
            dense.foods = torch.zeros_like(dense.foods); compact.foods = torch.zeros_like(dense.foods)
    compact.check_consistency()


def test_graphed_stepper_on_the_compact_state():
    from wurm_b200 import GraphedStepper
    E, K, S, steps = 64, 4, 25, 15
    dense = make_env(E, K, S, 'partial_4', seed=31, respawn_mode='any')
    compact = make_env(E, K, S, 'partial_4', seed=31, respawn_mode='any', state='compact')
    compact.agent_colours = dense.agent_colours.clone()
    acts = torch.randint(0, 8, (steps, K, E), generator=torch.Generator().manual_seed(2)).to(DEV)
    static = {f'agent_{k}': acts[0, k].clone() for k in range(K)}
    stepper = GraphedStepper(compact, static, warmup=2)
    check_state(compact, env_state(dense), 'state after construction')        # (materialises the tensors ...)
    for t in range(steps):
        for k in range(K):
            static[f'agent_{k}'].copy_(acts[t, k])
        obs, rewards, dones, info = stepper.step()                          # (... which the replay must drop again)
        obs2, rewards2, dones2, info2 = dense.step({f'agent_{k}': acts[t, k] for k in range(K)}, auto_reset=True)
        for k in range(K):
            assert_same(np_(obs[f'agent_{k}']), np_(obs2[f'agent_{k}']), f'step {t}: obs {k}')
        assert_same(stack_dict(rewards, K), stack_dict(rewards2, K), f'step {t}: rewards')
        if t % 5 == 0:
            check_state(compact, env_state(dense), f'step {t}')
    check_state(compact, env_state(dense), 'final state')


SHADOW_CASES = [
    (96, 4, 25, 'partial_4', dict(respawn_mode='any')),
    (64, 3, 12, 'full', dict(food_mode='random_rate', food_rate=5e-3, food_on_death_prob=1.0)),
    (16, 16, 64, 'partial_4', dict(respawn_mode='any', food_on_death_prob=0.33)),
    (8, 32, 40, 'partial_2', dict()),
]


@pytest.mark.parametrize('E,K,S,mode,rules', SHADOW_CASES)
def test_shadow_records_track_the_tensors(E, K, S, mode, rules):
    """state='dense' (the default): the reference's tensors are the state and the library keeps its records beside them.
    Stepped next to a 'dense_scan' env (tensors only) with the same seed and actions -- fused and two-call steps, observes and
    consistency checks in between -- every output and every state tensor must be identical after every call, and the records
    must always equal what a fresh conversion of the tensors gives."""
    shadow = make_env(E, K, S, mode, seed=17, state='dense', **rules)
    scan = make_env(E, K, S, mode, seed=17, state='dense_scan', **rules)
    scan.agent_colours = shadow.agent_colours.clone()
    probe = make_env(E, K, S, mode, seed=1, state='compact', manual_setup=True)
    g = torch.Generator().manual_seed(8)
    assert shadow._cells is not None and scan._cells is None
    for t in range(40):
        acts = torch.randint(0, 8, (K, E), generator=g).to(DEV)
        fused = t % 3 == 1
        outs = []
        for env in (shadow, scan):
            obs, rewards, dones, info = env.step({f'agent_{k}': acts[k] for k in range(K)}, auto_reset=fused)
            if not fused:
                env.reset(dones['__all__'], return_observations=False)
            outs.append((obs, rewards, dones, info))
        tag = f'step {t}: '
        assert shadow._shadow_ok
        for k in range(K):
            assert_same(np_(outs[0][0][f'agent_{k}']), np_(outs[1][0][f'agent_{k}']), tag + f'obs {k}')
        for i, name in ((1, 'rewards'), (2, 'dones')):
            assert_same(stack_dict(outs[0][i], K), stack_dict(outs[1][i], K), tag + name)
        assert_same(np_(outs[0][2]['__all__']), np_(outs[1][2]['__all__']), tag + '__all__')
        for name in ('snake_collision', 'edge_collision', 'boost', 'food', 'size'):
            assert_same(stack_dict(outs[0][3], K, name + '_'), stack_dict(outs[1][3], K, name + '_'), tag + name)
        check_state(shadow, env_state(scan), tag + 'state')
        if t % 4 == 3:
            o1, o2 = shadow._observe('full'), scan._observe('full')
            for k in range(K):
                assert_same(np_(o1[f'agent_{k}']), np_(o2[f'agent_{k}']), tag + f'full observation {k}')
            assert_same(np_(shadow._get_env_images()), np_(scan._get_env_images()), tag + 'env images')
            shadow.check_consistency()
            # the records against a fresh conversion of the tensors
            probe.foods, probe.heads, probe.bodies = shadow.foods.clone(), shadow.heads.clone(), shadow.bodies.clone()
            probe.dones = shadow.dones.clone()
            probe._state()
            assert_same(np_(shadow._cells), np_(probe._cells), tag + 'records')
            assert_same(np_(shadow._head_hints), np_(probe._head_hints), tag + 'head cells')
    shadow.check_status()
    scan.check_status()


def test_shadow_records_are_dropped_when_the_caller_writes_to_the_tensors():
    """A write through torch (in place, or a replaced tensor) is seen by the version counters: the records are not trusted
    for the next call, which streams the tensors and re-emits them.  Adding a food cell and a whole extra snake body
    segment -- things the records' presence check alone could not notice -- must change the trajectory exactly as it does
    for the tensors-only env and the oracle."""
    E, K, S, mode = 48, 4, 25, 'partial_4'
    rules = dict(food_mode='random_rate', food_rate=1e-3)
    env = make_env(E, K, S, mode, seed=23, **rules)
    cfg = orc.multi_cfg(E, K, S, **rules)
    st = orc.MultiState(E, K, S)
    assert orc.multi_reset(cfg, st, np.ones(E, np.uint8), None, seed=23, step=env._draws) == 0
    st.agent_colours[:] = np_(env.agent_colours)
    g = torch.Generator().manual_seed(4)

    def run(steps, tag):
        for t in range(steps):
            acts = torch.randint(0, 8, (E, K), generator=g)
            obs, rewards, dones, info = env.step({f'agent_{k}': acts[:, k].contiguous().to(DEV) for k in range(K)})
            out = orc.multi_step(cfg, st, acts.numpy(), None, seed=23, step=env._draws)
            check_state(env, st, f'{tag} step {t}')
            o, _ = orc.multi_observe(cfg, st, mode)
            assert_same(np.stack([np_(obs[f'agent_{k}']) for k in range(K)]), o, f'{tag} step {t}: obs')
            assert_same(stack_dict(rewards, K), out['rewards'], f'{tag} step {t}: rewards')
            env.reset(dones['__all__'], return_observations=False)
            orc.multi_reset(cfg, st, out['all_done'], None, seed=23, step=env._draws)

    run(5, 'warm')
    assert env._shadow_ok
    # food in front of every living snake's head (in place: version bump)
    heads = env.heads.view(E, K, S, S)
    for e in range(E):
        for k in range(K):
            if not bool(env.dones[e * K + k]):
                y, x = divmod(int(heads[e, k].flatten().argmax()), S)
                for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
                    if 0 < yy < S - 1 and 0 < xx < S - 1 and float(env.bodies.view(E, K, S, S)[e, :, yy, xx].sum()) == 0:
                        env.foods[e, 0, yy, xx] = 1.0
                        st.foods[e, 0, yy, xx] = 1.0
    run(5, 'food added')
    assert env._shadow_ok
    # replaced tensors (new objects, same content but one more food cell in env 0)
    f = env.foods.clone()
    free = ((env.bodies.view(E, K, S, S)[0].sum(0) + f[0, 0]) == 0).nonzero()
    y, x = [int(v) for v in free[(free[:, 0] > 0) & (free[:, 0] < S - 1) & (free[:, 1] > 0) & (free[:, 1] < S - 1)][0]]
    f[0, 0, y, x] = 1.0
    st.foods[0, 0, y, x] = 1.0
    env.foods = f
    run(5, 'foods replaced')
    env.check_status()


def test_graphed_stepper_on_the_shadowed_state_sees_caller_edits():
    """The captured launch is the record-loading one; an edit of the tensors between replays is noticed on the host (version
    counters) and handed to the captured kernel as 'unknown' head hints, which makes it re-load those envs from the tensors."""
    from wurm_b200 import GraphedStepper
    E, K, S, steps = 64, 4, 25, 18
    plain = make_env(E, K, S, 'partial_4', seed=31, state='dense_scan', respawn_mode='any')
    graphed = make_env(E, K, S, 'partial_4', seed=31, state='dense', respawn_mode='any')
    graphed.agent_colours = plain.agent_colours.clone()
    acts = torch.randint(0, 8, (steps + 1, K, E), generator=torch.Generator().manual_seed(2)).to(DEV)
    static = {f'agent_{k}': acts[0, k].clone() for k in range(K)}
    stepper = GraphedStepper(graphed, static, warmup=2)
    assert graphed._shadow_ok                                   # derived at construction: the capture loads records
    check_state(graphed, env_state(plain), 'state after construction')
    for t in range(1, steps + 1):
        if t % 5 == 0:                                          # the caller drops food onto a free cell of every env
            for env in (graphed, plain):
                occ = env.bodies.view(E, K, S, S).sum(1) + env.foods.view(E, S, S)
                occ[:, 0, :] = 1; occ[:, -1, :] = 1; occ[:, :, 0] = 1; occ[:, :, -1] = 1
                cell = (occ == 0).view(E, -1).float().argmax(dim=1)
                env.foods.view(E, -1)[torch.arange(E, device=DEV), cell] = 1.0
        for k in range(K):
            static[f'agent_{k}'].copy_(acts[t, k])
        obs, rewards, dones, info = stepper.step()
        obs2, rewards2, dones2, info2 = plain.step({f'agent_{k}': acts[t, k] for k in range(K)}, auto_reset=True)
        for k in range(K):
            assert_same(np_(obs[f'agent_{k}']), np_(obs2[f'agent_{k}']), f'step {t}: obs {k}')
        assert_same(stack_dict(rewards, K), stack_dict(rewards2, K), f'step {t}: rewards')
        check_state(graphed, env_state(plain), f'step {t}: state')
    graphed.check_status()
