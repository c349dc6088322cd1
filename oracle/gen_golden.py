"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the unmodified reference.

    python oracle/gen_golden.py            (container-only: needs the reference tree)

Each file holds short trajectories of the REFERENCE (run through oracle/reference_loader.py):
inputs (initial state, actions, the replay tape of the reference's own random draws) and every
output the reference produced (state after step, sanitised actions, reward, done, info, observation,
state and observation after reset).  The committed vectors are what pins the oracle on machines
where the reference tree is absent (tests/test_oracle_golden.py) and what the CUDA path is checked
against directly (tests/test_single_gpu.py).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import reference_loader as rl   # noqa: E402
from oracle import replay                   # noqa: E402

import torch                                # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def single_trajectory(ref, N, S, mode, steps, seed, reset_every=1, manual=None, actions=None):
    torch.manual_seed(seed)
    rl.take_tape()
    out = {'N': N, 'S': S, 'mode': mode, 'steps': steps}
    if manual is None:
        env = ref.SingleSnake(num_envs=N, size=S, observation_mode=mode)
        out['init_spawn'] = replay.single_reset_tape(rl.take_tape(), np.ones(N), N, S)
    else:
        env = ref.SingleSnake(num_envs=N, size=S, observation_mode=mode, manual_setup=True)
        env.envs = manual.clone()
    out['init_envs'] = env.envs.numpy().astype(np.int16)
    for t in range(steps):
        a = torch.randint(0, 4, (N,)) if actions is None else actions[t].clone()
        out[f'{t}/actions_in'] = a.numpy().copy()
        obs, reward, done, info = env.step(a)
        out[f'{t}/food_cell'] = replay.single_step_tape(rl.take_tape(), reward, N, S)
        out[f'{t}/envs'] = env.envs.numpy().astype(np.int16)
        assert np.array_equal(out[f'{t}/envs'].astype(np.float32), env.envs.numpy())
        out[f'{t}/actions_out'] = a.numpy().copy()
        out[f'{t}/reward'] = reward.numpy().reshape(-1)
        out[f'{t}/done'] = done.numpy().reshape(-1).astype(np.uint8)
        out[f'{t}/self_collision'] = info['self_collision'].numpy().astype(np.uint8)
        out[f'{t}/edge_collision'] = info['edge_collision'].numpy().astype(np.uint8)
        out[f'{t}/obs'] = obs.numpy()
        do_reset = (t % reset_every) == reset_every - 1
        out[f'{t}/did_reset'] = np.array(do_reset)
        if do_reset:
            obs2 = env.reset(done)
            out[f'{t}/spawn'] = replay.single_reset_tape(rl.take_tape(), done.numpy(), N, S)
            out[f'{t}/reset_envs'] = env.envs.numpy().astype(np.int16)
            out[f'{t}/reset_obs'] = obs2.numpy()
    return out


def multi_state(env):
    return {'foods': env.foods.numpy().astype(np.int16), 'heads': env.heads.numpy().astype(np.int16),
            'bodies': env.bodies.numpy().astype(np.int16), 'dones': env.dones.numpy().astype(np.uint8),
            'orientations': env.orientations.numpy().astype(np.int64),
            'boost_this_step': env.boost_this_step.numpy().astype(np.uint8),
            'agent_colours': env.agent_colours.numpy().astype(np.int16)}


def put(out, prefix, d):
    for k, v in d.items():
        if v is not None:
            out[f'{prefix}/{k}'] = np.asarray(v)


def multi_trajectory(ref, E, K, S, mode, steps, seed, **rules):
    torch.manual_seed(seed)
    rl.take_tape()
    env = rl.instrument_multi(ref.MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode=mode, **rules))
    out = {'E': E, 'K': K, 'S': S, 'mode': mode, 'steps': steps, 'rules': repr(sorted(rules.items()))}
    for k, v in rules.items():
        out[f'rule/{k}'] = v
    create, rest = replay.multi_create_tape(rl.take_tape(), np.arange(E), E, K, S)
    out['init/create'] = create
    put(out, 'init', multi_state(env))
    for t in range(steps):
        acts = torch.randint(0, 8, (E, K))
        out[f'{t}/actions'] = acts.numpy().copy()
        obs, rewards, dones, info = env.step({f'agent_{k}': acts[:, k].clone() for k in range(K)})
        draws = replay.multi_step_tape(rl.take_tape(), E, K, S)
        put(out, f'{t}/draws', {k: v for k, v in draws.items() if k != 'u_rate_dense'})   # dense = scatter(u_rate, selected)
        put(out, f'{t}/state', multi_state(env))
        out[f'{t}/obs'] = np.stack([obs[f'agent_{k}'].numpy() for k in range(K)])
        out[f'{t}/rewards'] = np.stack([rewards[f'agent_{k}'].numpy() for k in range(K)], axis=1)
        out[f'{t}/dones'] = np.stack([dones[f'agent_{k}'].numpy() for k in range(K)], axis=1).astype(np.uint8)
        out[f'{t}/all_done'] = dones['__all__'].numpy().astype(np.uint8)
        for name in ('snake_collision', 'edge_collision', 'food', 'boost', 'size'):
            a = np.stack([info[f'{name}_{k}'].numpy() for k in range(K)], axis=1)
            out[f'{t}/{name}'] = a.astype(np.uint8) if a.dtype == bool else a
        if t % 4 == 0:
            out[f'{t}/env_images'] = env._get_env_images().numpy()
        dones_before = env.dones.numpy().copy()
        env_done = dones['__all__'].numpy().copy()
        obs2 = env.reset(dones['__all__'])
        rdraws = replay.multi_reset_tape(rl.take_tape(), env_done, dones_before, env.agent_colours.numpy(),
                                         rules.get('respawn_mode', 'all') == 'any', E, K, S)
        put(out, f'{t}/reset_draws', rdraws)
        put(out, f'{t}/reset_state', multi_state(env))
        if t % 4 == 0:
            out[f'{t}/reset_obs'] = np.stack([obs2[f'agent_{k}'].numpy() for k in range(K)])
    return out


def grid_trajectory(ref, N, S, mode, steps, seed, manual=None, actions=None):
    torch.manual_seed(seed)
    rl.take_tape()
    start = (S // 2, S // 2)
    out = {'N': N, 'S': S, 'mode': mode, 'steps': steps, 'start': np.array(start)}
    if manual is None:
        env = ref.SimpleGridworld(num_envs=N, size=S, observation_mode=mode, start_location=start)
        out['init_food'] = replay.grid_food_tape(rl.take_tape(), np.ones(N), N, S)
    else:
        env = ref.SimpleGridworld(num_envs=N, size=S, observation_mode=mode, start_location=start, manual_setup=True)
        env.envs = manual.clone()
    out['init_envs'] = env.envs.numpy().astype(np.int16)
    for t in range(steps):
        a = torch.randint(0, 4, (N,)) if actions is None else actions[t].clone()
        out[f'{t}/actions'] = a.numpy().copy()
        obs, reward, done, info = env.step(a)
        out[f'{t}/food_cell'] = replay.grid_food_tape(rl.take_tape(), reward.numpy() != 0, N, S)
        out[f'{t}/envs'] = env.envs.numpy().astype(np.int16)
        out[f'{t}/reward'] = reward.numpy().reshape(-1)
        out[f'{t}/done'] = done.numpy().reshape(-1).astype(np.uint8)
        out[f'{t}/obs'] = obs.numpy()
        env.reset(done)
        out[f'{t}/reset_food'] = replay.grid_food_tape(rl.take_tape(), done.numpy(), N, S)
        out[f'{t}/reset_envs'] = env.envs.numpy().astype(np.int16)
    return out


def save(name, trajectories):
    flat = {'count': np.array(len(trajectories))}
    for i, tr in enumerate(trajectories):
        for k, v in tr.items():
            flat[f'{i}/{k}'] = np.asarray(v)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, **flat)
    print(path, os.path.getsize(path) // 1024, 'KiB')


def baseline_geometries(ref):
    """The BASELINE.json geometries themselves: C3 (size 36 `default`), the even sizes from 16
    up on which the CUDA side steps on body-only tiles, C5 (16 snakes, size 64) in both observation modes, and the
    32-snake maximum.  Kept in files of their own so that the round-1 fixtures stay byte-identical."""
    trs = [single_trajectory(ref, 10, 36, 'default', 14, seed=800),
           single_trajectory(ref, 8, 36, 'partial_2', 14, seed=801),
           single_trajectory(ref, 24, 16, 'partial_2', 20, seed=802),
           single_trajectory(ref, 12, 24, 'partial_2', 16, seed=803),
           single_trajectory(ref, 12, 24, 'one_channel', 16, seed=804),
           single_trajectory(ref, 12, 20, 'default', 21, seed=805, reset_every=7)]
    save('single_baseline.npz', trs)
    driver_rules = dict(food_mode='random_rate', food_rate=3e-4, respawn_mode='any', food_on_death_prob=0.33, boost_cost_prob=0.25)
    trs = [multi_trajectory(ref, 3, 16, 64, 'partial_4', 12, seed=900),
           multi_trajectory(ref, 2, 16, 64, 'full', 5, seed=901),
           multi_trajectory(ref, 2, 16, 64, 'partial_4', 10, seed=902, **driver_rules),
           multi_trajectory(ref, 2, 32, 48, 'partial_3', 8, seed=903),
           multi_trajectory(ref, 4, 4, 25, 'partial_4', 16, seed=904),
           multi_trajectory(ref, 4, 4, 25, 'partial_4', 16, seed=905, **driver_rules)]
    save('multi_baseline.npz', trs)


def main():
    ref = rl.load()
    if '--baseline-only' in sys.argv:
        baseline_geometries(ref)
        return
    trs = []
    for i, mode in enumerate(['partial_2', 'partial_3', 'default', 'raw', 'one_channel', 'positions']):
        trs.append(single_trajectory(ref, 24, 9, mode, 24, seed=100 + i))
        trs.append(single_trajectory(ref, 12, 13, mode, 16, seed=200 + i))
    # dead envs stepped again without reset: heads leave the grid, bodies decay to nothing
    for i, mode in enumerate(['default', 'one_channel', 'positions']):
        trs.append(single_trajectory(ref, 16, 9, mode, 21, seed=300 + i, reset_every=7))
    # the reference's own scenario fixtures (wurm/utils.py:68-110 get_test_env, size 12) under the
    # action sequences of tests/test_single_snake_env.py:55,89,122,146,174
    for orientation in ['up', 'right', 'down', 'left']:
        for seq in ([0, 0, 3, 0, 0, 1], [0, 3, 3, 0, 0], [1] * 10, [0, 3, 3, 2, 1, 0, 0, 0], [2, 2, 2, 3]):
            acts = torch.tensor(seq).unsqueeze(1).long()
            trs.append(single_trajectory(ref, 1, 12, 'default', len(seq), seed=400, reset_every=10 ** 6,
                                         manual=ref.utils.get_test_env(12, orientation), actions=acts))
    save('single.npz', trs)

    trs = [grid_trajectory(ref, 32, 7, 'default', 30, seed=700), grid_trajectory(ref, 16, 11, 'raw', 30, seed=701),
           grid_trajectory(ref, 1, 7, 'positions', 30, seed=702)]
    # the reference's own scenarios (tests/test_simple_gridworld.py:13-69), size 7
    for head, seq in [((3, 3), [0, 1, 2, 3, 2, 1]), ((2, 2), [0, 2, 2, 1]), ((3, 3), [0, 0, 0, 0])]:
        manual = torch.zeros((1, 2, 7, 7))
        manual[0, 0, 1, 1] = 1
        manual[0, 1, head[0], head[1]] = 1
        trs.append(grid_trajectory(ref, 1, 7, 'default', len(seq), seed=703, manual=manual,
                                   actions=torch.tensor(seq).unsqueeze(1).long()))
    save('gridworld.npz', trs)

    rule_sets = [
        dict(),                                                                          # constructor defaults
        dict(food_mode='random_rate', food_rate=3e-3, respawn_mode='any', food_on_death_prob=0.33, boost_cost_prob=0.25),
        dict(boost=False, food_on_death_prob=0.0, reward_on_death=-2),
        dict(respawn_mode='any', food_on_death_prob=1.0, boost_cost_prob=1.0, agent_colours='fixed'),
    ]
    trs = []
    for i, rules in enumerate(rule_sets):
        trs.append(multi_trajectory(ref, 6, 2, 12, 'full', 18, seed=500 + i, **rules))
        trs.append(multi_trajectory(ref, 4, 4, 14, 'partial_3', 18, seed=600 + i, **rules))
    save('multi.npz', trs)
    baseline_geometries(ref)


if __name__ == '__main__':
    main()
