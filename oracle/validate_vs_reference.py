"""TEST INFRASTRUCTURE ONLY -- pins the C oracle against the unmodified reference (container-only).

    python oracle/validate_vs_reference.py [--quick]

Runs the reference (through oracle/reference_loader.py) and the oracle side by side on the same
seeded states, actions and replayed draws, comparing every state tensor, reward, done flag, info
flag, sanitised action and observation bit for bit (int32 views of the fp32 data).  The summary it
prints is recorded in DESIGN.md.
"""
import os
import sys
import argparse

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import reference_loader as rl   # noqa: E402
from oracle import oracle as orc            # noqa: E402
from oracle import replay                   # noqa: E402

import torch                                # noqa: E402


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int32) if a.dtype == np.float32 else a


def same(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


class Tally(object):
    def __init__(self):
        self.checks = 0
        self.fails = []

    def check(self, name, a, b):
        self.checks += 1
        if not same(a, b):
            self.fails.append(name)
            if len(self.fails) < 5:
                print('MISMATCH', name, np.asarray(a).shape, np.asarray(b).shape)


def validate_single(ref, tally, N, S, mode, steps, seed, reset_every_step=True):
    """step(a); reset(done) loop, every output compared.  With reset_every_step=False dead envs keep
    being stepped (heads leave the grid, bodies decay to nothing) -- the degenerate states."""
    torch.manual_seed(seed)
    rl.take_tape()
    env = ref.SingleSnake(num_envs=N, size=S, observation_mode=mode)
    spawn = replay.single_reset_tape(rl.take_tape(), np.ones(N), N, S)
    state = np.zeros((N, 3, S, S), np.float32)
    orc.single_reset(state, np.ones(N, np.uint8), spawn)
    tally.check(f'single/{mode}/create', env.envs.numpy(), state)
    env_steps = 0
    for t in range(steps):
        a_ref = torch.randint(0, 4, (N,))
        a_orc = a_ref.numpy().copy()
        obs, reward, done, info = env.step(a_ref)
        food_cell = replay.single_step_tape(rl.take_tape(), reward, N, S)
        r, d, sc, ec = orc.single_step(state, a_orc, food_cell)
        o, bad = orc.single_observe(state, mode)
        tag = f'single/{mode}/S{S}/t{t}'
        tally.check(tag + '/envs', env.envs.numpy(), state)
        tally.check(tag + '/actions', a_ref.numpy(), a_orc)
        tally.check(tag + '/reward', reward.numpy().reshape(-1), r)
        tally.check(tag + '/done', done.numpy().reshape(-1).astype(np.uint8), d)
        tally.check(tag + '/self_collision', info['self_collision'].numpy().astype(np.uint8), sc)
        tally.check(tag + '/edge_collision', info['edge_collision'].numpy().astype(np.uint8), ec)
        tally.check(tag + '/obs', obs.numpy(), o)
        assert bad == 0
        env_steps += N
        if reset_every_step or t % 7 == 6:
            obs2 = env.reset(done)
            spawn = replay.single_reset_tape(rl.take_tape(), done.numpy(), N, S)
            orc.single_reset(state, done.numpy().reshape(-1).astype(np.uint8), spawn)
            tally.check(tag + '/reset_envs', env.envs.numpy(), state)
            o2, bad = orc.single_observe(state, mode)
            tally.check(tag + '/reset_obs', obs2.numpy(), o2)
    return env_steps


import collections
EVENTS = collections.Counter()


def ref_multi_state(env):
    """The reference env's state as the dict of arrays the oracle's MultiState holds."""
    return dict(foods=env.foods.numpy(), heads=env.heads.numpy(), bodies=env.bodies.numpy(),
                dones=env.dones.numpy().astype(np.uint8), orientations=env.orientations.numpy(),
                boost_this_step=env.boost_this_step.numpy().astype(np.uint8), agent_colours=env.agent_colours.numpy())


def check_multi_state(tally, tag, env, st, colours=True):
    ref = ref_multi_state(env)
    for name in ('foods', 'heads', 'bodies', 'dones', 'orientations', 'boost_this_step') + (('agent_colours',) if colours else ()):
        tally.check(f'{tag}/{name}', ref[name], getattr(st, name))


def validate_multi(ref, tally, E, K, S, mode, steps, seed, **rules):
    """step(actions); reset(done['__all__']) loop for one rule set (constructor kwargs in `rules`)."""
    torch.manual_seed(seed)
    rl.take_tape()
    env = rl.instrument_multi(ref.MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode=mode, **rules))
    cfg = orc.multi_cfg(E, K, S, **{k: v for k, v in rules.items() if k != 'agent_colours'},
                        colour_mode=rules.get('agent_colours', 'random'))
    st = orc.MultiState(E, K, S)
    tape = rl.take_tape()
    create, rest = replay.multi_create_tape(tape, np.arange(E), E, K, S)
    assert len(rest) == 1 and rest[0][0] == 'rand'          # the colours (:145/148)
    st.dones[:] = 1
    st.agent_colours[:] = env.agent_colours.numpy()
    failed = orc.multi_reset(cfg, st, np.ones(E, np.uint8), dict(create=create, respawn=np.full((E, 2), -1, np.int32),
                                                                  colours=env.agent_colours.numpy()))
    assert failed == 0
    tag = f'multi/{mode}/K{K}S{S}/{sorted(rules.items())}'
    check_multi_state(tally, tag + '/create', env, st)
    env_steps = 0
    respawn_any = rules.get('respawn_mode', 'all') == 'any'
    for t in range(steps):
        acts = torch.randint(0, 8, (E, K))
        actions = {f'agent_{k}': acts[:, k].clone() for k in range(K)}
        obs, rewards, dones, info = env.step(actions)
        draws = replay.multi_step_tape(rl.take_tape(), E, K, S)
        out = orc.multi_step(cfg, st, acts.numpy(), draws)
        o, bad = orc.multi_observe(cfg, st, mode)
        assert bad == 0
        tg = f'{tag}/t{t}'
        check_multi_state(tally, tg, env, st)
        if draws['u_rate'] is not None:
            tally.check(tg + '/rate_selected', out['rate_selected'], draws['selected'])
        for k in range(K):
            tally.check(f'{tg}/obs_{k}', obs[f'agent_{k}'].numpy(), o[k])
            tally.check(f'{tg}/reward_{k}', rewards[f'agent_{k}'].numpy(), out['rewards'][:, k])
            tally.check(f'{tg}/done_{k}', dones[f'agent_{k}'].numpy().astype(np.uint8), st.dones.reshape(E, K)[:, k])
            tally.check(f'{tg}/snake_collision_{k}', info[f'snake_collision_{k}'].numpy().astype(np.uint8), out['snake_collision'][:, k])
            tally.check(f'{tg}/edge_collision_{k}', info[f'edge_collision_{k}'].numpy().astype(np.uint8), out['edge_collision'][:, k])
            tally.check(f'{tg}/food_{k}', info[f'food_{k}'].numpy(), out['food'][:, k])
            tally.check(f'{tg}/boost_{k}', info[f'boost_{k}'].numpy().astype(np.uint8), st.boost_this_step.reshape(E, K)[:, k])
            tally.check(f'{tg}/size_{k}', info[f'size_{k}'].numpy(), out['size'][:, k])
        tally.check(tg + '/all_done', dones['__all__'].numpy().astype(np.uint8), out['all_done'])
        tally.check(tg + '/env_images', env._get_env_images().numpy(), orc.multi_env_images(cfg, st))
        env_steps += E
        EVENTS['boost_phases'] += int(draws['boost_phase_ran'])
        EVENTS['boosting_agents'] += int(st.boost_this_step.sum())
        EVENTS['snake_collisions'] += int(out['snake_collision'].sum())
        EVENTS['edge_collisions'] += int(out['edge_collision'].sum())
        EVENTS['food_eaten'] += int(out['food'].sum())
        EVENTS['env_recreations'] += int(out['all_done'].sum())
        dones_before = env.dones.numpy().copy()
        env_done = dones['__all__'].numpy().copy()
        obs2 = env.reset(dones['__all__'])
        rdraws = replay.multi_reset_tape(rl.take_tape(), env_done, dones_before, env.agent_colours.numpy(), respawn_any, E, K, S)
        orc.multi_reset(cfg, st, env_done, rdraws)
        EVENTS['respawns'] += int((rdraws['respawn'][:, 0] >= 0).sum())
        EVENTS['failed_respawns'] += int(((rdraws['respawn'][:, 0] < 0) & (rdraws['respawn'][:, 1] >= 0)).sum())
        check_multi_state(tally, tg + '/reset', env, st)
        o2, bad = orc.multi_observe(cfg, st, mode)
        for k in range(K):
            tally.check(f'{tg}/reset_obs_{k}', obs2[f'agent_{k}'].numpy(), o2[k])
    return env_steps


def validate_grid(ref, tally, N, S, mode, steps, seed):
    """SimpleGridworld: step(a); reset(done) loop, every output compared."""
    torch.manual_seed(seed)
    rl.take_tape()
    start = (S // 2, S // 2)
    env = ref.SimpleGridworld(num_envs=N, size=S, observation_mode=mode, start_location=start)
    state = np.zeros((N, 2, S, S), np.float32)
    orc.grid_reset(state, np.ones(N, np.uint8), start, replay.grid_food_tape(rl.take_tape(), np.ones(N), N, S))
    tally.check(f'grid/{mode}/create', env.envs.numpy(), state)
    for t in range(steps):
        a = torch.randint(0, 4, (N,))
        obs, reward, done, info = env.step(a)
        r, d = orc.grid_step(state, a.numpy(), replay.grid_food_tape(rl.take_tape(), reward.numpy() != 0, N, S))
        tag = f'grid/{mode}/S{S}/t{t}'
        tally.check(tag + '/envs', env.envs.numpy(), state)
        tally.check(tag + '/reward', reward.numpy().reshape(-1), r)
        tally.check(tag + '/done', done.numpy().reshape(-1).astype(np.uint8), d)
        tally.check(tag + '/edge_collision', info['edge_collision'].numpy().astype(np.uint8), d)
        if mode != 'positions' or N == 1:
            tally.check(tag + '/obs', obs.numpy(), orc.grid_observe(state, mode))
        env.reset(done)
        orc.grid_reset(state, done.numpy().reshape(-1).astype(np.uint8), start,
                       replay.grid_food_tape(rl.take_tape(), done.numpy(), N, S))
        tally.check(tag + '/reset_envs', env.envs.numpy(), state)
    return N * steps


def validate_a2c(ref_path, tally):
    """A2C.loss of the reference (wurm/rl/a2c.py:32-79) against losses computed from the oracle's returns with
    the reference's own two closing lines (:70-73).  The reference's `return_returns` is broken (`tuple += Tensor`),
    so the returns are pinned through the two losses, compared bit for bit."""
    sys.path.insert(0, ref_path)
    from wurm.rl.a2c import A2C
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    n = 0
    for T, N in [(20, 512), (5, 33), (64, 1000)]:
        # (0.99, 0.92) and (0.9, 0.9): pairs whose fp32 product differs from the fp32 rounding of the double product
        # the reference forms at a2c.py:56
        for gamma, gae_lambda in ((0.99, None), (0.99, 0.95), (0.99, 0.92), (0.9, 0.9)):
            rewards = (torch.rand(T, N, 1, generator=g) < 0.1).float() - (torch.rand(T, N, 1, generator=g) < 0.05).float()
            values = torch.randn(T, N, 1, generator=g)
            log_probs = -torch.rand(T, N, 1, generator=g)
            dones = torch.rand(T, N, 1, generator=g) < 0.08
            bootstrap = torch.randn(N, 1, generator=g)
            a2c = A2C(gamma=gamma, use_gae=gae_lambda is not None, gae_lambda=gae_lambda)
            value_loss, policy_loss = a2c.loss(bootstrap, rewards, values, log_probs, dones)
            ret = torch.from_numpy(orc.a2c_returns(bootstrap.numpy().reshape(N), rewards.numpy().reshape(T, N),
                                                   values.numpy().reshape(T, N), dones.numpy().reshape(T, N), gamma,
                                                   gae_lambda)).reshape(T, N, 1)
            tally.check(f'a2c/T{T}N{N}/gae{gae_lambda}/value_loss', value_loss.numpy(), F.smooth_l1_loss(values, ret).mean().numpy())
            tally.check(f'a2c/T{T}N{N}/gae{gae_lambda}/policy_loss', policy_loss.numpy(), (-((ret - values).detach() * log_probs).mean()).numpy())
            n += T * N
    return n


def validate_baseline_single(ref, tally, scale):
    """The BASELINE.json geometries: C3 (size 36, `default`), and the even sizes from 16 up on which the CUDA side runs
    its body-only kernel (16, 24), plus size 36 with a partial observation."""
    total = 0
    for S, N, mode, steps in [(36, 12, 'default', 50), (36, 8, 'partial_2', 50), (16, 32, 'partial_2', 60), (24, 16, 'partial_2', 60),
                              (24, 8, 'one_channel', 40)]:
        total += validate_single(ref, tally, N * scale, S, mode, steps * scale, seed=S * 7 + len(mode))
    return total


def validate_baseline_multi(ref, tally, scale):
    """The BASELINE.json geometries: C5 (16 snakes, size 64) in both observation modes, under the constructor defaults and
    the multiagent.py defaults; and the 32-snake maximum."""
    total = 0
    driver_rules = dict(food_mode='random_rate', food_rate=3e-4, respawn_mode='any', food_on_death_prob=0.33, boost_cost_prob=0.25)
    for (E, K, S, mode, steps, rules) in [(3, 16, 64, 'partial_4', 25, dict()), (2, 16, 64, 'full', 15, dict()),
                                          (2, 16, 64, 'partial_4', 20, driver_rules), (2, 32, 48, 'partial_3', 15, dict()),
                                          (2, 32, 64, 'full', 8, driver_rules)]:
        total += validate_multi(ref, tally, E * scale, K, S, mode, steps * scale, seed=E + K + S, **rules)
    return total


def only(args):
    ref = rl.load()
    tally = Tally()
    scale = 1 if args.quick else 4
    if args.only == 'a2c':
        total = validate_a2c(rl.REFERENCE_PATH, tally)
    elif args.only == 'single':
        total = validate_baseline_single(ref, tally, scale)
    elif args.only == 'multi':
        total = validate_baseline_multi(ref, tally, scale)
    else:
        total = validate_grid(ref, tally, 64 * scale, 7, 'default', 60 * scale, seed=7)
    print(f'{args.only}: {total} units, {tally.checks} comparisons, {len(tally.fails)} mismatches')
    if args.only == 'multi':
        print('multi events covered:', dict(EVENTS))
    if tally.fails:
        print('first failures:', tally.fails[:10])
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--only', default=None, choices=['single', 'a2c', 'grid', 'multi'], help='run one section only')
    args = ap.parse_args()
    if args.only:
        only(args)
        return
    ref = rl.load()
    tally = Tally()
    total = 0
    scale = 1 if args.quick else 4
    for mode in ['partial_2', 'partial_3', 'default', 'raw', 'one_channel', 'positions']:
        for S, N in [(9, 64), (12, 48), (17, 16)]:
            total += validate_single(ref, tally, N * scale, S, mode, 60 * scale, seed=S * 100 + len(mode))
    total += validate_baseline_single(ref, tally, scale)
    # degenerate: dead envs stepped again and again (not for partial_n: the reference raises there)
    for mode in ['default', 'raw', 'one_channel', 'positions']:
        total += validate_single(ref, tally, 32 * scale, 9, mode, 42 * scale, seed=7, reset_every_step=False)
    print(f'single: {total} env-steps, {tally.checks} tensor comparisons, {len(tally.fails)} mismatches')
    if tally.fails:
        print('first failures:', tally.fails[:10])
        sys.exit(1)
    tally = Tally()
    total = validate_a2c(rl.REFERENCE_PATH, tally)
    print(f'a2c returns: {total} (t, env) pairs through both losses, {tally.checks} comparisons, {len(tally.fails)} mismatches')
    if tally.fails:
        print('first failures:', tally.fails[:10])
        sys.exit(1)
    tally = Tally()
    total = 0
    for mode in ['default', 'raw']:
        for S, N in [(7, 64), (12, 32)]:
            total += validate_grid(ref, tally, N * scale, S, mode, 60 * scale, seed=S)
    total += validate_grid(ref, tally, 1, 7, 'positions', 60 * scale, seed=1)
    print(f'gridworld: {total} env-steps, {tally.checks} tensor comparisons, {len(tally.fails)} mismatches')
    if tally.fails:
        print('first failures:', tally.fails[:10])
        sys.exit(1)
    tally = Tally()
    total = 0
    rule_sets = [
        dict(),                                                                          # constructor defaults
        dict(food_mode='random_rate', food_rate=3e-3, respawn_mode='any', food_on_death_prob=0.33, boost_cost_prob=0.25),
        dict(boost=False, food_on_death_prob=0.0, reward_on_death=-2),
        dict(respawn_mode='any', food_on_death_prob=1.0, boost_cost_prob=1.0, agent_colours='fixed'),
    ]
    for rules in rule_sets:
        for (E, K, S, mode) in [(24, 2, 12, 'full'), (16, 4, 25, 'partial_4'), (12, 4, 12, 'partial_5'), (6, 7, 20, 'full')]:
            total += validate_multi(ref, tally, E * scale, K, S, mode, 40 * scale, seed=E + K + S, **rules)
    total += validate_baseline_multi(ref, tally, scale)
    print(f'multi: {total} env-steps, {tally.checks} tensor comparisons, {len(tally.fails)} mismatches')
    print('multi events covered:', dict(EVENTS))
    if tally.fails:
        print('first failures:', tally.fails[:10])
        sys.exit(1)


if __name__ == '__main__':
    main()
