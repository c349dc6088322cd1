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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    args = ap.parse_args()
    ref = rl.load()
    tally = Tally()
    total = 0
    scale = 1 if args.quick else 4
    for mode in ['partial_2', 'partial_3', 'default', 'raw', 'one_channel', 'positions']:
        for S, N in [(9, 64), (12, 48), (17, 16)]:
            total += validate_single(ref, tally, N * scale, S, mode, 60 * scale, seed=S * 100 + len(mode))
    # degenerate: dead envs stepped again and again (not for partial_n: the reference raises there)
    for mode in ['default', 'raw', 'one_channel', 'positions']:
        total += validate_single(ref, tally, 32 * scale, 9, mode, 42 * scale, seed=7, reset_every_step=False)
    print(f'single: {total} env-steps, {tally.checks} tensor comparisons, {len(tally.fails)} mismatches')
    if tally.fails:
        print('first failures:', tally.fails[:10])
        sys.exit(1)


if __name__ == '__main__':
    main()
