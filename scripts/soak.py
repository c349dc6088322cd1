"""Randomised differential soak: CUDA path vs oracle on random geometries, rules and seeds for a time budget.

    python scripts/soak.py [seconds]      (GPU box; exits non-zero on the first mismatch)
"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import test_single_gpu as ts
import test_multi_gpu as tm

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = random.Random(int(time.time()))
t0 = time.time(); runs = 0
while time.time() - t0 < budget:
    state = rng.choice(['dense', 'dense_scan', 'compact'])
    if rng.random() < 0.5:
        S = rng.choice([9, 10, 11, 12, 13, 16, 18, 20, 22, 24, 30, 36, 40, 64, 90, 130])
        if S > 90:
            state = 'dense'
        mode = rng.choice(['default', 'raw', 'one_channel', 'positions', 'partial_1', 'partial_2', 'partial_3', 'partial_5'])
        N = rng.randint(1, 700)                      # (the rollout test seeds itself from N and S)
        steps = rng.choice([10, 25, 40])
        every = rng.choice([1, 1, 1, 3, 6])
        adt = rng.choice([torch.long, torch.int, torch.short])
        desc = f'single N={N} S={S} {mode} steps={steps} reset_every={every} {adt} state={state}'
        if rng.random() < 0.25 and every == 1:       # the fused step+reset launch against step-then-reset
            desc += ' (fused vs two calls)'
            ts.test_fused_step_reset_equals_step_then_reset(N, S, mode, state)
        else:
            ts.test_rollout_matches_oracle(N, S, mode, steps, every, adt, state)
    else:
        K = rng.choice([1, 2, 3, 4, 6, 8, 12, 16])
        S = rng.choice([8, 10, 12, 14, 20, 25, 30, 36, 48])
        E = rng.randint(1, 100)                      # (the rollout test seeds itself from E, K and S)
        mode = rng.choice(['full', 'partial_1', 'partial_2', 'partial_4'])
        steps = rng.choice([10, 25, 40])
        rules = rng.choice([dict(), dict(respawn_mode='any'), dict(respawn_mode='any', food_on_death_prob=1.0, boost_cost_prob=1.0),
                            dict(boost=False, food_on_death_prob=0.0, reward_on_death=-2),
                            dict(food_mode='random_rate', food_rate=3e-3, respawn_mode='any', food_on_death_prob=0.33, boost_cost_prob=0.25),
                            dict(food_mode='random_rate', food_rate=0.5, respawn_mode='any', food_on_death_prob=1.0)])     # flooded boards: the compact live list overflows
        desc = f'multi E={E} K={K} S={S} {mode} steps={steps} {rules} state={state}'
        if K > max(1, (S - 4) ** 2 // 25):           # leave room to place every snake
            continue
        if rng.random() < 0.25:
            desc += ' (fused vs two calls)'
            tm.test_fused_step_reset_equals_step_then_reset(E, K, S, mode, rules, state)
        else:
            tm.test_rollout_matches_oracle(E, K, S, mode, steps, rules, torch.long, state)
    runs += 1
    print(f'ok {runs}: {desc}', flush=True)
print(f'soak finished: {runs} rollouts in {time.time() - t0:.0f} s, no mismatch')
