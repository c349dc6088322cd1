"""GPU: the fused invariant checkers (wurm_single_check / wurm_multi_check) raise exactly when -- and with
the message with which -- the reference's env_consistency / check_consistency do (wurm/utils.py:113-178,
multi_snake.py:733-769)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def single_env(n=64, S=12):
    from wurm_b200.envs import SingleSnake
    env = SingleSnake(num_envs=n, size=S, observation_mode='one_channel', device=DEV, seed=5)
    g = torch.Generator().manual_seed(5)                 # a fixed rollout: the corruptions below must not depend on luck
    for _ in range(10):
        _, _, done, _ = env.step(torch.randint(0, 4, (n,), generator=g).to(DEV))
        env.reset(done, return_observations=False)
    return env


def torch_reference_check(envs):
    """The torch path of wurm_b200.utils (taken for non-contiguous input): the reference's formulas."""
    from wurm_b200.utils import env_consistency
    wide = torch.zeros((envs.shape[0], 3, envs.shape[2], envs.shape[3] + 1), device=envs.device)
    view = wide[..., :-1]
    view.copy_(envs)
    assert not view.is_contiguous()
    env_consistency(view)


CORRUPTIONS = [
    ('invalid food pixel', lambda e: e.__setitem__((3, 0, 2, 2), 0.5)),
    ('multiple num_heads', lambda e: e.__setitem__((3, 1, 1, 1), 1.0)),
    ("don't contain a snake", lambda e: (e[3, 2].zero_(), e[3, 1].zero_(), e[3, 1].__setitem__((5, 5), 1.0))),
    ('head not at the end', lambda e: (e[3, 1].zero_(), e[3, 1].__setitem__((1, 1), 1.0))),
    ('inconsistent values', lambda e: e[3, 2].mul_(2.0).sub_(e[3, 2].gt(0).float())),
    ('exactly one food', lambda e: e[3, 0].zero_()),
]


def test_single_checker_accepts_consistent_envs_and_honours_skip():
    from wurm_b200.utils import env_consistency
    env = single_env()
    env_consistency(env.envs)
    env.check_consistency()
    env.envs[7, 0].zero_()                               # env 7 loses its food
    with pytest.raises(RuntimeError):
        env.check_consistency()
    skip = torch.zeros(64, dtype=torch.bool, device=DEV)
    skip[7] = True
    env.check_consistency(skip=skip)                      # the driver's env.envs[~done] (main.py:215)
    env.check_consistency(skip=skip.unsqueeze(-1))


@pytest.mark.parametrize('fragment,corrupt', CORRUPTIONS)
def test_single_checker_messages_match_the_torch_path(fragment, corrupt):
    from wurm_b200.utils import env_consistency
    env = single_env()
    corrupt(env.envs)
    with pytest.raises(RuntimeError) as fused:
        env_consistency(env.envs)
    with pytest.raises(RuntimeError) as plain:
        torch_reference_check(env.envs)
    assert fragment in str(fused.value), str(fused.value)
    assert fragment in str(plain.value), str(plain.value)
    assert 'first at index 3' in str(fused.value)


def multi_env():
    from wurm_b200.envs import MultiSnake
    env = MultiSnake(num_envs=32, num_snakes=3, size=14, observation_mode='partial_2', device=DEV, seed=9)
    g = torch.Generator().manual_seed(9)                 # a fixed rollout: the corruptions below must not depend on luck
    for _ in range(15):
        actions = {f'agent_{k}': torch.randint(0, 8, (32,), generator=g).to(DEV) for k in range(3)}
        _, _, dones, _ = env.step(actions)
        env.reset(dones['__all__'], return_observations=False)
    return env


def living_agent(env):
    return int((~env.dones).nonzero()[0])


def cell_away_from_head(env, a):
    """An interior cell that is neither the head of snake `a` nor next to it."""
    hy, hx = divmod(int(env.heads[a, 0].flatten().argmax()), env.size)
    return (2, 2) if abs(hy - 2) + abs(hx - 2) > 2 else (env.size - 3, env.size - 3)


def test_multi_checker():
    env = multi_env()
    env.check_consistency()
    a = living_agent(env)
    # overlapping snakes: two individually consistent snakes crossing at (5,5) -- like the reference, the
    # per-snake checks come first, so only a state that passes them reports the overlap
    from wurm_b200.envs import MultiSnake
    env2 = MultiSnake(num_envs=2, num_snakes=2, size=12, observation_mode='full', device=DEV, manual_setup=True)
    for e in range(2):
        env2.foods[e, 0, 1, 1] = 1
        for v, (y, x) in enumerate([(5, 4), (5, 5), (5, 6)], 1):
            env2.bodies[2 * e, 0, y, x] = v
        env2.heads[2 * e, 0, 5, 6] = 1
        for v, (y, x) in enumerate([(8, 5), (9, 5), (10, 5)] if e == 0 else [(4, 5), (5, 5), (6, 5)], 1):
            env2.bodies[2 * e + 1, 0, y, x] = v
        env2.heads[2 * e + 1, 0, 10 if e == 0 else 6, 5] = 1
    with pytest.raises(RuntimeError, match='overlapping snakes.*first at index 1'):
        env2.check_consistency()
    # head not at the end of the body
    env3 = multi_env()
    a = living_agent(env3)
    y, x = cell_away_from_head(env3, a)
    env3.heads[a].zero_()
    env3.heads[a, 0, y, x] = 1
    with pytest.raises(RuntimeError, match='head not at the end'):
        env3.check_consistency()
    # dead snake with leftovers
    env4 = multi_env()
    env4.dones[5] = True
    if env4.bodies[5].sum() == 0:
        env4.bodies[5, 0, 3, 3] = 1
    with pytest.raises(RuntimeError, match='Dead snake'):
        env4.check_consistency()
    # two heads
    env5 = multi_env()
    a = living_agent(env5)
    y, x = cell_away_from_head(env5, a)
    env5.heads[a, 0, y, x] = 1
    with pytest.raises(RuntimeError, match='num_heads'):
        env5.check_consistency()


def test_compact_single_checker_works_on_the_records():
    """state='compact': check_consistency runs on the records (no fp32 materialisation), honours `skip`, and reports
    corruptions folded in from a caller-edited tensor with the reference's messages."""
    from wurm_b200.envs import SingleSnake
    n, S = 64, 12
    env = SingleSnake(num_envs=n, size=S, observation_mode='one_channel', device=DEV, seed=5, state='compact')
    g = torch.Generator().manual_seed(5)
    for _ in range(10):
        _, _, done, _ = env.step(torch.randint(0, 4, (n,), generator=g).to(DEV))
        env.check_consistency(skip=done)                     # the driver's per-step call: terminal envs skipped
        assert env._dense is None                            # nothing was materialised for it
        env.reset(done, return_observations=False)
    env.check_consistency()
    for needle, corrupt in [('multiple num_heads', lambda e: e.__setitem__((3, 1, 1, 1), 1.0)),
                            ('exactly one food', lambda e: e[3, 0].zero_()),
                            ('head not at the end', lambda e: (e[3, 1].zero_(), e[3, 1].__setitem__((1, 1), 1.0)))]:
        saved = env.envs.clone()
        corrupt(env.envs)                                    # through the materialised tensor: folded back on the next call
        with pytest.raises(RuntimeError, match=needle):
            env.check_consistency()
        skip = torch.zeros(n, dtype=torch.bool, device=DEV); skip[3] = True
        env.check_consistency(skip=skip)
        env.envs = saved
        env.check_consistency()
