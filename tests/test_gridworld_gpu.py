"""GPU: SimpleGridworld through the C ABI against the reference's golden vectors (replayed draws), the CPU
oracle (Philox draws) and the reference's own scenario tests (tests/test_simple_gridworld.py)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from golden_util import load, assert_same

pytestmark = pytest.mark.gpu
GRID = load('gridworld.npz')
DEV = 'cuda'


def np_(t):
    return t.detach().cpu().numpy()


def make_env(N, S, mode, **kw):
    from wurm_b200.envs import SimpleGridworld
    return SimpleGridworld(num_envs=N, size=S, observation_mode=mode, device=DEV, **kw)


@pytest.mark.parametrize('i', range(len(GRID)))
def test_golden_replay(i):
    tr = GRID[i]
    N, S, mode, start = tr.N, tr.S, tr.mode, tuple(int(v) for v in tr['start'])
    env = make_env(N, S, mode, start_location=start, manual_setup=True)
    if 'init_food' in tr:
        env.envs = env._create_envs(N, food_cell_replay=torch.from_numpy(tr['init_food']))
        assert_same(np_(env.envs), tr['init_envs'].astype(np.float32), 'created envs')
    else:
        env.envs = torch.from_numpy(tr['init_envs'].astype(np.float32)).to(DEV)
    for t in range(tr.steps):
        obs, reward, done, info = env.step(torch.from_numpy(tr[f'{t}/actions']).to(DEV),
                                           food_cell_replay=torch.from_numpy(tr[f'{t}/food_cell']))
        tag = f'gridworld trajectory {i} ({mode}, S={S}) step {t}: '
        assert_same(np_(env.envs), tr[f'{t}/envs'].astype(np.float32), tag + 'envs')
        assert_same(np_(reward).reshape(-1), tr[f'{t}/reward'], tag + 'reward')
        assert_same(np_(done).reshape(-1).astype(np.uint8), tr[f'{t}/done'], tag + 'done')
        assert_same(np_(info['edge_collision']).astype(np.uint8), tr[f'{t}/done'], tag + 'edge_collision')
        assert_same(np_(obs), tr[f'{t}/obs'], tag + 'observation')
        env.reset(done, food_cell_replay=torch.from_numpy(tr[f'{t}/reset_food']))
        assert_same(np_(env.envs), tr[f'{t}/reset_envs'].astype(np.float32), tag + 'envs after reset')


@pytest.mark.parametrize('N,S,mode,adtype', [(1000, 7, 'default', torch.long), (300, 13, 'raw', torch.int),
                                             (77, 32, 'positions', torch.short), (5, 5, 'default', torch.long)])
def test_rollout_matches_oracle(N, S, mode, adtype):
    seed, start = 99 + N, (S // 2, S // 2)
    env = make_env(N, S, mode, start_location=start, seed=seed)
    state = np.zeros((N, 2, S, S), np.float32)
    orc.grid_reset(state, np.ones(N, np.uint8), start, None, seed=seed, step=env._draws)
    assert_same(np_(env.envs), state, 'created envs')
    g = torch.Generator().manual_seed(seed)
    for t in range(40):
        a = torch.randint(0, 4, (N,), generator=g)
        obs, reward, done, info = env.step(a.to(device=DEV, dtype=adtype))
        r, d = orc.grid_step(state, a.numpy(), None, seed=seed, step=env._draws)
        assert_same(np_(env.envs), state, f'step {t}: envs')
        assert_same(np_(reward).reshape(-1), r, f'step {t}: reward')
        assert_same(np_(done).reshape(-1).astype(np.uint8), d, f'step {t}: done')
        assert_same(np_(obs), orc.grid_observe(state, mode), f'step {t}: observation')
        obs2 = env.reset(done)
        orc.grid_reset(state, d, start, None, seed=seed, step=env._draws)
        assert_same(np_(env.envs), state, f'step {t}: envs after reset')
        assert_same(np_(obs2), orc.grid_observe(state, mode), f'step {t}: observation after reset')


# ---- the reference's scenario tests (reference tests/test_simple_gridworld.py) ----
size = 7


def scenario(head):
    env = make_env(1, size, 'default', start_location=(3, 3), manual_setup=True)
    env.envs[0, 0, 1, 1] = 1
    env.envs[0, 1, head[0], head[1]] = 1
    return env


def test_basic_movement():
    env = scenario((3, 3))
    expected = [[4, 3], [4, 2], [3, 2], [3, 3], [2, 3], [2, 2]]
    for i, a in enumerate([0, 1, 2, 3, 2, 1]):
        env.step(torch.tensor([a], device=DEV))
        idx = env.envs[0, 1].flatten().argmax().item()
        assert [idx // size, idx % size] == expected[i]


def test_eat_food():
    env = scenario((2, 2))
    rewards = [env.step(torch.tensor([a], device=DEV))[1].item() for a in [0, 2, 2, 1]]
    assert rewards == [0, 0, 0, 1]
    assert env.envs[0, 0].sum().item() == 1


def test_edge_collision():
    env = scenario((3, 3))
    dones = [env.step(torch.tensor([0], device=DEV))[2].item() for _ in range(3)]
    assert dones == [False, False, True]
