"""GPU: SingleSnake through the C ABI (wurm_b200.envs.SingleSnake) against
  (1) the reference's golden vectors, replaying the reference's own random draws,
  (2) the CPU oracle on seeded random rollouts with Philox-derived draws (same seed on both sides),
  (3) the reference's own scenario tests (tests/test_single_snake_env.py in the reference tree),
  (4) size-independent invariants at BASELINE.json's full size (2^20 envs).
Everything is compared bit for bit (fp32 as int32 patterns).
"""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from golden_util import load, assert_same

pytestmark = pytest.mark.gpu

SINGLE = load('single.npz') + load('single_baseline.npz')      # round-1 fixtures + the BASELINE.json geometries
DEV = 'cuda'


def make_env(N, S, mode, **kw):
    from wurm_b200.envs import SingleSnake
    return SingleSnake(num_envs=N, size=S, observation_mode=mode, device=DEV, **kw)


def np_(t):
    return t.detach().cpu().numpy()


STATES = ['dense', 'compact']     # state='compact': records resident in HBM, `envs` materialised for the comparison


@pytest.mark.parametrize('state', STATES)
@pytest.mark.parametrize('i', range(len(SINGLE)))
def test_golden_replay(i, state):
    """CUDA path == reference on the recorded trajectories (states, actions, rewards, dones, info, obs)."""
    tr = SINGLE[i]
    N, S, mode = tr.N, tr.S, tr.mode
    env = make_env(N, S, mode, manual_setup=True, state=state)
    if 'init_spawn' in tr:
        env.envs = env._create_envs(N, spawn_replay=torch.from_numpy(tr['init_spawn']))
        assert_same(np_(env.envs), tr['init_envs'].astype(np.float32), 'created envs')
    else:
        env.envs = torch.from_numpy(tr['init_envs'].astype(np.float32)).to(DEV)
    for t in range(tr.steps):
        a = torch.from_numpy(tr[f'{t}/actions_in'].copy()).to(DEV)
        obs, reward, done, info = env.step(a, food_cell_replay=torch.from_numpy(tr[f'{t}/food_cell']))
        tag = f'trajectory {i} ({mode}, S={S}) step {t}: '
        assert_same(np_(env.envs), tr[f'{t}/envs'].astype(np.float32), tag + 'envs')
        assert_same(np_(a), tr[f'{t}/actions_out'], tag + 'sanitised actions')
        assert reward.shape == (N, 1) and done.shape == (N, 1) and done.dtype == torch.bool
        assert_same(np_(reward).reshape(-1), tr[f'{t}/reward'], tag + 'reward')
        assert_same(np_(done).reshape(-1).astype(np.uint8), tr[f'{t}/done'], tag + 'done')
        assert_same(np_(info['self_collision']).astype(np.uint8), tr[f'{t}/self_collision'], tag + 'self_collision')
        assert_same(np_(info['edge_collision']).astype(np.uint8), tr[f'{t}/edge_collision'], tag + 'edge_collision')
        assert_same(np_(obs), tr[f'{t}/obs'], tag + 'observation')
        if bool(tr[f'{t}/did_reset']):
            obs2 = env.reset(done, spawn_replay=torch.from_numpy(tr[f'{t}/spawn']))
            assert_same(np_(env.envs), tr[f'{t}/reset_envs'].astype(np.float32), tag + 'envs after reset')
            assert_same(np_(obs2), tr[f'{t}/reset_obs'], tag + 'observation after reset')
    env.check_status()


ROLLOUTS = [
    # N, S, mode, steps, reset_every, action dtype
    (1000, 9, 'partial_2', 40, 1, torch.long),      # ragged last tile (1000 % 64 != 0)
    (777, 9, 'partial_3', 30, 1, torch.int),
    (513, 12, 'default', 30, 1, torch.short),
    (300, 10, 'raw', 30, 1, torch.long),
    (257, 17, 'one_channel', 30, 2, torch.long),
    (129, 36, 'default', 40, 1, torch.long),        # BASELINE config 3 geometry
    (66, 36, 'partial_4', 40, 1, torch.long),
    (31, 64, 'positions', 25, 1, torch.long),
    (5, 100, 'one_channel', 12, 1, torch.long),
    (400, 9, 'default', 28, 7, torch.long),         # dead envs stepped again: heads leave the grid
    (400, 11, 'positions', 28, 7, torch.long),
    (1, 9, 'partial_2', 20, 1, torch.long),
    (333, 16, 'partial_2', 40, 1, torch.long),      # body-only tiles (even sizes from 16 up), ragged last tile
    (200, 18, 'default', 30, 7, torch.int),         # ... with dead envs stepped again (general step on global memory)
    (77, 24, 'one_channel', 30, 3, torch.long),
    (5, 150, 'partial_3', 10, 1, torch.long),       # sides above 128: no shared-memory tile, the general step on global memory
    (3, 200, 'default', 6, 1, torch.int),           # ... above 181 also without hints (cell indices exceed int16)
    (4, 140, 'positions', 9, 3, torch.long),
    (3, 130, 'raw', 6, 1, torch.long),
]


@pytest.mark.parametrize('state', STATES)
@pytest.mark.parametrize('N,S,mode,steps,reset_every,adtype', ROLLOUTS)
def test_rollout_matches_oracle(N, S, mode, steps, reset_every, adtype, state):
    """Seeded random rollout, draws derived from Philox on both sides: every tensor identical."""
    if state == 'compact' and S > 90:
        pytest.skip('compact records carry sizes up to 90')
    seed = 1234 + N + S
    env = make_env(N, S, mode, seed=seed, state=state)
    state = np.zeros((N, 3, S, S), np.float32)
    orc.single_reset(state, np.ones(N, np.uint8), None, seed=seed, step=env._draws)
    assert_same(np_(env.envs), state, 'created envs')
    g = torch.Generator().manual_seed(seed)
    for t in range(steps):
        a_cpu = torch.randint(0, 4, (N,), generator=g)
        if t % 5 == 4:
            a_cpu = torch.randint(0, 9, (N,), generator=g)     # out-of-range actions go through fmod 4
        a = a_cpu.to(device=DEV, dtype=adtype)
        a_orc = a_cpu.numpy().astype(np.int64)
        obs, reward, done, info = env.step(a)
        r, d, sc, ec = orc.single_step(state, a_orc, None, seed=seed, step=env._draws)
        o, bad = orc.single_observe(state, mode)
        tag = f'step {t}: '
        assert_same(np_(env.envs), state, tag + 'envs')
        assert_same(np_(a).astype(np.int64), a_orc, tag + 'sanitised actions')
        assert_same(np_(reward).reshape(-1), r, tag + 'reward')
        assert_same(np_(done).reshape(-1).astype(np.uint8), d, tag + 'done')
        assert_same(np_(info['self_collision']).astype(np.uint8), sc, tag + 'self_collision')
        assert_same(np_(info['edge_collision']).astype(np.uint8), ec, tag + 'edge_collision')
        assert_same(np_(obs), o, tag + 'observation')
        if t % reset_every == reset_every - 1:
            obs2 = env.reset(done)
            orc.single_reset(state, d, None, seed=seed, step=env._draws)
            assert_same(np_(env.envs), state, tag + 'envs after reset')
            o2, _ = orc.single_observe(state, mode)
            assert_same(np_(obs2), o2, tag + 'observation after reset')
    if not mode.startswith('partial') and reset_every == 1:
        env.check_status()


def test_every_observe_mode_on_one_state():
    N, S = 200, 14
    env = make_env(N, S, 'partial_3', seed=5)
    for t in range(6):
        _, _, done, _ = env.step(torch.randint(0, 4, (N,), device=DEV))
        env.reset(done, return_observations=False)
    state = np_(env.envs)
    for mode in ['default', 'raw', 'one_channel', 'positions', 'partial_3']:
        o, bad = orc.single_observe(state, mode)
        assert bad == 0
        assert_same(np_(env._observe(mode)), o, mode)
    rgb = env._get_rgb()
    assert rgb.dtype == torch.short
    assert_same(np_(rgb).astype(np.float32) / np.float32(255), orc.single_observe(state, 'default')[0], '_get_rgb')


# ---- the reference's scenario tests, on the CUDA path (reference tests/test_single_snake_env.py) ----
size = 12


def get_test_env(orientation='up'):
    """The reference's fixture (wurm/utils.py:68-110)."""
    env = torch.zeros((1, 3, size, size))
    cells = {'up': ([(3, 3), (3, 4), (4, 4), (5, 4)], (6, 6)), 'right': ([(3, 3), (3, 4), (4, 4), (4, 5)], (6, 9)),
             'down': ([(8, 8), (7, 8), (6, 8), (5, 8)], (7, 2)), 'left': ([(8, 7), (7, 7), (6, 7), (6, 6)], (1, 2))}
    body, food = cells[orientation]
    for v, (y, x) in enumerate(body, 1):
        env[0, 2, y, x] = v
    env[0, 1, body[-1][0], body[-1][1]] = 1
    env[0, 0, food[0], food[1]] = 1
    return env


def head_position(env):
    idx = env.envs[0, 1].flatten().argmax().item()
    return [idx // size, idx % size]


def run_actions(actions, orientation='up'):
    env = make_env(1, size, 'one_channel', manual_setup=True, seed=20)
    env.envs = get_test_env(orientation).to(DEV)
    for a in actions:
        yield env, env.step(torch.tensor([a], device=DEV))


def test_basic_movement():
    expected = [[6, 4], [7, 4], [7, 5], [8, 5], [9, 5], [9, 4]]
    for i, (env, (obs, reward, done, info)) in enumerate(run_actions([0, 0, 3, 0, 0, 1])):
        assert head_position(env) == expected[i]
        assert not torch.any(done)


def test_cannot_move_backwards():
    expected = [[6, 4], [7, 4], [8, 4], [8, 5]]
    for i, (env, (obs, reward, done, info)) in enumerate(run_actions([2, 2, 2, 3])):
        assert head_position(env) == expected[i]
        assert not torch.any(done)


def test_eat_food():
    rewards = []
    for env, (obs, reward, done, info) in run_actions([0, 3, 3, 0, 0]):
        assert not torch.any(done)
        rewards.append(reward.item())
    assert rewards[2] == 1          # the food at (6,6) is reached by the third action
    assert env.envs[0, 2].max() == 4 + sum(rewards)     # (the respawned food may land on the path and be eaten too)
    assert env.envs[0, 0].sum() == 1
    assert_consistent(env.envs)


def test_hit_boundary():
    assert any(torch.any(done).item() for env, (obs, reward, done, info) in run_actions([1] * 10))


def test_hit_self():
    for env, (obs, reward, done, info) in run_actions([0, 3, 3, 2, 1, 0, 0, 0]):
        if torch.any(done):
            assert info['self_collision'].item()
            break
    else:
        assert False
    assert env.envs[0, 0].sum() == 1


def test_step_argument_errors():
    env = make_env(4, 9, 'default')
    with pytest.raises(TypeError):
        env.step(torch.zeros(4, device=DEV))
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(5, dtype=torch.long, device=DEV))
    with pytest.raises(NotImplementedError):
        make_env(4, 8, 'default')


def assert_consistent(envs):
    """The reference's invariants (wurm/utils.py:113-178 snake_consistency + env_consistency)."""
    n = envs.shape[0]
    food, head, body = envs[:, 0], envs[:, 1], envs[:, 2]
    assert torch.all((food == 0) | (food == 1))
    assert torch.all(head.reshape(n, -1).sum(-1) == 1)
    sizes = body.reshape(n, -1).max(-1)[0]
    assert torch.equal(sizes, (head * body).reshape(n, -1).sum(-1))
    totals = body.reshape(n, -1).sum(-1)
    assert torch.equal(totals, sizes * (sizes + 1) / 2)
    assert torch.all(totals >= 6)
    assert torch.all((head * food).reshape(n, -1).sum(-1) == 0)
    assert torch.all(food.reshape(n, -1).sum(-1) == 1)


def test_multiple_envs_invariants():
    """reference test_multiple_envs / test_setup / test_reset (:17-47)."""
    env = make_env(97, size, 'one_channel')
    assert_consistent(env.envs)
    assert torch.all(env.envs[:, 2].reshape(97, -1).sum(-1) == 6)
    actions = torch.randint(4, size=(100, 97), device=DEV)
    for a in actions:
        obs, reward, done, info = env.step(a)
        env.reset(done)
        assert_consistent(env.envs)
    env.check_status()


def test_full_size_invariants_and_food_uniformity():
    """BASELINE config 2 (2^20 envs, size 9, partial_2): invariants hold after every step+reset,
    episode statistics are plausible, and the Philox food placement covers exactly the free cells."""
    N, S = 1 << 20, 9
    env = make_env(N, S, 'partial_2', seed=99)
    assert_consistent(env.envs)
    food0 = env.envs[:, 0].reshape(N, -1).argmax(-1)
    counts = torch.bincount(food0, minlength=S * S).float().reshape(S, S)
    assert counts[0].sum() == 0 and counts[-1].sum() == 0 and counts[:, 0].sum() == 0 and counts[:, -1].sum() == 0
    assert counts[4, 4] == 0                           # the seed cell (4,4) always holds the snake
    assert (counts > 0).sum().item() == 48             # 49 interior cells minus the seed cell
    # cells never covered by the snake are chosen with probability 1/46 each
    corner = counts[1, 1].item()
    assert abs(corner - N / 46) < 6 * (N / 46) ** 0.5
    ate = 0
    for t in range(12):
        a = torch.randint(0, 4, (N,), device=DEV)
        obs, reward, done, info = env.step(a)
        assert obs.shape == (N, 75)
        ate += int(reward.sum().item())
        assert torch.equal(done.squeeze(-1), info['self_collision'] | info['edge_collision'])
        env.reset(done, return_observations=False)
        assert_consistent(env.envs)
    assert ate > 0
    env.check_status()



@pytest.mark.parametrize('compact,adtype', [(True, torch.uint8), (True, torch.long), (False, torch.long)])
def test_host_stepper_matches_direct_stepping(compact, adtype):
    """wurm_b200.HostStepper (pinned host buffers, pipelined copies) == stepping with device tensors; with
    `compact=True` the results cross PCIe as one packed byte per env and are decoded on the host, and uint8 actions
    upload one byte per env."""
    from wurm_b200 import HostStepper
    N, S, steps = 3000, 9, 12
    a = torch.randint(0, 4, (steps, N), generator=torch.Generator().manual_seed(3))
    direct = make_env(N, S, 'partial_2', seed=77)
    piped = make_env(N, S, 'partial_2', seed=77)
    stepper = HostStepper(piped, depth=2, return_actions=True, compact=compact, return_obs=True)
    expect = []
    for t in range(steps):
        acts = a[t].to(DEV)
        obs, reward, done, info = direct.step(acts)
        direct.reset(done, return_observations=False)
        expect.append((np_(obs), np_(reward), np_(done), np_(acts), np_(info['self_collision']), np_(info['edge_collision'])))
    tickets = [stepper.submit(a[t].to(adtype).pin_memory()) for t in range(steps)]      # far ahead of the waits
    for t, ticket in enumerate(tickets):
        ticket.wait()
        if t >= steps - 3:              # the slots of the last `depth + 1` tickets have not been recycled
            assert ticket.reward.shape == (N, 1) and ticket.done.shape == (N, 1) and ticket.done.dtype == torch.bool
            assert_same(ticket.reward.numpy(), expect[t][1], f'step {t}: reward')
            assert_same(ticket.done.numpy(), expect[t][2], f'step {t}: done')
            assert_same(ticket.actions.numpy().astype(np.int64), expect[t][3], f'step {t}: sanitised actions')
            assert_same(ticket.obs_host.numpy(), expect[t][0], f'step {t}: observation on the host')
            if compact:
                assert_same(ticket.self_collision.numpy(), expect[t][4], f'step {t}: self collision')
                assert_same(ticket.edge_collision.numpy(), expect[t][5], f'step {t}: edge collision')
    assert_same(np_(piped.envs), np_(direct.envs), 'final state')
    width = {torch.uint8: 1, torch.long: 8}[adtype]
    assert stepper.h2d_bytes_per_step == N * width
    assert stepper.d2h_bytes_per_step == N * ((1 if compact else 5) + width + 300)


def test_packed_result_byte_and_uint8_actions():
    """The packed byte carries done / self / edge / reward of every env, and uint8 actions step (and are sanitised)
    exactly like int64 ones -- on the tile kernel (size 9) and on the body-only kernel (size 36)."""
    for N, S, mode in [(777, 9, 'partial_2'), (130, 36, 'default')]:
        wide = make_env(N, S, mode, seed=5)
        narrow = make_env(N, S, mode, seed=5)
        g = torch.Generator().manual_seed(9)
        packed = torch.empty(N, dtype=torch.uint8, device=DEV)
        for t in range(25):
            a64 = torch.randint(0, 4, (N,), generator=g).to(DEV)
            a8 = a64.to(torch.uint8)
            o1, r1, d1, i1 = wide.step(a64, auto_reset=True)
            o2, r2, d2, i2 = narrow.step(a8, auto_reset=True, packed_out=packed)
            assert_same(np_(o2), np_(o1), f'{S} step {t}: obs')
            assert_same(np_(a8).astype(np.int64), np_(a64), f'{S} step {t}: sanitised actions')
            assert_same(np_(narrow.envs), np_(wide.envs), f'{S} step {t}: state')
            p = np_(packed)
            assert_same((p & 1) != 0, np_(d1).reshape(-1), 'packed done')
            assert_same((p & 2) != 0, np_(i1['self_collision']), 'packed self collision')
            assert_same((p & 4) != 0, np_(i1['edge_collision']), 'packed edge collision')
            assert_same(((p >> 3) & 3).astype(np.float32), np_(r1).reshape(-1), 'packed reward')
            assert not (p >> 5).any()


def test_graphed_stepper_is_bit_identical_to_call_by_call_stepping():
    """wurm_b200.GraphedStepper: one CUDA-graph launch per step+reset; same states, rewards, dones and
    observations as stepping the twin env call by call (the device-side call counter keeps the Philox
    draws in step)."""
    from wurm_b200 import GraphedStepper
    N, S, steps = 512, 9, 40
    plain = make_env(N, S, 'partial_2', seed=123)
    graphed = make_env(N, S, 'partial_2', seed=123)
    acts = torch.randint(0, 4, (steps + 2, N), generator=torch.Generator().manual_seed(1)).to(DEV)
    static_actions = torch.zeros(N, dtype=torch.long, device=DEV)
    # the stepper warms up with 2 real steps before capturing and then RESTORES the env (state, hints, statistics,
    # call counter): construction has no visible side effect
    static_actions.copy_(acts[0])
    stats_before = graphed.stats()
    stepper = GraphedStepper(graphed, static_actions, warmup=2)
    assert_same(np_(graphed.envs), np_(plain.envs), 'state after construction')
    assert graphed._draws == plain._draws and graphed.stats() == stats_before
    for t in range(1, steps + 1):
        a = acts[t].clone()
        if t % 7 == 3:
            # a DIRECT call interleaved with the replays: both paths keep drawing from one counter sequence
            a_direct = acts[t].clone()
            obs, reward, done, info = graphed.step(a_direct)
            graphed.reset(done, return_observations=False)
            sanitised = a_direct
        else:
            static_actions.copy_(acts[t])
            obs, reward, done, info = stepper.step()
            sanitised = static_actions
        obs2, reward2, done2, info2 = plain.step(a)
        assert_same(np_(obs), np_(obs2), f'step {t}: obs')
        assert_same(np_(reward), np_(reward2), f'step {t}: reward')
        assert_same(np_(done), np_(done2), f'step {t}: done')
        assert_same(np_(sanitised), np_(a), f'step {t}: sanitised actions')
        plain.reset(done2, return_observations=False)
        assert_same(np_(graphed.envs), np_(plain.envs), f'step {t}: state after reset')
        assert graphed._draws == plain._draws


@pytest.mark.parametrize('N,S,mode', [(1000, 9, 'partial_2'), (300, 12, 'default'), (130, 36, 'one_channel'), (257, 16, 'partial_3'),
                                      (65, 36, 'default'), (40, 20, 'positions')])
@pytest.mark.parametrize('state', STATES)
def test_fused_step_reset_equals_step_then_reset(N, S, mode, state):
    """step(a, auto_reset=True) == step(a); reset(done): same outputs, same state afterwards, same draws."""
    two_calls = make_env(N, S, mode, seed=55)
    fused = make_env(N, S, mode, seed=55, state=state)
    g = torch.Generator().manual_seed(8)
    for t in range(30):
        a = torch.randint(0, 4, (N,), generator=g).to(DEV)
        a2 = a.clone()
        obs, reward, done, info = two_calls.step(a)
        two_calls.reset(done, return_observations=False)
        obs_f, reward_f, done_f, info_f = fused.step(a2, auto_reset=True)
        tag = f'step {t}: '
        assert_same(np_(obs_f), np_(obs), tag + 'observation (terminal for finished envs)')
        assert_same(np_(reward_f), np_(reward), tag + 'reward')
        assert_same(np_(done_f), np_(done), tag + 'done')
        assert_same(np_(a2), np_(a), tag + 'sanitised actions')
        assert_same(np_(info_f['edge_collision']), np_(info['edge_collision']), tag + 'edge_collision')
        assert_same(np_(fused.envs), np_(two_calls.envs), tag + 'state after reset')
    assert fused._draws == two_calls._draws


@pytest.mark.parametrize('i', [i for i in range(len(SINGLE)) if 'init_spawn' in SINGLE[i] and bool(SINGLE[i]['0/did_reset'])])
def test_fused_step_reset_golden_replay(i):
    """The fused launch against the reference's recorded step+reset trajectories (both tapes replayed)."""
    tr = SINGLE[i]
    N, S, mode = tr.N, tr.S, tr.mode
    env = make_env(N, S, mode, manual_setup=True)
    env.envs = env._create_envs(N, spawn_replay=torch.from_numpy(tr['init_spawn']))
    for t in range(tr.steps):
        a = torch.from_numpy(tr[f'{t}/actions_in'].copy()).to(DEV)
        obs, reward, done, info = env.step(a, auto_reset=True, food_cell_replay=torch.from_numpy(tr[f'{t}/food_cell']),
                                           spawn_replay=torch.from_numpy(tr[f'{t}/spawn']))
        tag = f'trajectory {i} ({mode}, S={S}) step {t}: '
        assert_same(np_(obs), tr[f'{t}/obs'], tag + 'observation')
        assert_same(np_(reward).reshape(-1), tr[f'{t}/reward'], tag + 'reward')
        assert_same(np_(done).reshape(-1).astype(np.uint8), tr[f'{t}/done'], tag + 'done')
        assert_same(np_(env.envs), tr[f'{t}/reset_envs'].astype(np.float32), tag + 'envs after the fused reset')


def test_obs_out_renders_into_a_trajectory_buffer():
    """step(..., obs_out=buffer[t]): the kernel writes the observation straight into the caller's buffer."""
    N, S, T = 500, 9, 6
    a = make_env(N, S, 'partial_2', seed=4)
    b = make_env(N, S, 'partial_2', seed=4)
    ring = torch.zeros((T, N, 75), device=DEV)
    acts = torch.randint(0, 4, (T, N), generator=torch.Generator().manual_seed(4)).to(DEV)
    for t in range(T):
        obs, _, done, _ = a.step(acts[t].clone())
        a.reset(done, return_observations=False)
        obs_b, _, done_b, _ = b.step(acts[t].clone(), obs_out=ring[t])
        b.reset(done_b, return_observations=False)
        assert obs_b.data_ptr() == ring[t].data_ptr()
        assert_same(np_(ring[t]), np_(obs), f'step {t}')
    with pytest.raises(RuntimeError):
        b.step(acts[0].clone(), obs_out=torch.zeros((N, 74), device=DEV))


def boustrophedon(S):
    """Interior cells of an S x S grid in snake order (row 1 left to right, row 2 right to left, ...)."""
    cells = []
    for r in range(1, S - 1):
        cols = range(1, S - 1) if r % 2 == 1 else range(S - 2, 0, -1)
        cells += [(r, c) for c in cols]
    return cells


def test_nearly_full_board_food_respawn_and_no_free_cell():
    """A snake that fills almost the whole interior eats: the respawn has 0..6 free cells to choose from, which
    exercises the rejection-sampling misses, the explicit ranking fallback and the 'no free cell' case
    (the reference places nothing then, single_snake.py:306-320)."""
    S, path = 9, boustrophedon(9)
    delta = {(1, 0): 0, (0, -1): 1, (-1, 0): 2, (0, 1): 3}            # head displacement -> action
    lengths = [42, 43, 44, 45, 46, 47, 48] * 40                       # interior = 49 cells; food takes one more
    N = len(lengths)
    state = np.zeros((N, 3, S, S), np.float32)
    actions = np.zeros(N, np.int64)
    for e, L in enumerate(lengths):
        for v, (y, x) in enumerate(path[:L], 1):
            state[e, 2, y, x] = v
        hy, hx = path[L - 1]
        fy, fx = path[L]
        state[e, 1, hy, hx] = 1
        state[e, 0, fy, fx] = 1
        actions[e] = delta[(fy - hy, fx - hx)]
    env = make_env(N, S, 'partial_2', manual_setup=True, seed=2024)
    env.envs = torch.from_numpy(state.copy()).to(DEV)
    a = torch.from_numpy(actions.copy()).to(DEV)
    obs, reward, done, info = env.step(a)
    r, d, sc, ec = orc.single_step(state, actions, None, seed=2024, step=env._draws)
    assert (r == 1).all() and not d.any()
    assert_same(np_(env.envs), state, 'state after eating on a nearly full board')
    assert_same(np_(reward).reshape(-1), r, 'reward')
    assert_same(np_(obs), orc.single_observe(state, 'partial_2')[0], 'observation')
    food_left = state[:, 0].reshape(N, -1).sum(-1)
    assert (food_left[np.array(lengths) == 48] == 0).all()            # 49 body cells: nowhere to put the food
    assert (food_left[np.array(lengths) < 48] == 1).all()


def test_graphed_stepper_with_host_io():
    """GraphedStepper(host_io=True): the graph carries H2D actions and D2H rewards / done flags; results in the pinned
    host buffers equal call-by-call stepping of a twin env."""
    from wurm_b200 import GraphedStepper
    N, S, steps = 512, 9, 25
    plain = make_env(N, S, 'partial_2', seed=77)
    graphed = make_env(N, S, 'partial_2', seed=77)
    acts = torch.randint(0, 4, (steps + 1, N), generator=torch.Generator().manual_seed(5))
    static_actions = acts[0].to(DEV)
    stepper = GraphedStepper(graphed, static_actions, warmup=2, host_io=True)    # construction leaves the env untouched
    assert_same(np_(graphed.envs), np_(plain.envs), 'state after construction')
    for t in range(1, steps + 1):
        stepper.host_actions.copy_(acts[t])
        obs, _, _, _ = stepper.step_host()
        obs2, reward2, done2, _ = plain.step(acts[t].to(DEV))
        assert_same(stepper.host_reward.numpy(), np_(reward2), f'step {t}: host reward')
        assert_same(stepper.host_done.numpy(), np_(done2), f'step {t}: host done')
        assert_same(np_(obs), np_(obs2), f'step {t}: obs')
        plain.reset(done2, return_observations=False)
        assert_same(np_(graphed.envs), np_(plain.envs), f'step {t}: state after reset')
    with pytest.raises(RuntimeError):
        GraphedStepper(plain, static_actions.clone(), warmup=1).step_host()


@pytest.mark.parametrize('S', [9, 16])
def test_state_edited_between_calls_makes_hints_stale_not_wrong(S):
    """`envs` is a plain tensor the caller may write between calls (the reference's tests do).  The kernels keep
    (head cell, size) hints per env from one call to the next; after the caller shuffles, grows or replaces envs the
    hints no longer match and must only cost the fast path, never the result."""
    N, seed = 512, 99
    env = make_env(N, S, 'partial_2', seed=seed)
    g = torch.Generator().manual_seed(seed)
    for t in range(6):                                   # populate the hints
        _, _, done, _ = env.step(torch.randint(0, 4, (N,), generator=g).to(DEV))
        env.reset(done, return_observations=False)
    fresh = make_env(N, S, 'partial_2', seed=seed + 1)
    for edit in ('roll', 'grow', 'replace', 'roll'):
        if edit == 'roll':
            env.envs.copy_(env.envs.roll(1, dims=0))     # every env now sits under its neighbour's hints
        elif edit == 'grow':
            env.envs[:, 2].mul_(2.0)                     # body values 2, 4, 6, ...: sizes differ from the hinted ones
        else:
            env.envs.copy_(fresh.envs)                   # brand-new envs under the old hints
        for t in range(3):
            state = np_(env.envs).copy()
            a_cpu = torch.randint(0, 4, (N,), generator=g)
            a = a_cpu.to(DEV)
            obs, reward, done, info = env.step(a)
            a_orc = a_cpu.numpy().astype(np.int64)
            r, d, sc, ec = orc.single_step(state, a_orc, None, seed=seed, step=env._draws)
            tag = f'{edit} step {t}: '
            assert_same(np_(env.envs), state, tag + 'envs')
            assert_same(np_(a).astype(np.int64), a_orc, tag + 'sanitised actions')
            assert_same(np_(reward).reshape(-1), r, tag + 'reward')
            assert_same(np_(done).reshape(-1).astype(np.uint8), d, tag + 'done')
            if edit != 'grow':                           # (doubled bodies have no head on them: partial obs would flag it)
                o, _ = orc.single_observe(state, 'partial_2')
                assert_same(np_(obs), o, tag + 'observation')
            env.reset(done, return_observations=False)


@pytest.mark.parametrize('S,mode', [(36, 'default'), (16, 'partial_2'), (9, 'partial_2'), (24, 'one_channel')])
def test_second_food_cell_written_by_the_caller_is_eaten_like_in_the_reference(S, mode):
    """The reference's step handles any number of food cells (single_snake.py:242,270-272; only env_consistency
    rejects them).  In steady state the kernels run on verified hints (body-only tiles from size 16 up) that do not
    re-check that the hinted food cell is the ONLY one: a write through `env.envs` must drop the hints, so that a
    second food cell placed right in front of the head is eaten exactly as the oracle eats it."""
    N, seed = 96, 99
    env = make_env(N, S, mode, seed=seed)
    state = np.zeros((N, 3, S, S), np.float32)
    orc.single_reset(state, np.ones(N, np.uint8), None, seed=seed, step=env._draws)
    g = torch.Generator().manual_seed(3)

    def both(a_cpu, tag):
        a = a_cpu.to(DEV)
        obs, reward, done, info = env.step(a)
        a_np = a_cpu.numpy().copy()
        r, d, sc, ec = orc.single_step(state, a_np, None, seed=seed, step=env._draws)
        assert_same(np_(env.envs), state, tag + 'envs')
        assert_same(np_(reward).reshape(-1), r, tag + 'reward')
        assert_same(np_(done).reshape(-1).astype(np.uint8), d, tag + 'done')
        assert_same(np_(obs), orc.single_observe(state, mode)[0], tag + 'observation')
        return r, d

    for t in range(3):                                   # steady state: hints valid
        r, d = both(torch.randint(0, 4, (N,), generator=g), f'warm step {t}: ')
        env.reset(torch.from_numpy(d != 0).to(DEV), return_observations=False)
        orc.single_reset(state, d, None, seed=seed, step=env._draws)
    # second food cell on the cell the head will enter when it keeps going straight (action = 2-opposite of the
    # orientation is sanitised to "straight on": use the direction head - neck)
    body = state[:, 2].reshape(N, -1)
    head_cell = body.argmax(1)
    size = body.max(1)
    neck_cell = (body == (size - 1)[:, None]).argmax(1)
    step_to = 2 * head_cell - neck_cell                    # head + (head - neck)
    hy, hx = head_cell // S, head_cell % S
    ty, tx = step_to // S, step_to % S
    ok = (np.abs(ty - hy) + np.abs(tx - hx) == 1) & (ty >= 1) & (ty <= S - 2) & (tx >= 1) & (tx <= S - 2)
    ok &= state[:, 0].reshape(N, -1)[np.arange(N), np.clip(step_to, 0, S * S - 1)] == 0
    assert ok.sum() > N // 4
    rows = np.flatnonzero(ok)
    state[rows, 0, ty[rows], tx[rows]] = 1.0
    env.envs[torch.from_numpy(rows).to(DEV), 0, torch.from_numpy(ty[rows]).to(DEV), torch.from_numpy(tx[rows]).to(DEV)] = 1.0
    # the action that moves head -> step_to: DELTA[a] = [(+1,0),(0,-1),(-1,0),(0,+1)]
    dy, dx = ty - hy, tx - hx
    a = np.where(dy == 1, 0, np.where(dx == -1, 1, np.where(dy == -1, 2, 3))).astype(np.int64)
    r, d = both(torch.from_numpy(a), 'step onto the second food cell: ')
    assert (r[rows] == 1).all()
    for t in range(4):                                   # and the env carries on with two food cells on the board
        r, d = both(torch.randint(0, 4, (N,), generator=g), f'after step {t}: ')


def test_raw_pointer_writers_can_invalidate_hints_by_hand():
    """`invalidate_hints()` is the documented escape hatch for writes torch's version counter cannot see."""
    N, S, seed = 64, 36, 5
    env = make_env(N, S, 'default', seed=seed)
    a = torch.zeros(N, dtype=torch.long, device=DEV)
    env.step(a)
    assert int((env._hints[:, 0] >= 0).sum()) > 0
    env.invalidate_hints()
    assert int((env._hints >= 0).sum()) == 0
    key = env._hint_key
    env.step(a)
    assert env._hint_key == key                          # the env's own launches do not look like caller edits
    env.envs[0, 0, 1, 1] = 0.0                           # a torch write does
    assert (env.envs.data_ptr(), env.envs._version) != env._hint_key


def test_compact_state_takes_caller_edits_and_refuses_what_it_cannot_carry():
    """state='compact': `envs` is materialised on access; an assigned tensor (how the reference's tests install their
    fixtures) or an in-place write is folded back into the records before the next call, and a state the records cannot
    carry exactly raises instead of being rounded."""
    from wurm_b200.utils import get_test_env
    S = 12
    twin = make_env(1, S, 'default', manual_setup=True, seed=3)
    env = make_env(1, S, 'default', manual_setup=True, seed=3, state='compact')
    twin.envs = get_test_env(S, 'up').to(DEV)
    env.envs = get_test_env(S, 'up').to(DEV)
    for t, a in enumerate([0, 3, 3, 0, 0]):                 # the reference's test_eat_food sequence
        act = torch.tensor([a], device=DEV)
        o1, r1, d1, _ = twin.step(act.clone())
        o2, r2, d2, _ = env.step(act.clone())
        assert_same(np_(o2), np_(o1), f'step {t}: obs')
        assert_same(np_(r2), np_(r1), f'step {t}: reward')
        assert_same(np_(env.envs), np_(twin.envs), f'step {t}: state')
    env.check_consistency()
    twin.envs[0, 0, 1, 1] = 1.0; env.envs[0, 0, 1, 1] = 1.0  # an in-place write through the materialised tensor
    act = torch.tensor([1], device=DEV)
    twin.step(act.clone()); env.step(act.clone())
    assert env._dense is None                                # dropped by the state-changing call
    assert_same(np_(env.envs), np_(twin.envs), 'state after the edit')
    env.envs.data[0, 0, 2, 2] = 1.0; twin.envs[0, 0, 2, 2] = 1.0   # a write torch's version counter cannot see (`.data`) ...
    env.invalidate_hints()                                   # ... needs the manual form: the tensor is folded back in
    twin.step(act.clone()); env.step(act.clone())
    assert_same(np_(env.envs), np_(twin.envs), 'state after the invisible edit + invalidate_hints()')
    env.envs[0, 2, 3, 3] = 0.5                               # a non-integral body value: not a record
    with pytest.raises(RuntimeError, match='compact'):
        env.step(act.clone())


def test_graphed_and_host_steppers_on_the_compact_state():
    """GraphedStepper and HostStepper drive a state='compact' env exactly like a dense one (same draws, same results)."""
    from wurm_b200 import GraphedStepper, HostStepper
    N, S, steps = 512, 9, 20
    dense = make_env(N, S, 'partial_2', seed=21)
    graphed = make_env(N, S, 'partial_2', seed=21, state='compact')
    piped = make_env(N, S, 'partial_2', seed=21, state='compact')
    acts = torch.randint(0, 4, (steps, N), generator=torch.Generator().manual_seed(6))
    static = acts[0].to(DEV)
    gs = GraphedStepper(graphed, static, warmup=2)
    hs = HostStepper(piped, depth=2)
    assert_same(np_(graphed.envs), np_(dense.envs), 'state after constructing the graphed stepper')
    for t in range(steps):
        o1, r1, d1, _ = dense.step(acts[t].to(DEV), auto_reset=True)
        static.copy_(acts[t])
        o2, r2, d2, _ = gs.step()
        ticket = hs.submit(acts[t].to(torch.uint8).pin_memory()).wait()
        assert_same(np_(o2), np_(o1), f'step {t}: graphed obs')
        assert_same(np_(r2), np_(r1), f'step {t}: graphed reward')
        assert_same(np_(ticket.obs), np_(o1), f'step {t}: piped obs')
        assert_same(ticket.reward.numpy(), np_(r1), f'step {t}: piped reward')
        assert_same(ticket.done.numpy(), np_(d1), f'step {t}: piped done')
    assert_same(np_(graphed.envs), np_(dense.envs), 'final state (graphed)')
    assert_same(np_(piped.envs), np_(dense.envs), 'final state (piped)')
    out = torch.empty((N, 75), device=DEV)
    o3, _, _, _ = piped.step(acts[0].to(DEV), obs_out=out)
    assert o3.data_ptr() == out.data_ptr()
