"""GPU: the caller side of the boundary -- the rollout driver (the reference's experiments/main.py call
sequence) and the MultiSnake speed sweep (experiments/speeds.py) run on the CUDA path."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('agent,observation', [('random', 'partial_2'), ('feedforward', 'partial_2'),
                                               ('feedforward', 'default')])
def test_rollout_driver(agent, observation):
    """BASELINE config 1 (README example): snake, 512 envs, size 9, consistency checked every step."""
    from experiments.main import main
    summary = main(['--env', 'snake', '--num-envs', '512', '--size', '9', '--agent', agent, '--observation', observation,
                    '--total-steps', str(512 * 150), '--seed', '3'])
    assert summary['steps'] == 512 * 150
    assert summary['episodes'] > 0 and summary['edge_collisions'] > 0
    assert 3.0 <= summary['avg_size'] < 6.0


def test_drivers_on_the_compact_state():
    """`--state compact`: both drivers run their reference loops (consistency checked every step, on the records)."""
    from experiments.main import main
    summary = main(['--env', 'snake', '--num-envs', '512', '--size', '9', '--agent', 'feedforward', '--observation', 'partial_2',
                    '--total-steps', str(512 * 150), '--seed', '3', '--state', 'compact'])
    assert summary['steps'] == 512 * 150 and summary['episodes'] > 0 and 3.0 <= summary['avg_size'] < 6.0
    from experiments.multiagent import main as multi_main
    summary = multi_main(['--n-envs', '256', '--n-agents', '4', '--size', '25', '--obs', 'partial_4', '--total-steps',
                          str(256 * 120), '--seed', '2', '--state', 'compact'])
    assert summary['steps'] == 256 * 120 and summary['edge_collisions'] > 0 and summary['food'] > 0


def test_multisnake_speed_sweep():
    from experiments.speeds import sweep
    fps = sweep(num_agents=4, size=20, min_log2=4, max_log2=8, num_steps=5, check=True, verbose=False)
    assert [n for n, _ in fps] == [16, 32, 64, 128, 256]
    assert all(rate > 0 for _, rate in fps)


def test_rollout_driver_gridworld():
    from experiments.main import main
    summary = main(['--env', 'gridworld', '--num-envs', '256', '--size', '7', '--agent', 'feedforward', '--observation',
                    'default', '--total-steps', str(256 * 100), '--seed', '3'])
    assert summary['steps'] == 256 * 100 and summary['episodes'] > 0


def test_multiagent_rollout_driver_with_annealing():
    """The reference's multiagent.py defaults (random_rate food, respawn any) with food-rate annealing."""
    from experiments.multiagent import main
    summary = main(['--n-envs', '256', '--n-agents', '4', '--size', '25', '--obs', 'partial_4', '--total-steps',
                    str(256 * 120), '--food-rate', '3e-3', '--food-rate-min', '1e-3', '--seed', '2'])
    assert summary['steps'] == 256 * 120 and summary['edge_collisions'] > 0 and summary['food'] > 0
    assert abs(summary['food_rate'] - 1e-3) < 1e-4


def test_rollout_driver_trains_with_a2c():
    """`--train true`: the reference's A2C update (main.py:232-246) on the device return scan; losses stay finite and
    the trajectory store behaves like the reference's (stacked (T, N, 1) tensors)."""
    import math
    import torch
    from experiments import main as driver
    from wurm_b200.trajectory_store import TrajectoryStore
    torch.manual_seed(0)
    summary = driver.main(['--env', 'snake', '--num-envs', '256', '--size', '9', '--agent', 'feedforward', '--observation',
                           'partial_2', '--train', 'true', '--update-steps', '10', '--total-steps', str(256 * 200),
                           '--seed', '3'])
    assert summary['steps'] == 256 * 200
    assert math.isfinite(summary['value_loss']) and math.isfinite(summary['policy_loss'])

    store = TrajectoryStore()
    for t in range(3):
        store.append(reward=torch.full((4, 1), float(t)), done=torch.zeros(4, 1, dtype=torch.bool))
    assert store.rewards.shape == (3, 4, 1) and store.dones.dtype == torch.bool and len(store) == 3
    store.clear()
    assert len(store) == 0


def test_multiagent_driver_trains_with_a2c():
    """`--agent feedforward --train true`: the reference's multi-agent A2C update (multiagent.py:424-463) on the device
    return scan; finite losses, consistent envs throughout."""
    import math
    import torch
    from experiments import multiagent as driver
    torch.manual_seed(0)
    summary = driver.main(['--n-envs', '64', '--n-agents', '4', '--size', '25', '--obs', 'partial_4', '--agent', 'feedforward',
                           '--train', 'true', '--update-steps', '5', '--total-steps', str(64 * 100), '--seed', '3'])
    assert summary['steps'] == 64 * 100
    assert math.isfinite(summary['value_loss']) and math.isfinite(summary['policy_loss'])
