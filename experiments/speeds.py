"""MultiSnake env-steps/s over a sweep of batch sizes: the reference's own speed script
(`experiments/speeds.py:15-44`: 10 agents, size 36, num_envs 2^4..2^12, respawn_mode 'any',
step + reset(done['__all__']) + check_consistency) on the B200-native env, CUDA-event timed.

    python -m experiments.speeds [--num-agents 10] [--size 36] [--max-log2 16] [--check false]
"""
import argparse

import torch

from wurm_b200.envs import MultiSnake


def sweep(num_agents=10, size=36, min_log2=4, max_log2=12, num_steps=10, check=True, device='cuda', verbose=True):
    fps = []
    for n in [2 ** i for i in range(min_log2, max_log2 + 1)]:
        env = MultiSnake(num_envs=n, num_snakes=num_agents, size=size, manual_setup=False, boost=True, verbose=False,
                         device=device, respawn_mode='any')
        all_actions = {f'agent_{i}': torch.randint(8, size=(num_steps + 2, n), device=device) for i in range(num_agents)}

        def one_step(i):
            actions = {agent: agent_actions[i] for agent, agent_actions in all_actions.items()}
            observations, reward, done, info = env.step(actions)
            env.reset(done['__all__'])
            if check:
                env.check_consistency()

        one_step(0); one_step(1)            # warm-up
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for i in range(2, num_steps + 2):
            one_step(i)
        stop.record()
        torch.cuda.synchronize()
        rate = n * num_steps / (start.elapsed_time(stop) * 1e-3)
        if verbose:
            print(n, rate)
        fps.append((n, rate))
    return fps


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--num-agents', type=int, default=10)
    parser.add_argument('--size', type=int, default=36)
    parser.add_argument('--max-log2', type=int, default=12)
    parser.add_argument('--check', type=lambda x: x.lower()[0] == 't', default=True)
    args = parser.parse_args()
    sweep(args.num_agents, args.size, max_log2=args.max_log2, check=args.check)
