"""Multi-agent rollout driver: the caller of the MultiSnake hot path.

Own copy of the *rollout* half of the reference's `experiments/multiagent.py` (flags that concern the
env :47-93, env construction :247-251, hyper-parameter annealing by attribute assignment :337-345, the
call sequence :350-378): per step every agent samples an action from its policy's probabilities,
`observations, reward, done, info = env.step(actions)`, `env.reset(done['__all__'],
return_observations=False)`, `env.check_consistency()`.  The env comes from `wurm_b200.envs`; the policy
stand-in is a uniform random agent over the 8 actions (4 moves x boost) -- policies, species, DIAYN
and the A2C learner are outside this repository's scope.

    python -m experiments.multiagent --n-envs 4096 --n-agents 4 --size 25 --obs partial_4 --total-steps 1e6
"""
import argparse
from itertools import count
from time import time

import torch
from torch.distributions import Categorical

from wurm_b200.envs import MultiSnake

LOG_INTERVAL = 100


def boolean(x):
    return x.lower()[0] == 't'


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--env', type=str, default='snake')
    parser.add_argument('--n-envs', type=int, default=512)
    parser.add_argument('--n-agents', type=int, default=4)
    parser.add_argument('--size', type=int, default=25)
    parser.add_argument('--agent', type=str, nargs='+', default=['random'])
    parser.add_argument('--obs', type=str, default='partial_4')
    parser.add_argument('--boost', default=True, type=boolean)
    parser.add_argument('--train', default=False, type=boolean)
    parser.add_argument('--total-steps', default=float('inf'), type=float)
    parser.add_argument('--total-episodes', default=float('inf'), type=float)
    parser.add_argument('--device', default='cuda', type=str)
    parser.add_argument('--boost-cost', type=float, default=0.25)
    parser.add_argument('--food-on-death', type=float, default=0.33)
    parser.add_argument('--food-on-death-min', type=float, default=None)
    parser.add_argument('--reward-on-death', type=float, default=-1)
    parser.add_argument('--food-mode', type=str, default='random_rate')
    parser.add_argument('--food-rate', type=float, default=3e-4)
    parser.add_argument('--food-rate-min', type=float, default=None)
    parser.add_argument('--respawn-mode', type=str, default='any')
    parser.add_argument('--colour-mode', type=str, default='random')
    parser.add_argument('--check-consistency', default=True, type=boolean)
    parser.add_argument('--seed', default=None, type=int)
    args = parser.parse_args(argv)

    if args.env != 'snake':
        raise ValueError('Unrecognised environment')
    if args.train or args.agent != ['random']:
        raise NotImplementedError('policies and the A2C learner are outside the scope of wurm_b200 (DESIGN.md section 7); '
                                  'use --agent random --train false')

    env = MultiSnake(num_envs=args.n_envs, num_snakes=args.n_agents, size=args.size, device=args.device,
                     observation_mode=args.obs, boost=args.boost, boost_cost_prob=args.boost_cost,
                     food_on_death_prob=args.food_on_death, reward_on_death=args.reward_on_death, food_mode=args.food_mode,
                     food_rate=args.food_rate, respawn_mode=args.respawn_mode, agent_colours=args.colour_mode, seed=args.seed)

    num_actions = 8 if args.boost else 4
    uniform = torch.full((args.n_envs, num_actions), 1.0 / num_actions, device=args.device)
    observations = env.reset()
    num_steps = 0
    t0 = time()
    summary = {}
    for i_step in count(1):
        # hyper-parameter annealing: the reference mutates the env's attributes between steps (:337-345)
        if args.food_rate_min is not None:
            env.food_rate -= (args.food_rate - args.food_rate_min) / args.total_steps * args.n_envs
        if args.food_on_death_min is not None:
            env.food_on_death_prob -= (args.food_on_death - args.food_on_death_min) / args.total_steps * args.n_envs

        actions = {agent: Categorical(uniform).sample().clone().long() for agent, obs in observations.items()}

        observations, reward, done, info = env.step(actions)

        env.reset(done['__all__'], return_observations=False)
        if args.check_consistency:
            env.check_consistency()

        num_steps += args.n_envs
        if i_step % LOG_INTERVAL == 0 or num_steps >= args.total_steps:
            stats = env.stats()
            summary = dict(steps=num_steps, episodes=stats['episodes'], food=stats['reward'],
                           snake_collisions=stats['self_collisions'], edge_collisions=stats['edge_collisions'],
                           food_rate=env.food_rate, food_on_death_prob=env.food_on_death_prob,
                           fps=num_steps / (time() - t0))
            print('\t'.join(f'{k}={v:.4g}' if isinstance(v, float) else f'{k}={v}' for k, v in summary.items()))
        if num_steps >= args.total_steps or summary.get('episodes', 0) >= args.total_episodes:
            break
    env.check_status()
    return summary


if __name__ == '__main__':
    main()
