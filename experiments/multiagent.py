"""Multi-agent rollout driver: the caller of the MultiSnake hot path.

Own copy of the *rollout* half of the reference's `experiments/multiagent.py` (flags that concern the
env :47-93, env construction :247-251, hyper-parameter annealing by attribute assignment :337-345, the
call sequence :350-378): per step every agent samples an action from its policy's probabilities,
`observations, reward, done, info = env.step(actions)`, `env.reset(done['__all__'],
return_observations=False)`, `env.check_consistency()`.  The env comes from `wurm_b200.envs`; the policy
stand-ins are a uniform random agent over the 8 actions (4 moves x boost) and one shared 2x64 feed-forward
network (`--agent feedforward`).  `--train true` runs the reference's A2C update (:424-463: per-agent
trajectories flattened agent-major, one bootstrap pass, Adam, gradient clipping) with the return scan on the
device (`wurm_b200.rl.A2C`); species, shared backbones, recurrent agents and DIAYN stay outside this
repository's scope.

    python -m experiments.multiagent --n-envs 4096 --n-agents 4 --size 25 --obs partial_4 --total-steps 1e6

Multi-GPU (BASELINE config 5 is 32 768 envs per GPU across 8 GPUs): under `torch.distributed.run
--nproc-per-node N` every rank owns an independent slice of the `--n-envs` environments, its own Philox seed and its own
policy copy; the only collective is the all-reduce of the episode counters every LOG_INTERVAL steps, printed by rank 0.
"""
import argparse
from itertools import count
from time import time

import torch
from torch import nn
from torch.distributions import Categorical

from experiments.main import FeedforwardAgent
from wurm_b200.distributed import env_slice, finish, init_from_env, rank_seed
from wurm_b200.envs import MultiSnake
from wurm_b200.rl import A2C
from wurm_b200.trajectory_store import TrajectoryStore

LOG_INTERVAL = 100
MAX_GRAD_NORM = 0.5          # reference multiagent.py:31-32
VALUE_LOSS_COEFF = 0.5


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
    parser.add_argument('--lr', default=1e-3, type=float)
    parser.add_argument('--gamma', default=0.99, type=float)
    parser.add_argument('--update-steps', default=5, type=int)
    parser.add_argument('--entropy', default=0.0, type=float)
    parser.add_argument('--state', default='dense', type=str, choices=['dense', 'dense_scan', 'compact'],
                        help="'compact': the env lives in HBM as small records instead of the reference's dense fp32 tensors")
    args = parser.parse_args(argv)

    if args.env != 'snake':
        raise ValueError('Unrecognised environment')
    agent_type = args.agent[0]
    if agent_type not in ('random', 'feedforward'):
        raise ValueError('Unrecognised agent (this driver covers random and feedforward)')
    train = bool(args.train) and agent_type != 'random'

    ranks = init_from_env(args.device)
    args.device = ranks.device
    total_envs = args.n_envs
    _, args.n_envs = env_slice(total_envs, ranks.rank, ranks.world_size)         # this rank's slice of the environments
    if args.seed is not None and ranks.world_size > 1:
        args.seed = rank_seed(args.seed, ranks.rank)

    env = MultiSnake(num_envs=args.n_envs, num_snakes=args.n_agents, size=args.size, device=args.device,
                     observation_mode=args.obs, boost=args.boost, boost_cost_prob=args.boost_cost,
                     food_on_death_prob=args.food_on_death, reward_on_death=args.reward_on_death, food_mode=args.food_mode,
                     food_rate=args.food_rate, respawn_mode=args.respawn_mode, agent_colours=args.colour_mode, seed=args.seed,
                     state=args.state)

    num_actions = 8 if args.boost else 4
    K, E = args.n_agents, args.n_envs
    agents = [f'agent_{k}' for k in range(K)]
    uniform = torch.full((K * E, num_actions), 1.0 / num_actions, device=args.device)
    observations = env.reset()
    model = None
    if agent_type == 'feedforward':
        model = FeedforwardAgent(num_actions, 2, 64, num_inputs=observations['agent_0'][0].numel()).to(args.device)
    if train:
        optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
        a2c = A2C(gamma=args.gamma)
        trajectories = TrajectoryStore(capacity=args.update_steps)       # rewards / dones / actions in a device ring
    losses = {}

    def flat(d):                                     # the reference's flatten_dict: agent-major (K*E, 1)
        return torch.stack([d[a] for a in agents]).reshape(K * E, 1)
    num_steps = 0
    t0 = time()
    summary = {}
    for i_step in count(1):
        # hyper-parameter annealing: the reference mutates the env's attributes between steps (:337-345)
        if args.food_rate_min is not None:
            env.food_rate -= (args.food_rate - args.food_rate_min) / args.total_steps * total_envs
        if args.food_on_death_min is not None:
            env.food_on_death_prob -= (args.food_on_death - args.food_on_death_min) / args.total_steps * total_envs

        if model is None:
            dist = Categorical(uniform)
        else:
            with torch.set_grad_enabled(train):
                probs, values = model(torch.cat([observations[a] for a in agents]))       # one pass for all agents
            dist = Categorical(probs)
        sampled = dist.sample().clone().long()
        actions = {a: sampled[k * E:(k + 1) * E].clone() for k, a in enumerate(agents)}

        observations, reward, done, info = env.step(actions)

        if train:
            trajectories.append(action=sampled, log_prob=dist.log_prob(sampled).unsqueeze(-1), value=values.reshape(-1, 1),
                                reward=flat(reward), done=flat(done), entropy=dist.entropy().mean())

        env.reset(done['__all__'], return_observations=False)
        if args.check_consistency:
            env.check_consistency()

        if train and i_step % args.update_steps == 0:                 # reference multiagent.py:424-463
            with torch.no_grad():
                _, bootstrap_values = model(torch.cat([observations[a] for a in agents]))
            value_loss, policy_loss = a2c.loss(bootstrap_values.reshape(-1, 1), trajectories.rewards, trajectories.values,
                                               trajectories.log_probs, trajectories.dones)
            entropy_loss = - trajectories.entropies.mean()
            optimizer.zero_grad()
            (VALUE_LOSS_COEFF * value_loss + policy_loss + args.entropy * entropy_loss).backward()
            nn.utils.clip_grad_norm_(model.parameters(), MAX_GRAD_NORM)
            optimizer.step()
            trajectories.clear()
            losses = dict(value_loss=value_loss.item(), policy_loss=policy_loss.item())

        num_steps += total_envs                       # job-wide: every rank steps its slice in lockstep
        if i_step % LOG_INTERVAL == 0 or num_steps >= args.total_steps:
            stats = env.stats(reduce_group=True if ranks.world_size > 1 else None)      # summed over ranks (one all-reduce)
            summary = dict(steps=num_steps, episodes=stats['episodes'], food=stats['reward'],
                           snake_collisions=stats['self_collisions'], edge_collisions=stats['edge_collisions'],
                           food_rate=env.food_rate, food_on_death_prob=env.food_on_death_prob,
                           fps=num_steps / (time() - t0), ranks=ranks.world_size, env_steps=stats['env_steps'], **losses)
            if ranks.is_main:
                print('\t'.join(f'{k}={v:.4g}' if isinstance(v, float) else f'{k}={v}' for k, v in summary.items()))
        if num_steps >= args.total_steps or summary.get('episodes', 0) >= args.total_episodes:
            break
    env.check_status()
    finish(ranks)
    return summary


if __name__ == '__main__':
    main()
