"""Single-agent rollout driver: the caller of the hot path.

Own copy of the *rollout* half of the reference's `experiments/main.py` (its call sequence, :195-227 and
:252-320, and the CLI flags that concern the env, :30-54), with the three defects of the shipped driver
repaired (trap 2: `A2C(model, ...)` signature, `RandomAgent` on flat
observations, log-prob/value broadcasting).  The env comes from `wurm_b200.envs`; the policy stand-ins
(uniform random, 2x64 feed-forward) are plain torch.  `--train true` runs the reference's A2C update
(main.py:232-246) with the return scan on the device (`wurm_b200.rl.A2C`); the
reference's conv / recurrent agents, model saving, video and CSV logging stay outside this repository's scope.

    python -m experiments.main --env snake --num-envs 512 --size 9 --agent random --observation partial_2 \
        --total-steps 1e6

Call sequence kept from the reference:  state = env.reset();  loop:  probs, value = model(state);
action ~ Categorical(probs);  state, reward, done, info = env.step(action);
env_consistency(env.envs[~done]);  env.reset(done)  -- whose observation is discarded (main.py:227), so
`return_observations=False` is passed here.

Multi-GPU: launched under `torch.distributed.run --nproc-per-node N` every rank owns an
independent slice of the `--num-envs` environments (`env_slice`), its own Philox seed (`rank_seed`) and its own copy of
the policy; there is no data-path collective.  Every LOG_INTERVAL steps the episode counters the step kernels keep are
summed over ranks with one small all-reduce (`env.stats(reduce_group=True)`) and rank 0 prints the job-wide line.
"""
import argparse
from itertools import count
from time import time

import torch
from torch import nn
from torch.distributions import Categorical

from wurm_b200.distributed import env_slice, finish, init_from_env, rank_seed
from wurm_b200.envs import SingleSnake, SimpleGridworld
from wurm_b200.rl import A2C
from wurm_b200.trajectory_store import TrajectoryStore

LOG_INTERVAL = 100
MAX_GRAD_NORM = 0.5          # reference main.py:27


def boolean(x):
    return x.lower()[0] == 't'


class RandomAgent(nn.Module):
    def __init__(self, num_actions, device):
        super().__init__()
        self.num_actions, self.device = num_actions, device

    def forward(self, x):
        n = x.shape[0]
        return torch.full((n, self.num_actions), 1.0 / self.num_actions, device=self.device), torch.zeros(n, device=self.device)


class FeedforwardAgent(nn.Module):
    """2 x 64 MLP on flat observations with action and value heads (the shape of the reference's agent)."""

    def __init__(self, num_actions, num_layers, hidden_units, num_inputs):
        super().__init__()
        layers, width = [], num_inputs
        for _ in range(num_layers):
            layers += [nn.Linear(width, hidden_units), nn.ReLU()]
            width = hidden_units
        self.body = nn.Sequential(*layers)
        self.action_head = nn.Linear(hidden_units, num_actions)
        self.value_head = nn.Linear(hidden_units, 1)

    def forward(self, x):
        x = self.body(x.reshape(x.shape[0], -1))
        return torch.softmax(self.action_head(x), dim=-1), self.value_head(x)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--env', type=str, default='snake')
    parser.add_argument('--num-envs', type=int, default=512)
    parser.add_argument('--size', type=int, default=9)
    parser.add_argument('--agent', type=str, default='random')
    parser.add_argument('--train', default=False, type=boolean)
    parser.add_argument('--observation', default='default', type=str)
    parser.add_argument('--render', default=False, type=boolean)
    parser.add_argument('--render-window-size', default=256, type=int)
    parser.add_argument('--render-cols', default=1, type=int)
    parser.add_argument('--render-rows', default=1, type=int)
    parser.add_argument('--total-steps', default=float('inf'), type=float)
    parser.add_argument('--total-episodes', default=float('inf'), type=float)
    parser.add_argument('--check-consistency', default=True, type=boolean,
                        help='run env_consistency on the live envs every step, as the reference driver does')
    parser.add_argument('--device', default='cuda', type=str)
    parser.add_argument('--seed', default=None, type=int)
    parser.add_argument('--lr', default=1e-3, type=float)
    parser.add_argument('--gamma', default=0.99, type=float)
    parser.add_argument('--update-steps', default=20, type=int)
    parser.add_argument('--entropy', default=0.0, type=float)
    parser.add_argument('--state', default='dense', type=str, choices=['dense', 'dense_scan', 'compact'],
                        help="'compact': the env lives in HBM as small records instead of the reference's dense fp32 tensors")
    args = parser.parse_args(argv)

    if args.env not in ('snake', 'gridworld'):
        raise ValueError('Unrecognised environment')
    if args.train and args.agent == 'random':
        raise ValueError('--train true needs a trainable agent')

    ranks = init_from_env(args.device)
    args.device = ranks.device
    total_envs = args.num_envs
    _, args.num_envs = env_slice(total_envs, ranks.rank, ranks.world_size)       # this rank's slice of the environments
    if args.seed is not None:
        args.seed = rank_seed(args.seed, ranks.rank) if ranks.world_size > 1 else args.seed
    render_args = {'size': args.render_window_size, 'num_rows': args.render_rows, 'num_cols': args.render_cols}
    if args.env == 'gridworld':                          # reference main.py:166-168
        env = SimpleGridworld(num_envs=args.num_envs, size=args.size, start_location=(args.size // 2, args.size // 2),
                              observation_mode=args.observation, device=args.device, seed=args.seed)
    else:
        env = SingleSnake(num_envs=args.num_envs, size=args.size, device=args.device, observation_mode=args.observation,
                          render_args=render_args, seed=args.seed, state=args.state)

    state = env.reset()
    if args.agent == 'random':
        model = RandomAgent(num_actions=4, device=args.device)
    elif args.agent == 'feedforward':
        model = FeedforwardAgent(num_actions=4, num_layers=2, hidden_units=64, num_inputs=state[0].numel()).to(args.device)
    else:
        raise ValueError('Unrecognised agent (this driver covers random and feedforward)')

    train = bool(args.train)
    if train:
        optimizer = torch.optim.Adam(model.parameters(), lr=args.lr)
        a2c = A2C(gamma=args.gamma)
        trajectories = TrajectoryStore(capacity=args.update_steps)       # rewards / dones / actions in a device ring

    num_steps = num_episodes = 0
    t0 = time()
    summary, losses = {}, {}
    for i_step in count(1):
        if args.render:
            env.render()

        with torch.set_grad_enabled(train):
            probs, state_value = model(state)
        action_distribution = Categorical(probs)
        action = action_distribution.sample().clone().long()

        state, reward, done, info = env.step(action)
        # NB `action` was sanitised in place by step() (single_snake.py:222), after sampling -- the log-prob below is that
        # of the action actually taken, as in the reference (main.py:218).

        if args.check_consistency and args.env == 'snake':
            env.check_consistency(skip=done)      # == env_consistency(env.envs[~done.squeeze(-1)]) (main.py:215), fused

        if train:
            trajectories.append(action=action, log_prob=action_distribution.log_prob(action).unsqueeze(-1),
                                value=state_value.reshape(-1, 1), reward=reward, done=done,
                                entropy=action_distribution.entropy().mean())

        env.reset(done, return_observations=False)

        if train and i_step % args.update_steps == 0:                 # reference main.py:232-246
            with torch.no_grad():
                _, bootstrap_values = model(state)
            value_loss, policy_loss = a2c.loss(bootstrap_values.reshape(-1, 1), trajectories.rewards, trajectories.values,
                                               trajectories.log_probs, trajectories.dones)
            entropy_loss = - trajectories.entropies.mean()
            optimizer.zero_grad()
            (value_loss + policy_loss + args.entropy * entropy_loss).backward()
            nn.utils.clip_grad_norm_(model.parameters(), MAX_GRAD_NORM)
            optimizer.step()
            trajectories.clear()
            losses = dict(value_loss=value_loss.item(), policy_loss=policy_loss.item())

        num_steps += total_envs                       # job-wide: every rank steps its slice in lockstep
        if i_step % LOG_INTERVAL == 0 or num_steps >= args.total_steps:
            # device counters kept by the step kernel, summed over ranks: one tiny all-reduce + D2H read
            stats = env.stats(reduce_group=True if ranks.world_size > 1 else None)
            num_episodes = stats['episodes']
            dt = time() - t0
            sizes = env.envs[:, -1].reshape(args.num_envs, -1).max(dim=-1)[0].sum().reshape(1)
            if ranks.world_size > 1:
                torch.distributed.all_reduce(sizes)
            summary = dict(steps=num_steps, episodes=num_episodes, reward_rate=stats['reward'] / max(stats['env_steps'], 1),
                           edge_collisions=stats['edge_collisions'], self_collisions=stats['self_collisions'],
                           avg_size=sizes.item() / total_envs, steps_per_second=num_steps / dt, ranks=ranks.world_size,
                           env_steps=stats['env_steps'], **losses)
            if ranks.is_main:
                print('\t'.join(f'{k}={v:.4g}' if isinstance(v, float) else f'{k}={v}' for k, v in summary.items()))

        if num_steps >= args.total_steps or num_episodes >= args.total_episodes:
            break
    env.check_status()
    finish(ranks)
    return summary


if __name__ == '__main__':
    main()
