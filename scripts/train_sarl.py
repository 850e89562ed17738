#!/usr/bin/env python
"""Value-network training (SARL by default; --policy cadrl | lstm_rl) on the B200 backend: mirror of crowd_nav/train.py:21-259
(config 4).

  imitation learning : ORCA robot (safety_space 0.15) -> replay (return-to-go targets) -> il_epochs of SGD
  reinforcement      : epsilon-greedy rollouts -> TD targets from the target network -> train_batches of SGD,
                       target sync every target_update_interval, evaluation on the val cases

One process per GPU (torchrun): every rank rolls out its own shard of episodes (disjoint case ids), keeps a
rank-local replay buffer, and all-reduces value-net gradients (one 386 kB bucket) and episode statistics over
NCCL -- never anything inside the env step / lookahead.

  python scripts/train_sarl.py --output_dir data/out [--il_episodes 500 --il_epochs 50 --train_episodes 200]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_sarl.py ...
"""
import argparse
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

from modelcrowdnav_b200.train_loop import TrainingLoop  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env_config", default=None)
    ap.add_argument("--policy_config", default=None)
    ap.add_argument("--output_dir", default="data/output")
    ap.add_argument("--il_episodes", type=int, default=500)       # train.config [imitation_learning]
    ap.add_argument("--il_epochs", type=int, default=50)
    ap.add_argument("--il_learning_rate", type=float, default=0.01)
    ap.add_argument("--safety_space", type=float, default=0.15)
    ap.add_argument("--rl_learning_rate", type=float, default=0.001)   # train.config [train]
    ap.add_argument("--train_batches", type=int, default=100)
    ap.add_argument("--train_episodes", type=int, default=200)
    ap.add_argument("--sample_episodes", type=int, default=64, help="episodes rolled out side by side per iteration")
    ap.add_argument("--target_update_interval", type=int, default=50)
    ap.add_argument("--evaluation_interval", type=int, default=100)
    ap.add_argument("--capacity", type=int, default=100000)
    ap.add_argument("--epsilon_start", type=float, default=0.5)
    ap.add_argument("--epsilon_end", type=float, default=0.1)
    ap.add_argument("--epsilon_decay", type=float, default=4000)
    ap.add_argument("--batch_size", type=int, default=100)
    ap.add_argument("--policy", default="sarl", choices=["sarl", "cadrl", "lstm_rl"])    # train.py --policy
    ap.add_argument("--precision", default="f16_tc")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--trainer", default="graph", choices=["eager", "graph", "fused"])
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    group = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    device = torch.device("cuda", local)
    torch.cuda.set_device(local)
    os.makedirs(a.output_dir, exist_ok=True)
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING,
                        format="%(asctime)s, %(levelname)s: %(message)s", datefmt="%Y-%m-%d %H:%M:%S")

    loop = TrainingLoop(device, rank, world, group, policy_name=a.policy, precision=a.precision, seed=a.seed,
                        env_config=a.env_config, policy_config=a.policy_config, capacity=a.capacity,
                        batch_size=a.batch_size, trainer_mode=a.trainer, sample_episodes=a.sample_episodes)
    env, policy, model, explorer, memory = loop.env, loop.policy, loop.model, loop.explorer, loop.memory

    # ---- imitation learning (train.py:144-178): rank r demonstrates cases [r * per_rank, (r + 1) * per_rank) ----
    t0 = time.time()
    loss = loop.imitation_learning(a.il_episodes, a.il_epochs, a.il_learning_rate, a.safety_space)
    logging.info("Imitation learning: %d experiences (rank 0), final epoch loss %.3e, %.1f s", len(memory), loss,
                 time.time() - t0)
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(a.output_dir, "il_model.pth"))

    # ---- reinforcement learning (train.py:180-246): iteration i, rank r rolls out its own block of sample_episodes cases ----
    loop.start_rl(a.rl_learning_rate)
    timers = {}
    while loop.iteration < a.train_episodes:
        episode = loop.iteration
        eps = (a.epsilon_start + (a.epsilon_end - a.epsilon_start) / a.epsilon_decay * episode
               if episode < a.epsilon_decay else a.epsilon_end)
        if episode % a.evaluation_interval == 0:
            env.case_counter["val"] = rank * (env.case_size["val"] // world)
            explorer.run_k_episodes(env.case_size["val"] // world, "val", episode=episode)
        loop.rl_iteration(eps, a.train_batches, a.target_update_interval, timers=timers)
        if rank == 0 and loop.iteration % 50 == 0:
            torch.save(model.state_dict(), os.path.join(a.output_dir, "rl_model.pth"))
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(a.output_dir, "rl_model.pth"))
        np.save(os.path.join(a.output_dir, "rl_model_flat.npy"), policy.flat_weights())
    # ---- final test (train.py:249) ----
    env.case_counter["test"] = rank * (env.case_size["test"] // world)
    out = explorer.run_k_episodes(env.case_size["test"] // world, "test", episode=loop.iteration, returnNav=True)
    if rank == 0:
        logging.info("rollout env-steps/s (rank 0, train phase incl. replay filling): %.0f",
                     loop.env_steps / max(timers.get("rollout_and_targets", 0.0), 1e-9))
        logging.info("seconds per phase over %d iterations: %s", loop.iteration, {k: round(v, 2) for k, v in timers.items()})
        logging.info("final test: reward %.4f success %.3f collision %.3f timeout %.3f nav %.2f", *out)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
