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
import configparser
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402
from modelcrowdnav_b200.trainer import Trainer  # noqa: E402

ENV_DEFAULT = dict(env=dict(time_limit=25, time_step=0.25, val_size=100, test_size=500, randomize_attributes="false"),
                   reward=dict(success_reward=1, collision_penalty=-0.25, discomfort_dist=0.2,
                               discomfort_penalty_factor=0.5),
                   sim=dict(train_val_sim="circle_crossing", test_sim="circle_crossing", square_width=10,
                            circle_radius=4, human_num=5),
                   humans=dict(visible="true", policy="orca", radius=0.3, v_pref=1, sensor="coordinates"),
                   robot=dict(visible="false", policy="none", radius=0.3, v_pref=1, sensor="coordinates"))
POLICY_DEFAULT = dict(rl=dict(gamma=0.9), om=dict(cell_num=4, cell_size=1, om_channel_size=3),
                      action_space=dict(kinematics="holonomic", speed_samples=5, rotation_samples=16,
                                        sampling="exponential", query_env="false"),
                      sarl=dict(mlp1_dims="150, 100", mlp2_dims="100, 50", attention_dims="100, 100, 1",
                                mlp3_dims="150, 100, 100, 1", multiagent_training="true", with_om="false",
                                with_global_state="true"),
                      cadrl=dict(mlp_dims="150, 100, 100, 1", multiagent_training="false"),
                      lstm_rl=dict(global_state_dim=50, mlp1_dims="150, 100, 100, 50", mlp2_dims="150, 100, 100, 1",
                                   multiagent_training="true", with_om="false", with_interaction_module="false"))


def make_config(default, path=None):
    cp = configparser.RawConfigParser()
    cp.read_dict({k: {kk: str(vv) for kk, vv in v.items()} for k, v in default.items()})
    if path:
        cp.read(path)
    return cp


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

    env_config, policy_config = make_config(ENV_DEFAULT, a.env_config), make_config(POLICY_DEFAULT, a.policy_config)
    torch.manual_seed(a.seed)
    policy = mcn.policy_factory[a.policy]()
    policy.configure(policy_config)
    if a.policy == "sarl" and not policy.with_om:
        policy.precision = a.precision            # CADRL, LSTM-RL and occupancy maps run on the FP32 path
    policy.set_device(device)
    env = mcn.CrowdSim()
    env.configure(env_config)
    env.device = local
    robot = mcn.Robot(env_config, "robot")
    env.set_robot(robot)
    # shard the case ids: rank r starts r * sample_episodes into the case list and strides by world
    env.case_counter["train"] = rank * a.sample_episodes
    memory = mcn.ReplayMemory(a.capacity)
    model = policy.get_model()
    trainer = Trainer(model, memory, device, a.batch_size, dist_group=group, policy=policy)
    trainer.broadcast_weights()
    explorer = mcn.Explorer(env, robot, device, memory, policy.gamma, target_policy=policy, dist_group=group)
    env.case_size["train"] = np.iinfo(np.uint32).max - 2000       # upstream CrowdNav value (the fork shrank it to 100)

    # ---- imitation learning (train.py:144-178) ----
    t0 = time.time()
    il_policy = mcn.policy_factory["orca"]()
    il_policy.multiagent_training = policy.multiagent_training
    il_policy.safety_space = a.safety_space
    il_policy.set_device(device)
    robot.set_policy(il_policy)
    per_rank = (a.il_episodes + world - 1) // world
    explorer.run_k_episodes(per_rank, "train", update_memory=True, imitation_learning=True)
    trainer.set_learning_rate(a.il_learning_rate)
    loss = trainer.optimize_epoch(a.il_epochs)
    logging.info("Imitation learning: %d experiences (rank 0), final epoch loss %.3e, %.1f s", len(memory), loss,
                 time.time() - t0)
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(a.output_dir, "il_model.pth"))
    explorer.update_target_model(model)

    # ---- reinforcement learning (train.py:180-246) ----
    robot.set_policy(policy)
    policy.set_env(env)
    trainer.set_learning_rate(a.rl_learning_rate)
    episode = 0
    steps_done, t_roll = 0, 0.0
    while episode < a.train_episodes:
        eps = (a.epsilon_start + (a.epsilon_end - a.epsilon_start) / a.epsilon_decay * episode
               if episode < a.epsilon_decay else a.epsilon_end)
        policy.set_epsilon(eps)
        if episode % a.evaluation_interval == 0:
            env.case_counter["val"] = rank * (env.case_size["val"] // world)
            explorer.run_k_episodes(env.case_size["val"] // world, "val", episode=episode)
        t1 = time.time()
        env.case_counter["train"] += (world - 1) * a.sample_episodes       # skip the other ranks' cases
        explorer.run_k_episodes(a.sample_episodes, "train", update_memory=True, episode=episode)
        t_roll += time.time() - t1
        steps_done += int(explorer.last_run["steps"].sum())
        trainer.optimize_batch(a.train_batches)
        episode += 1
        if episode % a.target_update_interval == 0:
            explorer.update_target_model(model)
        if rank == 0 and episode % 50 == 0:
            torch.save(model.state_dict(), os.path.join(a.output_dir, "rl_model.pth"))
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(a.output_dir, "rl_model.pth"))
        np.save(os.path.join(a.output_dir, "rl_model_flat.npy"), policy.flat_weights())
    # ---- final test (train.py:249) ----
    env.case_counter["test"] = rank * (env.case_size["test"] // world)
    out = explorer.run_k_episodes(env.case_size["test"] // world, "test", episode=episode, returnNav=True)
    if rank == 0:
        logging.info("rollout env-steps/s (rank 0, train phase incl. replay filling): %.0f",
                     steps_done / max(t_roll, 1e-9))
        logging.info("final test: reward %.4f success %.3f collision %.3f timeout %.3f nav %.2f", *out)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
