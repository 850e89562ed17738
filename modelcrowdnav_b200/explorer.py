"""Batched episode driver: mirror of crowd_nav/utils/explorer.py:13-193 and crowd_nav/utils/memory.py:4-34.

`Explorer.run_k_episodes(k, phase, ...)` keeps the reference's signature, log lines and return values, but the
k episodes run side by side as k environments of one GPU batch (cases c, c+1, ... exactly as k successive
`env.reset(phase)` calls would pick them).  The inner loop of the reference (explorer.py:62-69: act, step) is one
`cn_rollout_step`-style sequence per step for all envs.  Phases that do not fill the replay memory (val / test, the bulk of
test.py) run WITHOUT a host sync per step: the kernels are enqueued back to back, the per-env episode accumulators of the
library (cn_env_read_episode_table) deliver every episode's outcome, end time, step count, danger count and discounted return
at the end, and the loop only looks at the frozen flags every few steps to stop early.  Replay-filling phases read rewards /
done / info codes back every step (the stored state sequence needs them).

Replay filling (`update_memory`, explorer.py:153-186) is batched as well: states are kept as (T, k, H, 13) device
tensors, IL targets are discounted returns-to-go, RL targets r + gamma_bar * V_target(s') with V_target evaluated
by the FP32 CUDA value network (cn_policy_forward).
"""
import copy
import logging

import numpy as np

from . import _capi, scenes
from .batch import BatchedCrowdSim
from .envs import ModelCrowdSim, batch_env_kwargs
from .policy import CADRL, ORCA, SARL, Linear


class ReplayMemory(object):
    """Ring buffer of (state (H,13) fp32, value (1,) fp32) pairs (memory.py:4-34), stored as two device tensors."""

    def __init__(self, capacity, device=None):
        self.capacity = capacity
        self.device = device
        self.states = None       # (capacity, H, 13)
        self.values = None       # (capacity, 1)
        self.size = 0
        self.position = 0

    def _ensure(self, state):
        import torch
        if self.states is None:
            dev = self.device if self.device is not None else state.device
            self.states = torch.zeros((self.capacity,) + tuple(state.shape), dtype=torch.float32, device=dev)
            self.values = torch.zeros((self.capacity, 1), dtype=torch.float32, device=dev)

    def push(self, item):
        state, value = item
        self._ensure(state)
        self.states[self.position] = state
        self.values[self.position] = value.reshape(1)
        self.size = max(self.size, self.position + 1) if self.size < self.capacity else self.capacity
        self.position = (self.position + 1) % self.capacity

    def push_batch(self, states, values):
        """states (n, H, 13), values (n,) -- same ring semantics as n successive push() calls."""
        import torch
        n = states.shape[0]
        if n == 0:
            return
        self._ensure(states[0])
        idx = (self.position + torch.arange(n, device=self.states.device)) % self.capacity
        if n > self.capacity:           # only the last `capacity` items survive
            states, values, idx = states[-self.capacity:], values[-self.capacity:], idx[-self.capacity:]
        self.states[idx] = states.to(self.states.device)
        self.values[idx] = values.to(self.states.device).reshape(-1, 1).float()
        self.size = min(self.capacity, max(self.size, self.position + n))
        self.position = (self.position + n) % self.capacity

    def is_full(self):
        return self.size == self.capacity

    def __getitem__(self, item):
        return self.states[item], self.values[item]

    def __len__(self):
        return self.size

    def clear(self):
        self.size = 0
        self.position = 0


def average(input_list):
    return sum(input_list) / len(input_list) if len(input_list) else 0


class Explorer(object):
    def __init__(self, env, robot, device, memory=None, gamma=None, target_policy=None, dist_group=None):
        self.env = env
        self.robot = robot
        self.device = device
        self.memory = memory
        self.gamma = gamma
        self.target_policy = target_policy
        self.target_model = None
        self.dist_group = dist_group     # optional torch.distributed group: statistics are all-reduced over ranks
        self._batches = {}
        self._target_handle = None
        self.profile = False             # True: synchronise between the phases of run_k_episodes and keep seconds in last_timing
        self.last_timing = {}

    def update_target_model(self, target_model):
        """explorer.py:24-25: keep a frozen copy for TD targets; its weights also go to the FP32 CUDA network."""
        self.target_model = copy.deepcopy(target_model)
        self._target_handle = None

    # -- GPU plumbing ------------------------------------------------------------------------------
    def _batch(self, k, H):
        key = (k, H, self.robot.kinematics)                       # kinematics follows the robot's policy (ORCA -> SARL)
        if key not in self._batches:
            dev = getattr(self.device, "index", None) or 0     # torch.device('cuda') has index None -> GPU 0
            self._batches[key] = BatchedCrowdSim(k, H, device=dev, gamma=self.gamma or 0.9,
                                                 **batch_env_kwargs(self.env))
        return self._batches[key]

    def _target_forward(self, states):
        """V_target(s) for a (n, H, 13) device tensor via cn_policy_forward (FP32 CUDA kernel)."""
        policy = self.target_policy if isinstance(self.target_policy, SARL) else self.robot.policy
        if self._target_handle is None:
            h = policy.handle(self.robot.v_pref, precision="f32")
            # a separate handle so that the behaviour policy's weights are untouched
            from .batch import BatchedSARL
            self._target_handle = BatchedSARL(device=h.device, precision="f32", gamma=policy.gamma,
                                              v_pref=self.robot.v_pref, kinematics=h.cfg.kinematics, **policy._net_kwargs)
            self._target_handle.load_weights(self.target_model.state_dict())
        return self._target_handle.forward(states.contiguous())

    # -- the reference entry point (explorer.py:36-151) ----------------------------------------------
    def run_k_episodes(self, k, phase, update_memory=False, imitation_learning=False, episode=None,
                       print_failure=False, update_raw_ob=False, stay=False, returnRate=True, test_case=None,
                       returnNav=False, cacheFile=None):
        import torch
        env, robot, policy = self.env, self.robot, self.robot.policy
        policy.set_phase(phase)
        if update_raw_ob or cacheFile is not None:
            raise NotImplementedError("raw-observation / SGAN caches belong to the model-based branch (out of scope)")
        if robot.kinematics not in ("holonomic", "unicycle", None):
            raise NotImplementedError("robot kinematics must be holonomic, unicycle or None (the fork's literal behaviour)")
        import time
        t_start = time.perf_counter()
        cases = env.next_cases(phase, k, test_case)
        kw = env.scene_kwargs(phase)
        if kw.get("rule") == "mixed":                                       # ragged: every scene draws its own human count
            scene_list = [scenes.generate_scene(phase, int(c), **kw) for c in cases]
        else:
            scene_list = list(scenes.generate_batch(phase, cases, **kw))   # native generator for plain seeded scenes
        t_scenes = time.perf_counter()
        sizes = sorted(set(a.shape[0] for a in scene_list))
        if len(sizes) > 1:
            # 'mixed' scenes (crowd_sim.py:111-161) draw their own human count: one batch per count, merged in case order
            if update_memory:
                raise NotImplementedError("replay filling over scenes with different human counts (the reference's "
                                          "DataLoader cannot collate them either)")
            r = dict(final_info=np.zeros(k, np.int64), end_time=np.zeros(k), steps=np.zeros(k, np.int64), returns=np.zeros(k),
                     too_close=0, min_dist_sum=0.0, states_t=None, R=None, M=None)
            for n in sizes:
                idx = [i for i, a in enumerate(scene_list) if a.shape[0] == n]
                part = self._rollout(np.stack([scene_list[i] for i in idx]), phase, False, imitation_learning, stay)
                for key in ("final_info", "end_time", "steps", "returns"):
                    r[key][idx] = part[key]
                r["too_close"] += part["too_close"]; r["min_dist_sum"] += part["min_dist_sum"]
        else:
            r = self._rollout(np.stack(scene_list), phase, update_memory, imitation_learning, stay)
        if self.profile:
            torch.cuda.synchronize()
        t_roll = time.perf_counter()
        out = self._summarise(k, phase, cases, r, update_memory, imitation_learning, episode, print_failure, stay, returnRate,
                              returnNav)
        if self.profile:
            torch.cuda.synchronize()
            self.last_timing = {"scenes": t_scenes - t_start, "rollout": t_roll - t_scenes,
                                "targets_and_replay": time.perf_counter() - t_roll}
        return out

    def _rollout(self, agents, phase, update_memory, imitation_learning, stay):
        """k episodes with the same human count side by side (explorer.py:53-69 for every env of the batch)."""
        env, robot, policy = self.env, self.robot, self.robot.policy
        k = agents.shape[0]
        b = self._batch(k, agents.shape[1] - 1)
        b.set_state(agents, np.zeros(k))
        dt, v_pref = env.time_step, robot.v_pref
        robot.time_step = dt                      # CrowdSim.reset does this for every agent (crowd_sim.py:307-309)
        policy.time_step = dt
        is_sarl = isinstance(policy, SARL)
        if is_sarl:
            if policy.action_space is None:
                policy.build_action_space(v_pref)
            if policy.device is None:
                raise AttributeError("Phase, device attributes have to be set!")
            if phase == "train" and policy.epsilon is None:
                raise AttributeError("Epsilon attribute has to be set in training phase")
            handle = policy.handle(v_pref)
            eps = float(policy.epsilon) if phase == "train" else 0.0
        elif not isinstance(policy, (ORCA, Linear)) and not stay:
            raise NotImplementedError("robot policy %r is not on the B200 hot path" % type(policy).__name__)
        tr_policy = self.target_policy if (imitation_learning and self.target_policy is not None) else policy
        world_env = isinstance(env, ModelCrowdSim)

        max_steps = int(round(env.time_limit / dt)) + 2
        gamma = self.gamma if self.gamma is not None else 1.0
        if stay or is_sarl or isinstance(policy, ORCA):
            world = None
            if world_env:                              # ModelCrowdSim: the humans' next velocities come from the world model
                if not hasattr(env.sim_world, "handle"):
                    raise NotImplementedError("sim_world must be a modelcrowdnav_b200.world_model MlpWorld / AttentionWorld")
                world = env.sim_world.handle(b.device)
            record = None
            if update_memory:
                # policy.last_state = transform(state) (multi_human_rl.py:60-61) / target_policy.transform (explorer.py:163)
                if not isinstance(tr_policy, SARL):
                    raise ValueError("update_memory needs a SARL policy (or target_policy) to transform states")
                # RL: predict()'s last_state (LSTM-RL rows in its sorted human order); IL: plain transform()
                record = (tr_policy.handle(v_pref), not imitation_learning, isinstance(tr_policy, CADRL))
            return self._rollout_device(b, k, max_steps, stay, is_sarl, policy, handle if is_sarl else None,
                                        eps if is_sarl else 0.0, world, record)

        active = np.ones(k, bool)
        rewards_t, states_t, active_t = [], [], []
        final_info = np.zeros(k, np.int64)
        end_time = np.zeros(k)
        too_close, min_dist_sum = 0, 0.0
        for _ in range(max_steps):
            if not active.any():
                break
            if update_memory:
                # policy.last_state = transform(state) (multi_human_rl.py:60-61) / target_policy.transform (explorer.py:163)
                if not isinstance(tr_policy, SARL):
                    raise ValueError("update_memory needs a SARL policy (or target_policy) to transform states")
                # RL: predict()'s last_state (LSTM-RL rows in its sorted human order); IL: plain transform()
                st = tr_policy.handle(v_pref).transform(b, last_state=not imitation_learning)
                if isinstance(tr_policy, CADRL):
                    assert st.shape[1] == 1                        # cadrl.py:209: CADRL trains on single-human states
                    st = st[:, 0]
                states_t.append(st)
            if world_env:                  # ModelCrowdSim: the humans' next velocities come from the world model
                env.world_step_batch(b)
            else:
                b.orca()
            if stay:
                b.set_actions(np.zeros((k, 2)))
            elif is_sarl:
                handle.lookahead(b, query_env=policy.query_env, epsilon=eps)
                try:
                    handle.read(b, values=False)
                except _capi.CrowdNavError as e:
                    if e.code == _capi.CN_EVALUE:
                        raise ValueError("Value network is not well trained. ")
                    raise
            elif isinstance(policy, Linear):
                b.set_actions(Linear.batch_actions(b.get_state()[0]))
            else:
                b.robot_orca(policy.safety_space)
            reward, done, info, dmin = b.step(update=True)
            rewards_t.append(np.where(active, reward, 0.0))
            active_t.append(active.copy())
            danger = active & (info == _capi.DANGER)
            too_close += int(danger.sum())
            min_dist_sum += float(dmin[danger].sum())
            finished = active & (done != 0)
            if finished.any():
                final_info[finished] = info[finished]
                _, times = b.get_state()
                end_time[finished] = times[finished]
            active &= ~finished
        if active.any():
            raise ValueError("Invalid end signal from environment")
        R, M = np.stack(rewards_t), np.stack(active_t)
        disc = np.array([pow(gamma, t * dt * v_pref) for t in range(R.shape[0])])
        return dict(R=R, M=M, final_info=final_info, end_time=end_time, too_close=too_close, min_dist_sum=min_dist_sum,
                    states_t=states_t, returns=(R * disc[:, None]).sum(0), steps=M.sum(0))

    def _rollout_device(self, b, k, max_steps, stay, is_sarl, policy, handle, eps, world, record=None):
        """Episodes of a batch with NO host sync per step: the whole loop of explorer.py:53-69 is enqueued natively
        (cn_rollout_episodes), outcomes come from the per-env episode accumulators.  Finished envs freeze (auto_reset off), so
        extra steps are harmless; the count of running envs is read back asynchronously to stop soon after the last episode
        has ended.  record = (transform handle, last_state, squeeze): replay-filling phases also keep, per step, the transformed
        state and the reward / done outputs as DEVICE tensors."""
        import torch
        b.stats(reset=True)                             # zero the accumulators (set_state cleared the per-episode parts)
        if is_sarl:
            bad0 = handle.bad_count()
        if stay:
            b.set_actions(np.zeros((k, 2)))
        mode = _capi.ROBOT_KEEP if stay else (_capi.ROBOT_POLICY if is_sarl else _capi.ROBOT_ORCA)
        steps_run, S, R, D = b.run_episodes(
            max_steps, policy=handle if mode == _capi.ROBOT_POLICY else None, world=world, robot_mode=mode,
            safety_space=0.0 if is_sarl or stay else policy.safety_space, query_env=bool(is_sarl and policy.query_env), epsilon=eps,
            record=None if record is None else record[:2])
        states_t = None
        if record is not None:
            if record[2]:
                assert S.shape[2] == 1                             # cadrl.py:209: CADRL trains on single-human states
                S = S[:, :, 0]
            states_t = S
        t = b.episode_table()
        if is_sarl and handle.bad_count() != bad0:
            raise ValueError("Value network is not well trained. ")         # multi_human_rl.py:57-58
        if not t["frozen"].all() or not (t["episodes"] == 1).all():
            raise ValueError("Invalid end signal from environment")
        final_info = np.where(t["success"] == 1, _capi.REACHGOAL, np.where(t["collision"] == 1, _capi.COLLISION, _capi.TIMEOUT))
        end_time = np.where(t["success"] == 1, t["sum_success_time"],
                            np.where(t["collision"] == 1, t["sum_collision_time"], t["sum_timeout_time"]))
        M = None
        if record is not None:
            # env e was active in step t iff it had not finished before it: done stays 1 once an env is frozen
            M = torch.cat([torch.ones((1, k), dtype=torch.bool, device=D.device), D[:-1] == 0])
        return dict(R=R, M=M, final_info=final_info.astype(np.int64), end_time=end_time,
                    too_close=int(t["too_close"].sum()), min_dist_sum=float(t["sum_min_dist"].sum()), states_t=states_t,
                    returns=t["sum_return"].copy(), steps=t["steps"].copy())

    def _summarise(self, k, phase, cases, r, update_memory, imitation_learning, episode, print_failure, stay, returnRate,
                   returnNav):
        """Counters, log lines, replay filling and return values of explorer.py:70-151."""
        import torch
        env, robot = self.env, self.robot
        final_info, end_time, too_close = r["final_info"], r["end_time"], r["too_close"]
        R, M, states_t = r["R"], r["M"], r["states_t"]
        success = final_info == _capi.REACHGOAL
        collision = final_info == _capi.COLLISION
        timeout = final_info == _capi.TIMEOUT
        success_times = end_time[success].tolist()
        collision_times = end_time[collision].tolist()
        timeout_times = [env.time_limit] * int(timeout.sum())
        collision_cases = np.nonzero(collision)[0].tolist()
        timeout_cases = np.nonzero(timeout)[0].tolist()
        cumulative_rewards = r["returns"].tolist() if self.gamma is not None else \
            (R.sum(0).tolist() if R is not None else r["returns"].tolist())

        if update_memory:
            if self.memory is None or self.gamma is None:
                raise ValueError("Memory or gamma value is not set!")
            keep = np.nonzero(success | collision)[0]              # explorer.py:110-113
            if len(keep):
                S = torch.stack(states_t) if isinstance(states_t, list) else states_t     # (T, k, H, 13) on the device
                self._update_memory_batched(S, R, M, keep, imitation_learning)

        counts = np.array([success.sum(), collision.sum(), timeout.sum(), too_close, k], dtype=np.float64)
        sums = np.array([sum(success_times), sum(collision_times), sum(timeout_times), r["min_dist_sum"],
                         sum(cumulative_rewards)], dtype=np.float64)
        if self.dist_group is not None:
            counts, sums = self._all_reduce(counts, sums)
        n_success, n_collision, n_timeout, too_close, k_all = (int(x) for x in counts)
        success_rate, collision_rate = n_success / k_all, n_collision / k_all
        timeout_rate = (k_all - n_success - n_collision) / k_all
        assert n_success + n_collision + n_timeout == k_all
        avg_nav_time = sums[0] / n_success if n_success else env.time_limit
        avg_return = sums[4] / k_all

        extra_info = "" if episode is None else "in episode {} ".format(episode)
        if not stay:
            logging.info("{:<5} {}has success rate: {:.2f}, collision rate: {:.2f}, nav time: {:.2f}, total reward: {:.4f}".
                         format(phase.upper(), extra_info, success_rate, collision_rate, avg_nav_time, avg_return))
        if phase in ["val", "test"]:
            num_step = (sums[0] + sums[1] + sums[2]) / robot.time_step
            logging.info("Frequency of being in danger: %.2f and average min separate distance in danger: %.2f",
                         too_close / num_step, sums[3] / too_close if too_close else 0)
        if print_failure:
            logging.info("Collision cases: " + " ".join([str(x) for x in collision_cases]))
            logging.info("Timeout cases: " + " ".join([str(x) for x in timeout_cases]))
        self.last_run = dict(cases=cases, info=final_info, end_time=end_time, steps=np.asarray(r["steps"]), too_close=too_close,
                             returns=np.array(cumulative_rewards))
        if returnRate and returnNav:
            return avg_return, success_rate, collision_rate, timeout_rate, avg_nav_time
        if returnRate:
            return avg_return, success_rate, collision_rate, timeout_rate
        return avg_return, n_success, n_collision, (k_all - n_success - n_collision)

    def _all_reduce(self, counts, sums):
        """Episode statistics summed over ranks: one small all-reduce per run_k_episodes."""
        import torch
        import torch.distributed as dist
        dev = self.device if dist.get_backend(self.dist_group) == "nccl" else "cpu"
        t = torch.tensor(np.concatenate([counts, sums]), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.dist_group)
        t = t.cpu().numpy()
        return t[:len(counts)], t[len(counts):]

    def _update_memory_batched(self, S, R, M, keep, imitation_learning):
        """explorer.py:153-186 for every kept episode at once; pushes in episode-then-time order.  Everything stays on the
        device: the IL return-to-go is a reversed cumulative sum, the RL targets one batched target-network forward, and the
        (episode, time) pairs are gathered with ONE index and pushed with ONE ring-buffer write."""
        import torch
        T = R.shape[0]
        dev = S.device
        gamma_bar = pow(self.gamma, self.robot.time_step * self.robot.v_pref)
        keep_t = torch.as_tensor(np.asarray(keep), device=dev)
        Rk = torch.as_tensor(R, dtype=torch.float64, device=dev)[:, keep_t]        # (T, n); zero after an episode's end
        Mk = torch.as_tensor(M, device=dev)[:, keep_t]
        if imitation_learning:
            # value_i = sum_{t >= i} gamma_bar^(t - i) r_t (explorer.py:161-166) = revcumsum(r_t gamma_bar^t)_i / gamma_bar^i
            wgt = torch.pow(torch.tensor(gamma_bar, dtype=torch.float64, device=dev),
                            torch.arange(T, dtype=torch.float64, device=dev)).unsqueeze(1)
            V = torch.flip(torch.cumsum(torch.flip(Rk * wgt, [0]), 0), [0]) / wgt
        else:
            # value_i = r_i + gamma_bar * V_target(s_{i+1}); terminal: r (explorer.py:168-174)
            Sk = S[:, keep]                                                        # (T, n, H, 13)
            nxt = torch.zeros((T,) + tuple(Rk.shape[1:]), dtype=torch.float64, device=dev)
            if T > 1:
                flat = Sk[1:].reshape((-1,) + tuple(Sk.shape[2:]))
                nxt[:-1] = self._target_forward(flat).reshape(T - 1, -1).double()
            last = Mk & ~torch.cat([Mk[1:], torch.zeros_like(Mk[:1])])
            V = torch.where(last, Rk, Rk + gamma_bar * nxt)
        jj, tt = torch.nonzero(Mk.t(), as_tuple=True)                              # episode-major, time-minor
        ee = keep_t[jj]
        self.memory.push_batch(S[tt, ee], V[tt, jj].float())
