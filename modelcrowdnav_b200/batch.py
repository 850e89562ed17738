"""Batched host API over the C ABI: E environments per GPU, one SARL policy.

This is the layer the drop-in façades (``crowd_sim.CrowdSim``, ``policy.SARL``, ``explorer.Explorer``)
and ``bench.py`` are built on.  Host arrays are numpy; device tensors (torch) are only used for hand-off
(`transform`, `forward`).  Layout of the exchange format: ``agents[E, H+1, 8]`` float64 with columns
``px py vx vy gx gy radius v_pref`` and agent 0 = robot (crowd_sim/envs/utils/state.py FullState order
minus theta, which is unused for holonomic kinematics).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _stream(stream):
    return C.c_void_p(int(stream)) if stream else None


class BatchedCrowdSim(object):
    """E CrowdSim environments in HBM (replaces CrowdSim, crowd_sim/envs/crowd_sim.py:15-434)."""

    def __init__(self, num_envs, human_num=5, device=0, **cfg):
        self.lib = _capi.load()
        self.cfg = _capi.default_env_cfg(num_envs=num_envs, human_num=human_num, **cfg)
        self.E, self.H, self.device = num_envs, human_num, device
        self.handle = C.c_void_p()
        check(self.lib.cn_env_create(C.byref(self.cfg), device, C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cn_env_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ------------------------------------------------------------------------------------
    def set_theta(self, theta, stream=None):
        """Robot headings (E,), only meaningful for robot_kinematics != holonomic; resets put pi / 2."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        assert theta.shape == (self.E,)
        check(self.lib.cn_env_set_theta(self.handle, _ptr(theta), _stream(stream)))

    def get_theta(self, stream=None):
        theta = np.empty(self.E, np.float64)
        check(self.lib.cn_env_get_theta(self.handle, _ptr(theta), _stream(stream)))
        return theta

    def set_state(self, agents, times=None, stream=None):
        agents = np.ascontiguousarray(agents, dtype=np.float64)
        assert agents.shape == (self.E, self.H + 1, _capi.AGENT_STRIDE), agents.shape
        if times is not None:
            times = np.ascontiguousarray(times, dtype=np.float64)
            assert times.shape == (self.E,)
        check(self.lib.cn_env_set_state(self.handle, _ptr(agents), _ptr(times), _stream(stream)))

    def get_state(self, stream=None):
        agents = np.empty((self.E, self.H + 1, _capi.AGENT_STRIDE), np.float64)
        times = np.empty(self.E, np.float64)
        check(self.lib.cn_env_get_state(self.handle, _ptr(agents), _ptr(times), _stream(stream)))
        return agents, times

    def reset_device(self, stream=None):
        """CrowdSim.reset on the GPU (Philox scenes; crowd_sim.py:165-217 distributions)."""
        check(self.lib.cn_env_reset(self.handle, _stream(stream)))

    # -- ORCA / step ------------------------------------------------------------------------------
    def orca(self, stream=None):
        check(self.lib.cn_env_orca(self.handle, _stream(stream)))

    def human_actions(self, stream=None):
        out = np.empty((self.E, self.H, 2), np.float64)
        check(self.lib.cn_env_read_human_actions(self.handle, _ptr(out), _stream(stream)))
        return out

    def set_human_actions(self, human_vxy, stream=None):
        """(E, H, 2) human velocities for the next step instead of an ORCA solve (a world model's prediction,
        model_crowd_sim.py:397-425)."""
        human_vxy = np.ascontiguousarray(human_vxy, dtype=np.float64)
        assert human_vxy.shape == (self.E, self.H, 2)
        check(self.lib.cn_env_set_human_actions(self.handle, _ptr(human_vxy), _stream(stream)))

    def robot_orca(self, safety_space=0.0, stream=None):
        check(self.lib.cn_env_robot_orca(self.handle, float(safety_space), _stream(stream)))

    def pending_actions(self, stream=None):
        """(E,2) action xy and (E,) table index of the action chosen by the lookahead / robot ORCA."""
        xy = np.empty((self.E, 2), np.float64)
        idx = np.empty(self.E, np.int32)
        check(self.lib.cn_env_read_actions(self.handle, _ptr(xy), _ptr(idx), _stream(stream)))
        return xy, idx

    def set_actions(self, actions, stream=None):
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        assert actions.shape == (self.E, 2)
        check(self.lib.cn_env_set_actions(self.handle, _ptr(actions), _stream(stream)))

    def step(self, actions=None, update=True, read=True, stream=None):
        """CrowdSim.step for every env. actions: (E,2) host array or None (= pending action)."""
        if actions is not None:
            self.set_actions(actions, stream)
        check(self.lib.cn_env_step(self.handle, None, int(bool(update)), _stream(stream)))
        return self.read_outputs(stream) if read else None

    def read_outputs(self, stream=None):
        reward = np.empty(self.E, np.float64)
        done = np.empty(self.E, np.uint8)
        info = np.empty(self.E, np.uint8)
        dmin = np.empty(self.E, np.float64)
        check(self.lib.cn_env_read_outputs(self.handle, _ptr(reward), _ptr(done), _ptr(info), _ptr(dmin),
                                           _stream(stream)))
        return reward, done, info, dmin

    def next_obs(self, stream=None):
        out = np.empty((self.E, self.H, 5), np.float64)
        check(self.lib.cn_env_read_next_obs(self.handle, _ptr(out), _stream(stream)))
        return out

    def views(self):
        v = _capi.EnvViews()
        check(self.lib.cn_env_get_views(self.handle, C.byref(v)))
        return v

    def copy_outputs_to(self, reward=None, done=None, info=None, stream=None):
        """reward (E,) float64 / done, info (E,) uint8 CUDA tensors <- outputs of the last step, device to device, async."""
        def p(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        check(self.lib.cn_env_copy_outputs(self.handle, p(reward), p(done), p(info), _stream(stream)))

    def run_episodes(self, max_steps, policy=None, world=None, robot_mode=_capi.ROBOT_POLICY, safety_space=0.0,
                     query_env=False, epsilon=0.0, check_every=8, record=None, stream=None):
        """One episode per env, enqueued natively (cn_rollout_episodes; explorer.py:53-69 for the whole batch).
        record = (transform BatchedSARL, last_state flag): also returns the device records (states (T, E, H, F) fp32,
        reward (T, E) f64, done (T, E) u8), T = steps enqueued.  Returns (steps_run, states, reward, done)."""
        rec_p, S, R, D = None, None, None, None
        if record is not None:
            import torch
            th, last_state = record
            dev = torch.device("cuda", self.device)
            S = torch.empty((max_steps, self.E, self.H, th.cfg.input_dim), dtype=torch.float32, device=dev)
            R = torch.zeros((max_steps, self.E), dtype=torch.float64, device=dev)
            D = torch.ones((max_steps, self.E), dtype=torch.uint8, device=dev)
            rec = _capi.RolloutRecord(th.handle, int(bool(last_state)), S.data_ptr(), R.data_ptr(), D.data_ptr())
            rec_p = C.byref(rec)
        n = C.c_int32(0)
        check(self.lib.cn_rollout_episodes(policy.handle if policy is not None else None, self.handle, world, int(robot_mode),
                                           float(safety_space), int(bool(query_env)), float(epsilon), int(max_steps),
                                           int(check_every), rec_p, C.byref(n), _stream(stream)))
        n = n.value
        if record is not None:
            S, R, D = S[:n], R[:n], D[:n]
        return n, S, R, D

    def episode_table(self, stream=None):
        """Per-env episode accumulators (un-reduced cn_env_read_stats) + frozen flags: dict of (E,) arrays.  Blocking."""
        E = self.E
        raw = np.empty(E * 11, np.int64)
        frozen = np.empty(E, np.uint8)
        check(self.lib.cn_env_read_episode_table(self.handle, _ptr(raw), _ptr(frozen), _stream(stream)))
        names_i = ("episodes", "success", "collision", "timeout", "steps", "too_close")
        names_d = ("sum_min_dist", "sum_success_time", "sum_collision_time", "sum_timeout_time", "sum_return")
        out = {k: raw[i * E:(i + 1) * E] for i, k in enumerate(names_i)}
        out.update({k: raw[(6 + i) * E:(7 + i) * E].view(np.float64) for i, k in enumerate(names_d)})
        out["frozen"] = frozen
        return out

    def all_done(self, stream=None):
        frozen = np.empty(self.E, np.uint8)
        check(self.lib.cn_env_read_episode_table(self.handle, None, _ptr(frozen), _stream(stream)))
        return bool(frozen.all())

    def stats(self, reset=False, stream=None):
        s = _capi.Stats()
        check(self.lib.cn_env_read_stats(self.handle, C.byref(s), int(reset), _stream(stream)))
        return {k: getattr(s, k) for k, _ in _capi.Stats._fields_}


class BatchedSARL(object):
    """SARL lookahead policy on the GPU (replaces SARL/MultiHumanRL.predict, sarl.py:68-89,
    multi_human_rl.py:11-63)."""

    def __init__(self, device=0, precision="f32", **cfg):
        self.lib = _capi.load()
        prec = {"f32": _capi.PREC_F32, "f16_tc": _capi.PREC_F16_TC}[precision]
        if isinstance(cfg.get("network"), str):      # "sarl" | "cadrl" | "lstm_rl" (policy_factory names)
            cfg["network"] = {"sarl": _capi.NET_SARL, "cadrl": _capi.NET_CADRL, "lstm_rl": _capi.NET_LSTM_RL}[cfg["network"]]
        self.cfg = _capi.default_sarl_cfg(precision=prec, **cfg)
        self.device = device
        self.handle = C.c_void_p()
        check(self.lib.cn_policy_create(C.byref(self.cfg), device, C.byref(self.handle)))
        self.n_params = int(self.lib.cn_policy_param_count(C.byref(self.cfg)))
        n = C.c_int32()
        check(self.lib.cn_policy_action_table(self.handle, None, C.byref(n)))
        self.action_table = np.empty((n.value, 2), np.float64)
        check(self.lib.cn_policy_action_table(self.handle, _ptr(self.action_table), C.byref(n)))
        self.A = n.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cn_policy_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weights(self, weights, stream=None):
        """weights: flat fp32 array in state-dict order, or a ValueNetwork state_dict (test.py:59)."""
        if isinstance(weights, dict):
            weights = np.concatenate([np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v,
                                                 dtype=np.float32).ravel() for v in weights.values()])
        weights = np.ascontiguousarray(weights, dtype=np.float32)
        check(self.lib.cn_policy_load_weights(self.handle, _ptr(weights), weights.size, _stream(stream)))

    def lookahead(self, env, query_env=False, epsilon=0.0, stream=None):
        check(self.lib.cn_policy_lookahead(self.handle, env.handle, int(bool(query_env)), float(epsilon),
                                           _stream(stream)))

    def read(self, env, values=True, stream=None):
        best = np.empty(env.E, np.int32)
        vals = np.empty((env.E, self.A), np.float64) if values else None
        check(self.lib.cn_policy_read(self.handle, env.handle, _ptr(best), _ptr(vals), _stream(stream)))
        return best, vals

    def bad_count(self, reset=False, stream=None):
        """Envs whose action values were all non-finite so far (the reference's ValueError, multi_human_rl.py:57-58): the
        device-resident / async rollout forms substitute action 0 and count here instead of raising."""
        n = C.c_int64()
        check(self.lib.cn_policy_bad_count(self.handle, C.byref(n), int(bool(reset)), _stream(stream)))
        return int(n.value)

    def kernel_timing(self, on=True):
        """Measurement hook: CUDA events around the kernels of every later tensor-core lookahead (bench.py roofline)."""
        check(self.lib.cn_debug_kernel_ms(self.handle, int(bool(on)), None))

    def kernel_ms(self):
        """ms of {features, rows, mlp3, argmax} kernels of the LAST lookahead (after kernel_timing(True)); blocks."""
        out = (C.c_float * 4)()
        check(self.lib.cn_debug_kernel_ms(self.handle, 1, out))
        return dict(zip(("features", "rows", "mlp3", "argmax"), (float(x) for x in out)))

    def transform(self, env, stream=None, last_state=False):
        """MultiHumanRL.transform for every env -> torch CUDA tensor (E, H, 13) fp32.  last_state=True: what predict()
        leaves in policy.last_state (LSTM-RL: rows in its sorted human order, lstm_rl.py:99-104)."""
        import torch
        out = torch.empty((env.E, env.H, self.cfg.input_dim), dtype=torch.float32, device="cuda:%d" % self.device)
        fn = self.lib.cn_policy_last_state if last_state else self.lib.cn_policy_transform
        check(fn(self.handle, env.handle, C.c_void_p(out.data_ptr()), _stream(stream)))
        return out

    def forward(self, x, stream=None):
        """ValueNetwork.forward on a CUDA tensor (B, H, 13) fp32 -> (B,) fp32 (FP32 kernel)."""
        import torch
        if x.dim() == 2:                              # CADRL trains on single (robot, human) rows (cadrl.py:202-216)
            x = x.unsqueeze(1)
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[2] == self.cfg.input_dim
        x = x.contiguous()
        out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        check(self.lib.cn_policy_forward(self.handle, C.c_void_p(x.data_ptr()), x.shape[0], x.shape[1],
                                         C.c_void_p(out.data_ptr()), _stream(stream)))
        return out


def rollout_step(policy, env, query_env=False, epsilon=0.0, stream=None):
    """orca -> lookahead -> step(update=True) [-> auto reset] without a host round trip."""
    check(policy.lib.cn_rollout_step(policy.handle, env.handle, int(bool(query_env)), float(epsilon),
                                     _stream(stream)))


class HostStepBuffers(object):
    """Pinned host buffers for ``rollout_step_host`` (the end-to-end path with HOST state)."""

    def __init__(self, env, pinned=True):
        E, A1 = env.E, env.H + 1
        shapes = dict(agents_in=((E, A1, 8), np.float64), times_in=((E,), np.float64),
                      agents_out=((E, A1, 8), np.float64), times_out=((E,), np.float64),
                      reward=((E,), np.float64), done=((E,), np.uint8), info=((E,), np.uint8),
                      action_idx=((E,), np.int32))
        self._keep = []
        for k, (shape, dt) in shapes.items():
            if pinned:
                import torch
                t = torch.empty(shape, dtype=getattr(torch, np.dtype(dt).name), pin_memory=True)
                self._keep.append(t)
                setattr(self, k, t.numpy())
            else:
                setattr(self, k, np.empty(shape, dt))
        self.h2d_bytes = self.agents_in.nbytes + self.times_in.nbytes
        self.d2h_bytes = (self.agents_out.nbytes + self.times_out.nbytes + self.reward.nbytes + self.done.nbytes
                          + self.info.nbytes + self.action_idx.nbytes)


    def swap(self):
        """Make the state just downloaded the next call's input without copying it (ping-pong of the two
        pinned buffer pairs): what a host-resident simulation loop does between two steps."""
        self.agents_in, self.agents_out = self.agents_out, self.agents_in
        self.times_in, self.times_out = self.times_out, self.times_in


class PackedHostStepBuffers(object):
    """Two pinned packed exchange blocks for ``rollout_step_host_packed`` (one H2D + one D2H copy per step).
    ``agents_in / times_in`` view the input block; ``agents_out, times_out, reward, action_idx, done, info`` view the
    output block; ``swap()`` makes the state just downloaded the next input without copying it."""

    def __init__(self, env):
        import torch
        self.E, self.A1 = env.E, env.H + 1
        self.in_bytes = int(env.lib.cn_host_step_bytes(env.handle, 0))
        self.out_bytes = int(env.lib.cn_host_step_bytes(env.handle, 1))
        nb = (self.out_bytes + 15) // 16 * 16
        self._blocks = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        self._views(0, 1)
        self.h2d_bytes, self.d2h_bytes = self.in_bytes, self.out_bytes

    def _views(self, i, o):
        E, A1 = self.E, self.A1
        self._in, self._out = i, o
        bi, bo = self._blocks[i].numpy(), self._blocks[o].numpy()
        n = E * A1 * 8
        self.agents_in = bi[:8 * n].view(np.float64).reshape(E, A1, 8)
        self.times_in = bi[8 * n:8 * (n + E)].view(np.float64)
        self.agents_out = bo[:8 * n].view(np.float64).reshape(E, A1, 8)
        self.times_out = bo[8 * n:8 * (n + E)].view(np.float64)
        off = 8 * (n + E)
        self.reward = bo[off:off + 8 * E].view(np.float64); off += 8 * E
        self.action_idx = bo[off:off + 4 * E].view(np.int32); off += 4 * E
        self.done = bo[off:off + E]; off += E
        self.info = bo[off:off + E]
        self.in_ptr, self.out_ptr = self._blocks[i].data_ptr(), self._blocks[o].data_ptr()

    def swap(self):
        self._views(self._out, self._in)


def rollout_step_host_packed(policy, env, buf, query_env=False, epsilon=0.0, stream=None):
    """One lookahead + env step through the packed host blocks: ONE H2D copy, kernels, ONE D2H copy (blocking)."""
    check(policy.lib.cn_rollout_step_host_packed(policy.handle, env.handle, int(bool(query_env)), float(epsilon),
                                                 C.c_void_p(buf.in_ptr), C.c_void_p(buf.out_ptr), _stream(stream)))


def rollout_step_host(policy, env, buf, query_env=False, epsilon=0.0, stream=None):
    """One lookahead + env step through host buffers: H2D state, kernels, D2H results (blocking)."""
    check(policy.lib.cn_rollout_step_host(policy.handle, env.handle, int(bool(query_env)), float(epsilon),
                                          _ptr(buf.agents_in), _ptr(buf.times_in), _ptr(buf.agents_out),
                                          _ptr(buf.times_out), _ptr(buf.reward), _ptr(buf.done), _ptr(buf.info),
                                          _ptr(buf.action_idx), _stream(stream)))


class PipelinedHostRollout(object):
    """Host-resident rollout of E envs split into `shards` env handles, each with its own stream, policy handle
    and pair of pinned packed blocks.  ``step()`` advances every env once: while one shard's kernels run, the other
    shards' host->device / device->host copies are in flight (cn_rollout_step_host_packed_async).  Per step every
    env's state is uploaded from and downloaded to host memory, exactly like ``rollout_step_host_packed``; only the
    order in which the shards are served overlaps.  Results do not depend on the shard count (envs are independent
    and keyed by their global id).  Measured on B200 (8192 envs x 5 humans): 1 handle blocking 8.8e6, 2 shards 9.3e6,
    4 shards 1.05e7 env-steps/s -- with 4 shards there is always a queued row kernel, so each shard's copies, host
    round trip, feature kernel and small kernels (which fit beside the persistent row kernel) are fully hidden."""

    def __init__(self, num_envs, human_num, weights, device=0, shards=4, precision="f16_tc", env_id_offset=0,
                 policy_cfg=None, use_graphs=True, **env_cfg):
        import torch
        assert num_envs % shards == 0
        self.E, self.shards, self.device = num_envs, shards, device
        self.Es = num_envs // shards
        self.envs, self.pols, self.bufs, self.streams = [], [], [], []
        for k in range(shards):
            env = BatchedCrowdSim(self.Es, human_num, device=device, env_id_offset=env_id_offset + k * self.Es, **env_cfg)
            pol = BatchedSARL(device=device, precision=precision, **(policy_cfg or {}))
            pol.load_weights(weights)
            self.envs.append(env); self.pols.append(pol)
            self.bufs.append(PackedHostStepBuffers(env))
            self.streams.append(torch.cuda.Stream(device=device))
        self.lib = self.envs[0].lib
        self.use_graphs = bool(use_graphs)
        self._graphs = {}                    # (shard, input block index, query_env, epsilon) -> torch.cuda.CUDAGraph
        self._eager_steps = [0] * shards
        self._busy = [False] * shards
        self.h2d_bytes = sum(b.h2d_bytes for b in self.bufs)
        self.d2h_bytes = sum(b.d2h_bytes for b in self.bufs)

    def reset_device(self):
        """CrowdSim.reset of every env on the GPU (scenes keyed by the global env id) -> host input blocks."""
        for env, b in zip(self.envs, self.bufs):
            env.reset_device()
            a, t = env.get_state()
            b.agents_in[...] = a; b.times_in[...] = t

    def load_state(self, agents, times):
        """Host state of all E envs -> the shards' input blocks."""
        for k, b in enumerate(self.bufs):
            sl = slice(k * self.Es, (k + 1) * self.Es)
            b.agents_in[...] = agents[sl]; b.times_in[...] = times[sl]

    def _wait(self, k):
        if self._busy[k]:
            check(self.lib.cn_stream_sync(self.device, C.c_void_p(self.streams[k].cuda_stream)))
            self._busy[k] = False
            self.bufs[k].swap()              # the state just downloaded is the next input (no host copy)

    def _enqueue(self, k, query_env, epsilon, stream_ptr):
        b = self.bufs[k]
        check(self.lib.cn_rollout_step_host_packed_async(
            self.pols[k].handle, self.envs[k].handle, int(bool(query_env)), float(epsilon),
            C.c_void_p(b.in_ptr), C.c_void_p(b.out_ptr), C.c_void_p(stream_ptr)))

    def step(self, query_env=False, epsilon=0.0):
        """Enqueue one step of every shard; a shard is re-launched as soon as its previous step has landed.
        With use_graphs, a shard's whole step (H2D copy, ~10 kernels on three streams, D2H copy) is captured once per
        block orientation into a CUDA graph and replayed: one launch per shard and step instead of ~25 driver calls,
        which keeps the host loop off the critical path on slow or busy hosts."""
        import torch
        for k in range(self.shards):
            self._wait(k)
            st = self.streams[k]
            if not self.use_graphs or self._eager_steps[k] < 2:          # the first steps allocate workspaces: run eagerly
                self._enqueue(k, query_env, epsilon, st.cuda_stream)
                self._eager_steps[k] += 1
            else:
                key = (k, self.bufs[k]._in, bool(query_env), float(epsilon))
                g = self._graphs.get(key)
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    st.synchronize()
                    with torch.cuda.graph(g, stream=st):
                        self._enqueue(k, query_env, epsilon, torch.cuda.current_stream().cuda_stream)
                    self._graphs[key] = g
                with torch.cuda.stream(st):
                    g.replay()
            self._busy[k] = True

    def step_device(self, query_env=False, epsilon=0.0):
        """One step of every shard with the state left in HBM (cn_rollout_step_sharded, no copies, no host sync): the
        device-resident counterpart of ``step()``; read the state back with ``envs[k].get_state()`` after ``sync()``."""
        for k in range(self.shards):
            check(self.lib.cn_rollout_step_sharded(self.pols[k].handle, self.envs[k].handle, int(bool(query_env)),
                                                   float(epsilon), C.c_void_p(self.streams[k].cuda_stream)))

    def sync_device(self):
        for k in range(self.shards):
            check(self.lib.cn_stream_sync(self.device, C.c_void_p(self.streams[k].cuda_stream)))

    def sync(self):
        """Wait for every shard; afterwards ``results()`` is valid."""
        for k in range(self.shards):
            self._wait(k)

    def results(self):
        """(agents, times, reward, action_idx, done, info) of the last completed step, concatenated over shards.
        After ``sync()`` the downloaded blocks are the *input* side of the ping-pong."""
        outs = []
        for b in self.bufs:
            E, A1 = b.E, b.A1
            blk = b._blocks[b._in].numpy()
            n = E * A1 * 8
            off = 8 * (n + E)
            outs.append((blk[:8 * n].view(np.float64).reshape(E, A1, 8), blk[8 * n:off].view(np.float64),
                         blk[off:off + 8 * E].view(np.float64), blk[off + 8 * E:off + 12 * E].view(np.int32),
                         blk[off + 12 * E:off + 13 * E], blk[off + 13 * E:off + 14 * E]))
        return tuple(np.concatenate(x) for x in zip(*outs))

    def close(self):
        self.sync()
        self._graphs.clear()
        for e, p in zip(self.envs, self.pols):
            e.close(); p.close()


def pin_to_gpu_numa(device_index):
    """Restrict this process to the CPUs NVML reports as local to the GPU (its NUMA node).  A host-resident driver moves
    6.5 MB per step and GPU through pinned memory; with every rank of a node on the same CPU set the far-socket ranks lose ~2 %
    (8-GPU end-to-end line of round 1).  Returns the number of CPUs kept, or 0 when nothing was changed (no NVML, no overlap
    with the allowed set, single-socket box)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        keep = local & allowed
        if not keep or keep == allowed:
            return 0
        os.sched_setaffinity(0, keep)
        return len(keep)
    except Exception:
        return 0
