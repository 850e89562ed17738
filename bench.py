#!/usr/bin/env python
"""bench.py -- env-steps/s of the hot path (ORCA step + SARL lookahead) on N B200s of one node.

A "step" = one cn_rollout_step over the whole batch resident on each GPU: ORCA for every human, the
81-action SARL lookahead, CrowdSim.step(update=True) and the auto-reset of finished episodes.
Workload at any N: BASELINE.json configs[1] per GPU -- 8192 batched envs x 5 humans (weak scaling; envs are
sharded by contiguous global id, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--precision f32|f16_tc] [--envs 8192] [--humans 5] [--query-env 0|1] [--sim circle|square]

--impl reference times the reference's CPU algorithm (the oracle port, oracle/crowdnav_oracle.c -- the
reference itself is Python + an un-vendored rvo2 and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/s (ORCA step + SARL lookahead)"
UNIT = "env-steps/s"


# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
NCU_DRAM_BYTES = {("f16_tc", 8192, 5): 218508032 + 56814592}   # tc_rows_pair_kernel<5>: X tiles in, J tiles out


def flops_per_env_step(H):
    """literal ValueNetwork formulation, 2 FLOP/MAC, 81 actions."""
    return 81 * (124100 * H + 67000)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_burst=d["bf16_tflops"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm_gbs=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(E, H, sim, query_env, n_threads, target_s, max_steps=None):
    """Oracle port (C restatement of the reference algorithm) on host cores: env-steps/s over a bounded sample."""
    import oracle
    oracle.build()
    ecfg, scfg = oracle.EnvCfg.default(), oracle.SarlCfg.default()
    rule = "circle_crossing" if sim == "circle" else "square_crossing"
    agents = np.stack([oracle.generate_scene("train", c, human_num=H, rule=rule) for c in range(E)])
    w = oracle.default_sarl_weights(0)
    table = oracle.action_space()
    times = np.zeros(E); done = np.zeros(E, np.uint8)
    oracle.batch_lookahead_step(ecfg, scfg, w, agents, times, table, query_env, done, n_threads=n_threads)  # warm
    t0 = time.perf_counter(); steps = 0
    while True:
        done[:] = 0
        oracle.batch_lookahead_step(ecfg, scfg, w, agents, times, table, query_env, done, n_threads=n_threads)
        steps += 1
        el = time.perf_counter() - t0
        if el >= target_s or (max_steps and steps >= max_steps):
            break
    return steps * E / el, steps, el


def run_reference(a):
    """Reference arm: the reference's CPU implementation of the path = oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    E = min(a.envs, 64 * cores)
    rate1, _, _ = cpu_port_rate(E, a.humans, a.sim, a.query_env, cores, 1.0)     # size K steps to a few minutes
    per_step = E / rate1
    t_all = []
    import oracle
    ecfg, scfg = oracle.EnvCfg.default(), oracle.SarlCfg.default()
    rule = "circle_crossing" if a.sim == "circle" else "square_crossing"
    agents = np.stack([oracle.generate_scene("train", c, human_num=a.humans, rule=rule) for c in range(E)])
    w = oracle.default_sarl_weights(0); table = oracle.action_space()
    times = np.zeros(E); done = np.zeros(E, np.uint8)
    for i in range(a.warmup + a.steps):
        done[:] = 0
        t0 = time.perf_counter()
        oracle.batch_lookahead_step(ecfg, scfg, w, agents, times, table, a.query_env, done, n_threads=cores)
        if i >= a.warmup:
            t_all.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(t_all))
    value = E / (ms / 1e3)
    sample = "%d envs x %d humans per step (bounded sample of the %d-env workload), %d OpenMP threads" % (
        E, a.humans, a.envs, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "envs_per_step": E, "humans": a.humans, "query_env": a.query_env,
                       "est_ms_per_step": 1e3 * per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_name(a):
    return "SARL lookahead + ORCA step, %d batched envs x %d humans per GPU, %s_crossing, query_env=%d" % (
        a.envs, a.humans, a.sim, a.query_env)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CN_BENCH_PRECISION", "f16_tc"), choices=["f32", "f16_tc"])
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--humans", type=int, default=5)
    ap.add_argument("--query-env", type=int, default=0)
    ap.add_argument("--sim", default="circle", choices=["circle", "square"])
    ap.add_argument("--e2e-shards", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import modelcrowdnav_b200 as mcn
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    E, H = a.envs, a.humans
    env = mcn.BatchedCrowdSim(E, H, device=local, auto_reset=1, seed=0, env_id_offset=rank * E,
                              sim_rule=0 if a.sim == "circle" else 1)
    pol = mcn.BatchedSARL(device=local, precision=a.precision)
    # SARL weights: default nn.Linear init under torch.manual_seed(0) (no trained weights ship with the reference)
    wpath = os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy")
    pol.load_weights(np.load(wpath))
    env.reset_device()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    # ---- warm-up, then K timed steps (device time per step, L2 flushed between steps) ----
    for _ in range(a.warmup):
        mcn.rollout_step(pol, env, a.query_env)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    launches0 = pol.lib.cn_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(a.steps):
        flush.zero_()
        ev[i][0].record()
        mcn.rollout_step(pol, env, a.query_env)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = pol.lib.cn_launch_count() - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel breakdown of the same step (events around each phase) ----
    phases = {"orca": 0.0, "lookahead": 0.0, "step": 0.0}
    nb = min(a.steps, 20)
    for _ in range(nb):
        flush.zero_()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        marks[0].record(); env.orca()
        marks[1].record(); pol.lookahead(env, a.query_env)
        marks[2].record(); env.step(update=True, read=False)
        marks[3].record()
        torch.cuda.synchronize()
        for k, name in enumerate(("orca", "lookahead", "step")):
            phases[name] += marks[k].elapsed_time(marks[k + 1]) / nb
    barrier()

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    # Every step, every env's state is uploaded from a pinned host block and its new state + reward + done + info +
    # action are downloaded (packed blocks: one H2D and one D2H copy per shard and step).  The E envs are served as
    # `e2e_shards` env handles on their own streams (cn_rollout_step_host_packed_async), so that the PCIe copies of
    # one shard overlap the kernels of the other; the blocking single-handle call is timed next to it.
    ne = max(5, min(a.steps, 50))
    buf = mcn.PackedHostStepBuffers(env)
    a0, t0 = env.get_state()
    buf.agents_in[...] = a0; buf.times_in[...] = t0
    for _ in range(2):
        mcn.rollout_step_host_packed(pol, env, buf, a.query_env)
        buf.swap()
    barrier()
    te0 = time.perf_counter()
    for _ in range(ne):
        mcn.rollout_step_host_packed(pol, env, buf, a.query_env)
        buf.swap()                                   # next step's input = the host state just downloaded (no host copy)
    barrier()
    e2e_blocking_s = time.perf_counter() - te0
    pipe = mcn.PipelinedHostRollout(E, H, np.load(wpath), device=local, shards=a.e2e_shards, precision=a.precision,
                                    env_id_offset=rank * E, auto_reset=1, seed=0,
                                    sim_rule=0 if a.sim == "circle" else 1)
    pipe.reset_device()
    for _ in range(6):                               # 2 eager steps per shard, then one graph capture per block orientation
        pipe.step(a.query_env)
    pipe.sync()
    barrier()
    te0 = time.perf_counter()
    for _ in range(ne):
        pipe.step(a.query_env)
    pipe.sync()
    barrier()
    e2e_s = time.perf_counter() - te0
    h2d_bytes, d2h_bytes = pipe.h2d_bytes, pipe.d2h_bytes
    pipe.close()

    t = torch.tensor([dev_ms, e2e_s, t_wall, e2e_blocking_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, t_wall, e2e_blocking_s = t.tolist()
    st = env.stats()
    if rank == 0:
        peaks = load_peaks()
        ms_per_step = dev_ms / a.steps
        value = world * E / (ms_per_step / 1e3)
        F = flops_per_env_step(H)
        la_ms = phases["lookahead"]
        ach = E * F / (la_ms / 1e3) / 1e12
        peak = peaks["tflops_sustained"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if a.precision == "f16_tc" else "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "envs_per_gpu": E, "humans": H, "actions": pol.A,
                       "query_env": a.query_env, "precision": a.precision,
                       "weights": "torch.manual_seed(0) default nn.Linear init",
                       "l2": "256 MiB memset between timed steps (outside the event pair)",
                       "wall_s_timed_region": t_wall, "phase_ms": phases,
                       "episodes_finished": st["episodes"]},
            "clocks": clocks,
            "e2e": {"value": world * E * ne / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": ne, "shards_per_gpu": a.e2e_shards,
                    "how": "PipelinedHostRollout: pinned host state in and out every step, the copies, host round trip and "
                           "small kernels of one env shard overlap the row kernels of the others; one CUDA-graph launch "
                           "per shard and step",
                    "blocking_single_handle_value": world * E * ne / e2e_blocking_s},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": NCU_DRAM_BYTES.get((a.precision, E, H)),
                         "traffic_source": "ncu --set full, tc_rows_pair_kernel, dram__bytes_read+write per launch "
                                           "(profiles/r01k_pair_kernels_ncu_summary.txt)",
                         "kernel": "lookahead (value network)", "kernel_ms": la_ms,
                         "flops_per_env_step": F, "peak_source": peaks["source"] + ", sustained bf16"},
        }
        # secondary figure: the ORCA + env-step kernels against the HBM roofline.  48 (H + 1) + 22 algorithmic bytes
        # per env step; these kernels are latency / FP32-issue bound, three orders of magnitude below the HBM ceiling.
        step_bytes = 48 * (H + 1) + 22
        step_ms = phases["orca"] + phases["step"]
        line["roofline_orca_step"] = {"bound": "hbm", "achieved": E * step_bytes / (step_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                                      "unit": "GB/s", "frac": E * step_bytes / (step_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                                      "bytes_per_env_step": step_bytes, "kernel_ms": step_ms,
                                      "kernel": "orca_humans_kernel + step_kernel (+ reset of finished episodes)"}
        if not a.no_cpu_baseline:
            cores = 1
            n = 16
            rate, steps, el = cpu_port_rate(n, H, a.sim, a.query_env, cores, a.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d steps of %d envs x %d humans, %.1f s, oracle C port, 1 thread" % (
                                        steps, n, H, el)}
        print(json.dumps(line))
    env.close(); pol.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
