"""world_size-2 gloo tests of the N>1 host logic: environments shard with no data-path collective;
the only collectives are the gradient all-reduce of the trainer and the statistics all-reduce of the explorer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = fn(rank, world)
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.asarray(res, dtype=np.float64))
    finally:
        dist.destroy_process_group()


def _run(fn, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), fn, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(world)]


def _make_model():
    from modelcrowdnav_b200.policy import make_value_network
    torch.manual_seed(0)
    return make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1])


class _Mem(object):
    def __init__(self, states, values):
        self.states, self.values = states, values

    def __len__(self):
        return self.states.shape[0]


def _data(n=200):
    g = torch.Generator().manual_seed(1)
    return torch.rand((n, 5, 13), generator=g), torch.rand((n, 1), generator=g)


def _trainer_rank(rank, world):
    from modelcrowdnav_b200.trainer import Trainer
    torch.set_num_threads(1)
    model = _make_model()
    states, values = _data()
    half = states.shape[0] // world
    mem = _Mem(states[rank * half:(rank + 1) * half], values[rank * half:(rank + 1) * half])
    tr = Trainer(model, mem, torch.device("cpu"), 100, dist_group=dist.group.WORLD)
    tr.broadcast_weights()
    tr.set_learning_rate(0.01)
    for _ in range(3):
        tr._step(torch.arange(half))          # each rank: its own shard, gradients averaged over ranks
    return torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()


def test_gradient_allreduce_equals_single_process(tmp_path):
    """2 ranks x 100 samples with averaged gradients == 1 process x 200 samples (MSE mean reduction)."""
    from modelcrowdnav_b200.trainer import Trainer
    res = _run(_trainer_rank, tmp_path)
    assert np.array_equal(res[0], res[1])                 # replicas stay in lock-step
    torch.set_num_threads(1)
    model = _make_model()
    states, values = _data()
    tr = Trainer(model, _Mem(states, values), torch.device("cpu"), 200)
    tr.set_learning_rate(0.01)
    for _ in range(3):
        tr._step(torch.arange(200))
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    assert np.max(np.abs(res[0] - ref)) < 1e-6
    assert ref.size == 96502


def _epoch_rank(rank, world):
    """optimize_epoch with UNEQUAL replay sizes per rank (rank-local buffers keep only success / collision episodes)."""
    from modelcrowdnav_b200.trainer import Trainer
    torch.set_num_threads(1)
    model = _make_model()
    states, values = _data(330)
    n = 130 if rank == 0 else 330          # ceil(130 / 100) = 2 steps vs ceil(330 / 100) = 4 steps if left unsynchronised
    tr = Trainer(model, _Mem(states[:n], values[:n]), torch.device("cpu"), 100, dist_group=dist.group.WORLD)
    tr.broadcast_weights()
    tr.set_learning_rate(0.01)
    calls = []
    orig = tr._sync_gradients
    tr._sync_gradients = lambda: (calls.append(1), orig())[1]
    loss = tr.optimize_epoch(2)
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    return np.concatenate([[len(calls), loss], w])


def test_optimize_epoch_unequal_memory_sizes(tmp_path):
    """every rank must issue the same number of gradient all-reduces per epoch, or NCCL hangs."""
    res = _run(_epoch_rank, tmp_path)
    assert res[0][0] == res[1][0] == 2 * 4                      # 2 epochs x ceil(max(130, 330) / 100) steps on BOTH ranks
    assert np.array_equal(res[0][2:], res[1][2:])               # replicas stay in lock-step
    assert np.isfinite(res[0][1]) and np.isfinite(res[1][1])


def _stats_rank(rank, world):
    from modelcrowdnav_b200.explorer import Explorer
    ex = Explorer(None, None, torch.device("cpu"), dist_group=dist.group.WORLD)
    counts = np.array([3 + rank, 1, 2 * rank, 7, 10], dtype=np.float64)
    sums = np.array([10.5 * (rank + 1), 2.0, 25.0 * rank, 0.3, -1.5 + rank], dtype=np.float64)
    c, s = ex._all_reduce(counts, sums)
    return np.concatenate([c, s])


def test_explorer_statistics_allreduce(tmp_path):
    res = _run(_stats_rank, tmp_path)
    exp = np.array([7, 2, 2, 14, 20, 31.5, 4.0, 25.0, 0.6, -2.0])
    assert np.allclose(res[0], exp) and np.array_equal(res[0], res[1])


def test_shards_are_disjoint_and_cover_all_cases():
    """Contiguous global-id sharding used by bench.py / the batched explorer: rank r owns [r*E, (r+1)*E)."""
    from modelcrowdnav_b200 import scenes
    world, E = 4, 6
    full = scenes.generate_batch("test", range(world * E))
    for r in range(world):
        shard = scenes.generate_batch("test", range(r * E, (r + 1) * E))
        assert np.array_equal(shard, full[r * E:(r + 1) * E])


def test_training_case_shards_are_disjoint():
    """imitation-learning demonstrations and RL roll-outs of different ranks never share a case id, for any
    world size (rank r demonstrates [r * per_rank, (r + 1) * per_rank); RL iteration i continues from a common base)."""
    from modelcrowdnav_b200.train_loop import shard_cases
    for world in (1, 2, 3, 8):
        il_episodes, k = 500, 64
        per_rank = (il_episodes + world - 1) // world
        seen = set()
        for r in range(world):
            cases = set(range(r * per_rank, (r + 1) * per_rank))
            assert not (cases & seen)
            seen |= cases
        base = world * per_rank
        assert max(seen) < base
        for it in range(5):
            for r in range(world):
                first = shard_cases(base, it, world, r, k)
                cases = set(range(first, first + k))
                assert not (cases & seen), (world, it, r)
                seen |= cases
        assert seen == set(range(base + 5 * world * k))       # no holes either


def test_product_scene_generator_matches_reference_fixture(oracle_mod):
    """modelcrowdnav_b200.scenes (product) == the reference's reset() scenes (pinned through the oracle goldens)."""
    from conftest import load_traj
    from modelcrowdnav_b200 import scenes
    tr = load_traj("square10_qfalse")
    for case, rec in tr["cases"].items():
        phase, c = case.split("_")
        s = scenes.generate_scene(phase, int(c), human_num=10, rule="square_crossing")
        assert np.array_equal(s, rec["agents"][0])
    tr = load_traj("circle5_qfalse")
    for case, rec in tr["cases"].items():
        phase, c = case.split("_")
        assert np.array_equal(scenes.generate_scene(phase, int(c)), rec["agents"][0])


def test_native_scene_generator_is_bit_identical():
    """cn_scenes_generate (csrc/scenes_host.cu: MT19937 + crowd_sim.py:165-217 in C++) == the Python generator that the
    reference fixtures pin, for every rule / attribute combination it covers; out-of-range seeds are refused like numpy's."""
    import ctypes as C
    from modelcrowdnav_b200 import _capi, scenes
    for kw in (dict(), dict(rule="square_crossing", human_num=10), dict(randomize_attributes=True, human_num=7),
               dict(rule="square_crossing", human_num=3, randomize_attributes=True), dict(human_num=1, circle_radius=5.0)):
        for phase in ("train", "val", "test"):
            cases = list(range(40)) + [499, 31337]
            a = scenes.generate_batch(phase, cases, **kw)
            b = np.stack([scenes.generate_scene(phase, c, **kw) for c in cases])
            assert np.array_equal(a, b), (kw, phase)
    seeds = np.array([-1], np.int64)
    out = np.zeros((1, 6, 8))
    rc = _capi.load().cn_scenes_generate(1, seeds.ctypes.data_as(C.c_void_p), 5, 0, 4.0, 10.0, 0.3, 1.0, 0.2, 0.3, 1.0, 0,
                                         out.ctypes.data_as(C.c_void_p))
    assert rc == _capi.CN_EINVAL
