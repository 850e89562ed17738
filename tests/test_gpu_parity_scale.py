"""Parity of the CUDA path at the BENCHMARKED sizes and on decisive argmax sets ("next round" item 1).

* `test_decisive_argmax`: every state of tests/golden/decisive_*.npz -- the reference's own SARL (trained weights, |V| up
  to 2) teacher-forced over ALL 500 test cases -- as ONE batch of several thousand envs: values within the relative bar,
  argmax identical on >= 99.9 % of >= 80 % decidable states.  A batch of 7,600 envs x 81 actions x 5 humans is 24,000 row
  tiles = 80+ rounds of the persistent CTA-pair kernels, so the X-slot recycling and mbarrier phase wrap of the steady state
  are compared with the reference, not with themselves.
* `test_benchmark_sizes_vs_oracle`: the three BASELINE.json single-GPU configurations at full size (8192 x 5 circle,
  8192 x 10 square, 4096 x 50 square), evolved for 3 steps, fp16 tensor-core path against the CPU oracle on a random sample
  of envs and against the FP32 CUDA path on ALL envs.
* `test_epsilon_greedy`: the Philox random-action branch of argmax_kernel (multi_human_rl.py:28-30) at epsilon = 0.3.
"""
import numpy as np
import pytest

from conftest import DECISIVE_NAMES, TIE_GAP, check_argmax, decidable, load_decisive, value_errors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcn():
    import modelcrowdnav_b200 as m
    assert m._capi.load().cn_device_count() > 0, "GPU tests need a CUDA device"
    return m


@pytest.mark.parametrize("precision", ["f32", "f16_tc"])
@pytest.mark.parametrize("name", DECISIVE_NAMES)
def test_decisive_argmax(mcn, weights_trained, name, precision):
    d = load_decisive(name)
    N, H = d["agents"].shape[0], int(d["H"])
    assert N >= 2500
    env = mcn.BatchedCrowdSim(N, H)
    pol = mcn.BatchedSARL(precision=precision)
    pol.load_weights(weights_trained)
    env.set_state(d["agents"], d["time"])
    env.orca()
    pol.lookahead(env, query_env=int(d["query_env"]))
    best, values = pol.read(env)
    ref = d["values"].astype(np.float64)                 # stored as fp32: 6e-8 relative, far inside both bars
    assert value_errors(values, ref, precision) <= 1.0
    agree, total = check_argmax(best, d["best"], ref, precision, min_decidable=int(0.8 * N))
    regret = ref.max(axis=1) - ref[np.arange(N), best]
    assert np.all(regret <= TIE_GAP[precision]), regret.max()
    print("decisive %s %s: %d / %d decidable states agree (of %d)" % (name, precision, agree, total, N))
    env.close(); pol.close()


# BASELINE.json configs[1], [2] (per GPU) and [4]: (envs, humans, sim rule, query_env, oracle sample size)
SIZES = [(8192, 5, 0, 0, 512), (8192, 10, 1, 1, 512), (4096, 50, 1, 0, 64)]


@pytest.mark.parametrize("E,H,rule,query_env,n_sample", SIZES)
def test_benchmark_sizes_vs_oracle(mcn, oracle_mod, weights_trained, E, H, rule, query_env, n_sample):
    o = oracle_mod
    ecfg, scfg = o.EnvCfg.default(), o.SarlCfg.default()
    env = mcn.BatchedCrowdSim(E, H, auto_reset=1, seed=11, sim_rule=rule)
    p16 = mcn.BatchedSARL(precision="f16_tc")
    p32 = mcn.BatchedSARL(precision="f32")
    p16.load_weights(weights_trained); p32.load_weights(weights_trained)
    env.reset_device()
    for _ in range(3):                                   # evolved states: humans moving, robot off its start
        mcn.rollout_step(p16, env, query_env)
    state, times = env.get_state()
    env.orca()
    hv = env.human_actions()
    p32.lookahead(env, query_env); b32, v32 = p32.read(env)
    p16.lookahead(env, query_env); b16, v16 = p16.read(env)
    # (1) every env: tensor-core path against the FP32 CUDA path
    assert value_errors(v16, v32, "f16_tc") <= 1.0
    check_argmax(b16, b32, v32, "f16_tc", min_decidable=int(0.6 * E))
    # (2) a random sample of envs against the CPU oracle, both precisions
    rs = np.random.RandomState(E + H)
    sample = np.sort(rs.choice(E, n_sample, replace=False))
    ob, ov = [], []
    for e in sample:
        obest, ovals, _ = o.lookahead(ecfg, scfg, weights_trained, state[e], times[e], p16.action_table, query_env, hv[e])
        ob.append(obest); ov.append(ovals)
    ov = np.stack(ov)
    assert value_errors(v32[sample], ov, "f32") <= 1.0
    assert value_errors(v16[sample], ov, "f16_tc") <= 1.0
    check_argmax(b32[sample], ob, ov, "f32", min_decidable=int(0.6 * n_sample))
    check_argmax(b16[sample], ob, ov, "f16_tc", min_decidable=int(0.6 * n_sample))
    # (3) the env step that follows (reward / done / info / next state) on the sampled envs, bit-exact
    reward, done, info, dmin = env.step(update=True)      # advances with the f16 path's pending actions
    acts, _ = p16.action_table[b16], None
    for k, e in enumerate(sample):
        r, dn, i, _ = o.step_outcome(ecfg, state[e], times[e], acts[e])
        assert (r, dn, i) == (reward[e], bool(done[e]), int(info[e])), e
    env.close(); p16.close(); p32.close()


def test_epsilon_greedy(mcn, weights_trained):
    """epsilon-greedy (multi_human_rl.py:28-30): with probability epsilon a uniformly random action of the 81 replaces the
    greedy one.  Rate within 3 sigma, uniform over the action table, per-env streams independent of the sharding."""
    E, H, eps = 16384, 5, 0.3
    env = mcn.BatchedCrowdSim(E, H, seed=5)
    pol = mcn.BatchedSARL(precision="f16_tc")
    pol.load_weights(weights_trained)
    env.reset_device()
    state, times = env.get_state()
    pol.lookahead(env, 0, epsilon=0.0); greedy, _ = pol.read(env)
    pol.lookahead(env, 0, epsilon=eps); picked, _ = pol.read(env)
    A = pol.A
    # a random pick coincides with the greedy action with probability 1 / A
    p_diff = eps * (A - 1) / A
    n_diff = int((picked != greedy).sum())
    sigma = np.sqrt(E * p_diff * (1 - p_diff))
    assert abs(n_diff - E * p_diff) <= 3 * sigma, (n_diff, E * p_diff, sigma)
    # uniform over the table: the differing picks spread over the A - 1 non-greedy actions.  The greedy action at reset is
    # (almost) the same for every env, so look at the raw histogram of the differing picks: each bin ~ n_diff / (A - 1)
    hist = np.bincount(picked[picked != greedy], minlength=A).astype(np.float64)
    dominant = np.bincount(greedy, minlength=A).argmax()
    bins = np.delete(hist, dominant)
    exp = bins.sum() / len(bins)
    chi2 = float(((bins - exp) ** 2 / exp).sum())
    assert chi2 < 80 + 4 * np.sqrt(2 * 80), chi2          # chi-square, 79 dof: mean 79, sigma 12.6
    assert bins.min() > 0
    # shard invariance: the same envs as two handles keyed by their global ids draw the same actions
    half = E // 2
    for k in range(2):
        env2 = mcn.BatchedCrowdSim(half, H, seed=5, env_id_offset=k * half)
        env2.set_state(state[k * half:(k + 1) * half], times[k * half:(k + 1) * half])
        pol.lookahead(env2, 0, epsilon=eps)
        picked2, _ = pol.read(env2)
        assert np.array_equal(picked2, picked[k * half:(k + 1) * half]), k
        env2.close()
    # a second lookahead on the same handle advances the per-env stream: different draws, same rate
    pol.lookahead(env, 0, epsilon=eps); again, _ = pol.read(env)
    assert not np.array_equal(again, picked)
    assert abs(int((again != greedy).sum()) - E * p_diff) <= 3 * sigma
    env.close(); pol.close()
