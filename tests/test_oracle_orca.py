"""Analytic known-answer tests for the rvo2 restatement (the reference pins nothing at that boundary,
SURVEY §8(c)): parity for this third-party piece is 'unpinned', these KATs anchor it."""
import numpy as np
import pytest


def solve(o, pos, vel, pref, radius=0.31, max_speed=1.0, **kw):
    pos = np.asarray(pos, np.float32); vel = np.asarray(vel, np.float32)
    rad = np.full(len(pos), radius, np.float32)
    return o.rvo_new_velocity(pos, vel, rad, 0, max_speed, pref, return_info=True, **kw)


def test_no_neighbor_clips_pref(oracle_mod):
    v, (n, fail) = solve(oracle_mod, [[0, 0]], [[0, 0]], (3.0, 4.0))
    assert n == 0 and np.allclose(v, [0.6, 0.8], atol=1e-7)
    v, _ = solve(oracle_mod, [[0, 0]], [[0, 0]], (0.3, -0.4))
    assert np.array_equal(v, np.float32([0.3, -0.4]))


def test_far_agent_ignored(oracle_mod):
    v, (n, _) = solve(oracle_mod, [[0, 0], [10.0, 0]], [[0, 0], [0, 0]], (1.0, 0.0))
    assert n == 0 and np.array_equal(v, np.float32([1, 0]))          # strict distSq < neighborDist^2
    v, (n, _) = solve(oracle_mod, [[0, 0], [9.99, 0]], [[0, 0], [0, 0]], (1.0, 0.0))
    assert n == 1


def test_max_neighbors_keeps_nearest(oracle_mod):
    rs = np.random.RandomState(0)
    pos = np.vstack([[0, 0], rs.uniform(-6, 6, (20, 2))])
    vel = np.zeros_like(pos)
    v, (n, _) = solve(oracle_mod, pos, vel, (1.0, 0.0), max_neighbors=10)
    assert n == 10
    # same result when the 10 farthest are removed
    d = np.sum(pos[1:] ** 2, axis=1)
    keep = np.sort(np.argsort(d)[:10]) + 1
    v2, (n2, _) = solve(oracle_mod, np.vstack([pos[:1], pos[keep]]), np.zeros((11, 2)), (1.0, 0.0))
    assert n2 == 10 and np.array_equal(v, v2)


def test_constraints_satisfied_and_speed_bound(oracle_mod):
    """When LP2 succeeds the result satisfies every half-plane and the speed disc."""
    rs = np.random.RandomState(1)
    ok = 0
    for _ in range(300):
        n = rs.randint(2, 8)
        pos = rs.uniform(-4, 4, (n, 2)); vel = rs.uniform(-1, 1, (n, 2))
        pref = rs.uniform(-3, 3, 2)
        v, (nl, fail) = solve(oracle_mod, pos, vel, pref)
        assert np.hypot(*v) <= 1.0 + 1e-5
        ok += fail == nl
    assert ok > 100


def test_head_on_symmetry(oracle_mod):
    """Two agents head-on: mirror symmetry about the x axis of the pair of solutions."""
    pos = [[-2, 0], [2, 0]]; vel = [[1, 0], [-1, 0]]
    va, _ = solve(oracle_mod, pos, vel, (4.0, 0.0))
    vb, _ = solve(oracle_mod, pos[::-1], vel[::-1], (-4.0, 0.0))
    assert np.allclose(va, -vb, atol=1e-6)
    assert va[0] < 1.0 and np.hypot(*va) <= 1 + 1e-6


def test_rotation_equivariance(oracle_mod):
    rs = np.random.RandomState(2)
    for _ in range(50):
        n = rs.randint(2, 6)
        pos = rs.uniform(-3, 3, (n, 2)); vel = rs.uniform(-1, 1, (n, 2)); pref = rs.uniform(-1, 1, 2)
        v, (nl, fail) = solve(oracle_mod, pos, vel, pref)
        R = np.array([[0, -1], [1, 0]], float)                # 90 deg: exact in float32
        vr, _ = solve(oracle_mod, pos @ R.T, vel @ R.T, R @ pref)
        assert np.allclose(R @ v, vr, atol=2e-5), (v, vr)


def test_overlap_branch(oracle_mod):
    """||rp|| < R: the collision branch pushes apart within one time step (1/timeStep)."""
    v, (n, fail) = solve(oracle_mod, [[0, 0], [0.3, 0]], [[0, 0], [0, 0]], (1.0, 0.0))
    assert n == 1
    # u = (R/dt - |w|) * unitW with w = -rp/dt => u = (0.62/0.25 - 1.2) * (-1, 0) ; point = 0.5 u
    assert v[0] <= 0.5 * -(0.62 / 0.25 - 0.3 / 0.25) + 1e-5


def test_permutation_invariance(oracle_mod):
    rs = np.random.RandomState(3)
    for _ in range(50):
        n = rs.randint(3, 8)
        pos = rs.uniform(-4, 4, (n, 2)); vel = rs.uniform(-1, 1, (n, 2)); pref = rs.uniform(-1, 1, 2)
        v, _ = solve(oracle_mod, pos, vel, pref)
        perm = np.concatenate([[0], 1 + rs.permutation(n - 1)])
        v2, _ = solve(oracle_mod, pos[perm], vel[perm], pref)
        assert np.array_equal(v, v2)


def test_reference_call_arguments(oracle_mod):
    """orca.py:99-104 issues (0.25, 10, 10, 5, 5, 0.3, 1) / radius 0.31."""
    import oracle.refshim  # noqa: F401  (rvo2 stand-in lives there)
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(oracle_mod.__file__), "refshim"))
    import rvo2
    sim = rvo2.PyRVOSimulator(0.25, 10, 10, 5, 5, 0.3, 1)
    sim.addAgent((0.0, 0.0), 10, 10, 5, 5, 0.3 + 0.01 + 0, 1.0, (0, 0))
    sim.addAgent((1.0, 0.2), 10, 10, 5, 5, 0.3 + 0.01 + 0, 1, (0, 0))
    sim.setAgentPrefVelocity(0, (2.0, 0.0)); sim.setAgentPrefVelocity(1, (0, 0))
    sim.doStep()
    v = sim.getAgentVelocity(0)
    agents = np.zeros((3, 8)); agents[0] = [9, 9, 0, 0, 9, 9, .3, 1]
    agents[1] = [0, 0, 0, 0, 2, 0, .3, 1]; agents[2] = [1.0, 0.2, 0, 0, 1.0, 0.2, .3, 1]
    hv = oracle_mod.human_actions(oracle_mod.EnvCfg.default(), agents)
    assert tuple(hv[0]) == v
    assert sim.getNumAgents() == 2 and np.allclose(sim.getAgentPosition(0), np.array(v) * 0.25)
