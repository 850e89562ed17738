"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/crowdnav_b200.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__ as g
    from modelcrowdnav_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        g.build()
    return _capi


def test_header_symbols_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "crowdnav_b200.h")).read()
    declared = set(re.findall(r"\b(cn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = capi.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(capi.EXPORTS) <= declared


def test_struct_layouts_match_defaults(capi):
    e = capi.default_env_cfg()
    assert (e.time_limit, e.time_step, e.human_num, e.max_neighbors) == (25.0, 0.25, 5, 10)
    assert (e.collision_penalty, e.discomfort_dist, e.circle_radius, e.square_width) == (-0.25, 0.2, 4.0, 10.0)
    assert e.gamma == 0.9 and e.auto_reset == 0 and e.robot_visible == 0
    s = capi.default_sarl_cfg()
    assert list(s.mlp1_dims) == [150, 100] and list(s.mlp3_dims) == [150, 100, 100, 1]
    assert (s.speed_samples, s.rotation_samples, s.gamma, s.v_pref) == (5, 16, 0.9, 1.0)
    assert capi.load().cn_policy_param_count(C.byref(s)) == 96502


def test_no_cpu_fallback(capi):
    """Without a CUDA device every compute entry point must fail, never silently compute on the host."""
    lib = capi.load()
    if lib.cn_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.cn_env_create(C.byref(capi.default_env_cfg(num_envs=4)), 0, C.byref(h))
    assert rc == capi.CN_ECUDA and b"no CPU fallback" in lib.cn_last_error()
    rc = lib.cn_policy_create(C.byref(capi.default_sarl_cfg()), 0, C.byref(h))
    assert rc == capi.CN_ECUDA
    from modelcrowdnav_b200 import BatchedCrowdSim, CrowdNavError
    with pytest.raises(CrowdNavError):
        BatchedCrowdSim(8, 5)


def test_product_never_imports_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "modelcrowdnav_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower().replace("# oracle-free", ""), os.path.join(dp, f)
