"""The torch mirrors of the fork's world models (crowd_nav/policy/world_model.py) against the reference's own outputs:
tests/golden/model_world_*.npz hold the reference AttentionWorld / MlpWorld weights and, per step, the velocities the
reference module predicted for the recorded human states.  Pure torch on the CPU: no CUDA involved."""
import numpy as np
import pytest

from conftest import MODEL_WORLD_NAMES, load_model_world


@pytest.mark.parametrize("name", MODEL_WORLD_NAMES)
def test_world_model_mirror_matches_reference(name):
    import torch
    from modelcrowdnav_b200.world_model import AttentionWorld, MlpWorld
    g = load_model_world(name)
    H = int(g["H"])
    world = MlpWorld(H) if str(g["world"]) == "mlp" else AttentionWorld()
    sd = world.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["world_keys"]]          # a reference checkpoint loads unchanged
    off, new = 0, {}
    for k, v in sd.items():
        new[k] = torch.from_numpy(g["world_weights"][off:off + v.numel()].reshape(tuple(v.shape)).copy())
        off += v.numel()
    assert off == g["world_weights"].size
    world.load_state_dict(new)
    world.eval()
    for t in range(len(g["reward"])):
        cur = torch.tensor(g["agents"][t][1:, :4], dtype=torch.float32).reshape(1, -1)     # px py vx vy per human
        with torch.no_grad():
            v = world(cur)[0].reshape(H, 2).numpy()
        assert np.max(np.abs(v - g["new_v"][t])) <= 1e-6, t
