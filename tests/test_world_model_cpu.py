"""Host side of the world models (crowd_nav/policy/world_model.py): the parameter containers carry the reference's state-dict
keys and shapes (its checkpoints load unchanged) and refuse to compute on the CPU -- the forward is cn_world_predict
(csrc/world_model.cu), checked against the reference's outputs in tests/test_gpu_facade.py."""
import numpy as np
import pytest

from conftest import MODEL_WORLD_NAMES, load_model_world


@pytest.mark.parametrize("name", MODEL_WORLD_NAMES)
def test_world_model_containers_take_reference_checkpoints(name):
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
    assert np.array_equal(world.flat_weights(), g["world_weights"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        world(torch.zeros((1, H * 4)))
