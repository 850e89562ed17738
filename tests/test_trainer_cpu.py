"""Trainer host logic on the CPU (no CUDA): the eager form of the optimisation step against the reference's own Trainer
(tests/golden/training.npz: 100 SGD-momentum steps on the batches the reference's DataLoader drew)."""
import os

import numpy as np
import torch

from conftest import GOLDEN


class _Mem(object):
    def __init__(self, states, values):
        self.states, self.values = states, values

    def __len__(self):
        return self.states.shape[0]


def test_trainer_eager_matches_reference_sgd(weights0):
    from modelcrowdnav_b200.policy import make_value_network
    from modelcrowdnav_b200.trainer import Trainer
    g = dict(np.load(os.path.join(GOLDEN, "training.npz"), allow_pickle=False))
    torch.set_num_threads(1)
    torch.manual_seed(0)
    model = make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1])
    flat0 = torch.cat([p.detach().reshape(-1) for p in model.state_dict().values()]).numpy()
    assert np.array_equal(flat0, weights0)
    mem = _Mem(torch.from_numpy(g["il_states"]), torch.from_numpy(g["il_values"]).reshape(-1, 1))
    tr = Trainer(model, mem, torch.device("cpu"), 100)
    assert tr.mode == "eager"
    tr.set_learning_rate(0.01)
    idx = torch.from_numpy(g["sgd_idx"].astype(np.int64))
    loss = sum(float(tr._step(idx[i])) for i in range(idx.shape[0])) / idx.shape[0]
    flat = torch.cat([p.detach().reshape(-1) for p in model.state_dict().values()]).numpy()
    assert abs(loss - float(g["sgd_loss"])) < 1e-6
    assert np.max(np.abs(flat - g["sgd_weights"])) <= 1e-5


def test_trainer_requires_learning_rate_and_data():
    import pytest
    from modelcrowdnav_b200.policy import make_value_network
    from modelcrowdnav_b200.trainer import Trainer
    model = make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1])
    tr = Trainer(model, _Mem(torch.zeros((0, 5, 13)), torch.zeros((0, 1))), torch.device("cpu"), 100)
    with pytest.raises(ValueError, match="Learning rate is not set"):          # trainer.py:37-38,62-63
        tr.optimize_batch(1)
    tr.set_learning_rate(0.01)
    with pytest.raises(ValueError):
        tr.optimize_batch(1)
