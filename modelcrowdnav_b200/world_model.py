"""Learned human-motion models of the fork (crowd_nav/policy/world_model.py:20-106): torch modules with the reference's
parameter names, so its checkpoints load unchanged.  `ModelCrowdSim` (envs.py) asks one of them for every human's next
velocity instead of solving ORCA; the env step itself stays on the GPU (cn_env_set_human_actions + cn_env_step).

SGANWorld (world_model.py:108-268) wraps the third-party Social-GAN generator and its trajectory datasets; it is outside
the hot path and not provided.
"""
import math

import torch
from torch import nn

from .policy import mlp


def init_weight(m):
    if type(m) == nn.Linear:
        nn.init.xavier_uniform_(m.weight)


class MlpWorld(nn.Module):
    """world_model.py:20-50: (B, num_human * 4) -> (B, num_human * 2), tanh-bounded velocities."""

    def __init__(self, num_human, drop_rate=0.5, multihuman=True):
        super().__init__()
        if not multihuman:
            num_human = 1
        self.mlp = nn.Sequential(
            nn.Linear(num_human * 4, 128), nn.ReLU(True), nn.Dropout(drop_rate),
            nn.Linear(128, 64), nn.ReLU(True), nn.Dropout(drop_rate),
            nn.Linear(64, 12), nn.ReLU(True), nn.Linear(12, num_human * 2), nn.Tanh())
        self.mse = 0
        self.device = None

    def forward(self, x):
        return self.mlp(x)

    def noise_pre(self, x):
        x = self.forward(x)
        mean = math.sqrt(self.mse)
        return x + (torch.randn(x.shape) * mean).to(self.device)


class AttentionWorld(nn.Module):
    """world_model.py:53-106: the SARL block structure on (px, py, vx, vy) rows, one (vx, vy) per human."""

    def __init__(self, input_dim=4, with_global_state=True):
        super().__init__()
        mlp1_dims, mlp2_dims, attention_dims, mlp3_dims = [150, 100], [100, 50], [100, 100, 1], [150, 100, 100, 2]
        self.input_dim = input_dim
        self.with_global_state = with_global_state
        self.global_state_dim = mlp1_dims[-1]
        self.mlp1 = mlp(input_dim, mlp1_dims, last_relu=True)
        self.mlp2 = mlp(mlp1_dims[-1], mlp2_dims)
        self.attention = mlp(mlp1_dims[-1] * 2 if with_global_state else mlp1_dims[-1], attention_dims)
        self.mlp3_input_dim = mlp2_dims[-1] + input_dim
        self.mlp3 = mlp(self.mlp3_input_dim, mlp3_dims)
        self.attention_weights = None
        self.output_func = nn.Tanh()

    def forward(self, in_state):
        state = in_state.view((in_state.shape[0], -1, self.input_dim))
        size = state.shape
        mlp1_output = self.mlp1(state.reshape((-1, size[2])))
        mlp2_output = self.mlp2(mlp1_output)
        if self.with_global_state:
            global_state = torch.mean(mlp1_output.view(size[0], size[1], -1), 1, keepdim=True)
            global_state = global_state.expand((size[0], size[1], self.global_state_dim)).contiguous().view(
                -1, self.global_state_dim)
            attention_input = torch.cat([mlp1_output, global_state], dim=1)
        else:
            attention_input = mlp1_output
        scores = self.attention(attention_input).view(size[0], size[1], 1).squeeze(dim=2)
        scores_exp = torch.exp(scores) * (scores != 0).float()
        weights = (scores_exp / torch.sum(scores_exp, dim=1, keepdim=True)).unsqueeze(2)
        self.attention_weights = weights[0, :, 0].data.cpu().numpy()
        features = mlp2_output.view(size[0], size[1], -1)
        weighted_feature = torch.sum(torch.mul(weights, features), dim=1, keepdim=True)
        mul_weighted_feature = torch.cat([weighted_feature] * size[1], dim=1)
        joint_state = torch.cat([state, mul_weighted_feature], dim=2)
        return self.mlp3(joint_state.view((-1, self.mlp3_input_dim))).view((size[0], -1))


class SGANWorld(object):
    def __init__(self, *a, **kw):
        raise NotImplementedError("SGANWorld wraps the third-party Social-GAN generator (outside the B200 hot path)")
