"""Network containers with the reference's names, constructor signatures and ``state_dict`` layout.

``MLP`` mirrors offlinerlkit/nets/mlp.py:9-33 and ``EnsembleLinear`` offlinerlkit/nets/ensemble_linear.py:8-53
(parameter names, shapes and initialisation), so the reference's run_example scripts build the same objects
and checkpoints are interchangeable.  These classes only HOLD parameters and provide a plain forward for
evaluation-time inference; the training step never runs them -- ``policy.learn`` goes through the CUDA engine,
which re-points the parameters at its arena (engine/nets.py).
"""
from typing import List, Optional, Sequence

import torch
import torch.nn as nn


class MLP(nn.Module):
    def __init__(self, input_dim: int, hidden_dims: Sequence[int], output_dim: Optional[int] = None,
                 activation: type = nn.ReLU, dropout_rate: Optional[float] = None) -> None:
        super().__init__()
        widths = [int(input_dim)] + [int(h) for h in hidden_dims]
        blocks: List[nn.Module] = []
        for fan_in, fan_out in zip(widths, widths[1:]):
            blocks.append(nn.Linear(fan_in, fan_out))
            blocks.append(activation())
            if dropout_rate is not None:
                blocks.append(nn.Dropout(p=dropout_rate))
        self.output_dim = widths[-1]
        if output_dim is not None:
            blocks.append(nn.Linear(widths[-1], int(output_dim)))
            self.output_dim = int(output_dim)
        self.model = nn.Sequential(*blocks)
        self.hidden_dims = widths[1:]
        self.has_dropout = dropout_rate is not None
        self.activation_type = activation

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)


class EnsembleLinear(nn.Module):
    """E independent affine maps; weight [E, in, out], bias [E, 1, out] (+ the reference's ``saved_*`` shadows)."""

    def __init__(self, input_dim: int, output_dim: int, num_ensemble: int, weight_decay: float = 0.0) -> None:
        super().__init__()
        self.num_ensemble = num_ensemble
        self.weight_decay = weight_decay
        self.weight = nn.Parameter(torch.zeros(num_ensemble, input_dim, output_dim))
        self.bias = nn.Parameter(torch.zeros(num_ensemble, 1, output_dim))
        nn.init.trunc_normal_(self.weight, std=1 / (2 * input_dim ** 0.5))
        # registered as Parameters in the reference (ensemble_linear.py:25-26): they appear in state_dict()
        self.saved_weight = nn.Parameter(self.weight.detach().clone())
        self.saved_bias = nn.Parameter(self.bias.detach().clone())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            return torch.einsum("ij,bjk->bik", x, self.weight) + self.bias
        return torch.baddbmm(self.bias, x, self.weight)

    def load_save(self) -> None:
        self.weight.data.copy_(self.saved_weight.data)
        self.bias.data.copy_(self.saved_bias.data)

    def update_save(self, indexes: List[int]) -> None:
        self.saved_weight.data[indexes] = self.weight.data[indexes]
        self.saved_bias.data[indexes] = self.bias.data[indexes]

    def get_decay_loss(self) -> torch.Tensor:
        return self.weight_decay * (0.5 * (self.weight ** 2).sum())
