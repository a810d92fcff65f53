"""Tabular models that consume the loader's packed ``[B, F]`` feature matrix."""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn as nn


class TabularMLP(nn.Module):
    """Plain MLP over the packed feature matrix (BASELINE.json's "64 float32
    columns" tables). Accepts fp32 / bf16 input; runs in bf16 autocast on GPU."""

    def __init__(self, num_features: int, hidden: Sequence[int] = (1024, 512, 256),
                 out_features: int = 1, dropout: float = 0.0):
        super().__init__()
        layers = []
        prev = num_features
        for h in hidden:
            layers += [nn.Linear(prev, h), nn.ReLU(inplace=True)]
            if dropout:
                layers.append(nn.Dropout(dropout))
            prev = h
        layers.append(nn.Linear(prev, out_features))
        self.net = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x.to(self.net[0].weight.dtype))


class EmbeddingTabularNet(nn.Module):
    """DLRM-style net for the reference's ``DATA_SPEC`` schema: one embedding
    table per int64 ``embeddings_name*`` / ``one_hot*`` column (cardinalities
    from the spec's ``high`` bound), concatenated into an MLP."""

    def __init__(self, cardinalities: Dict[str, int], embedding_dim: int = 16,
                 hidden: Sequence[int] = (512, 256), max_rows: int = 1 << 20):
        super().__init__()
        self.columns = list(cardinalities)
        self.tables = nn.ModuleList(
            [nn.Embedding(min(int(cardinalities[c]), max_rows), embedding_dim)
             for c in self.columns])
        self.sizes = [t.num_embeddings for t in self.tables]
        self.mlp = TabularMLP(embedding_dim * len(self.columns), hidden, 1)

    def forward(self, features: Sequence[torch.Tensor]) -> torch.Tensor:
        parts = []
        for table, size, col in zip(self.tables, self.sizes, features):
            idx = col.reshape(-1).long() % size
            parts.append(table(idx))
        return self.mlp(torch.cat(parts, dim=1))
