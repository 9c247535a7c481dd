"""Quotient-remainder embedding table: parameter container with the reference's names and shapes.

Mirrors ``QREmbeddingBag(num_categories, embedding_dim, num_collisions, operation, ..., mode, sparse)``
of the reference (model/QREmbeddingBag.py:111-154): two parameters ``weight_q`` (ceil(n/c), d) and
``weight_r`` (c, d), initialised ``uniform_(sqrt(1/n), 1)`` exactly as the reference's
``reset_parameters`` really does (it passes one positional bound, model/QREmbeddingBag.py:152-154).

The lookup itself -- q = idx // c, r = idx mod c, row = Wq[q] (*|+) Wr[r]
(model/QREmbeddingBag.py:156-174) -- is executed inside the fused CUDA gather kernel
(csrc/embed_fwfm.cu, ``fetch_row``) when the table belongs to a ``DeepFMs`` module; calling this
container directly runs the same kernel on a one-field problem.  ``operation='concat'`` changes the
embedding width and is rejected.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn.parameter import Parameter


class QREmbeddingBag(nn.Module):
    def __init__(self, num_categories, embedding_dim, num_collisions, operation="mult", max_norm=None,
                 norm_type=2.0, scale_grad_by_freq=False, mode="mean", sparse=False, _weight=None):
        super().__init__()
        if operation not in ("mult", "add"):
            raise ValueError(f"QREmbeddingBag operation {operation!r} is not supported on the B200 hot path "
                             "('concat' changes the embedding width)")
        if max_norm is not None or scale_grad_by_freq:
            raise ValueError("max_norm / scale_grad_by_freq are not supported")
        self.num_categories = int(num_categories)
        if not isinstance(embedding_dim, int):
            dims = list(embedding_dim)
            if len(dims) == 2 and dims[0] != dims[1]:
                raise ValueError("Embedding dimensions do not match!")
            embedding_dim = int(dims[0])
        self.embedding_dim = [embedding_dim, embedding_dim]
        self.num_collisions = int(num_collisions)
        self.operation = operation
        self.mode = mode
        self.sparse = sparse
        self.num_embeddings = [int(np.ceil(self.num_categories / self.num_collisions)), self.num_collisions]
        if _weight is None:
            self.weight_q = Parameter(torch.Tensor(self.num_embeddings[0], embedding_dim))
            self.weight_r = Parameter(torch.Tensor(self.num_embeddings[1], embedding_dim))
            self.reset_parameters()
        else:
            assert list(_weight[0].shape) == [self.num_embeddings[0], embedding_dim]
            assert list(_weight[1].shape) == [self.num_embeddings[1], embedding_dim]
            self.weight_q = Parameter(_weight[0])
            self.weight_r = Parameter(_weight[1])

    def reset_parameters(self):
        nn.init.uniform_(self.weight_q, np.sqrt(1 / self.num_categories))
        nn.init.uniform_(self.weight_r, np.sqrt(1 / self.num_categories))

    def forward(self, input, offsets=None, per_sample_weights=None):
        """input (B, 1) int64 on the table's CUDA device -> (B, d).  Bags of one index only."""
        if offsets is not None or per_sample_weights is not None:
            raise ValueError("only one-index bags (the DeepFMs usage) are supported")
        if input.dim() != 2 or input.shape[1] != 1:
            raise ValueError("expected input of shape (B, 1)")
        from .DeepFMs import _single_table_lookup
        return _single_table_lookup(self, input)

    def extra_repr(self):
        return f"{self.num_embeddings}, {self.embedding_dim}, operation={self.operation}, mode={self.mode}"
