"""Path configuration shared by the oracle modules (test infrastructure only).

Mirrors the constructor switches of the reference module that shape the forward
hot path (reference: model/DeepFMs.py:81-89) and derives the parameter names and
shapes the reference registers (model/DeepFMs.py:185-222, 246-283, 1066-1091;
model/QREmbeddingBag.py:135-141).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Tuple


@dataclass
class PathConfig:
    field_size: int
    feature_sizes: List[int]
    embedding_size: int = 10
    numerical: int = 13
    use_fm: bool = False
    use_fwfm: bool = True
    use_deep: bool = True
    use_fwlw: bool = False
    use_lw: bool = False
    embedding_bag: bool = False
    qr_flag: int = 0
    qr_operation: str = "mult"
    qr_collisions: int = 1
    qr_threshold: int = 200
    h_depth: int = 3
    deep_nodes: int = 400

    def __post_init__(self):
        self.feature_sizes = [int(n) for n in self.feature_sizes]
        assert len(self.feature_sizes) == self.field_size
        if self.qr_flag:
            # model/DeepFMs.py:125 -- qr forces the embedding-bag construction
            self.embedding_bag = True

    # -- helpers -----------------------------------------------------------
    def is_qr(self, f: int) -> bool:
        """model/DeepFMs.py:1071 -- QR only for tables larger than the threshold."""
        return bool(self.qr_flag) and self.feature_sizes[f] > self.qr_threshold

    def qr_rows(self, f: int) -> Tuple[int, int]:
        """model/QREmbeddingBag.py:135-136."""
        n = self.feature_sizes[f]
        return int(math.ceil(n / self.qr_collisions)), int(self.qr_collisions)

    def table_names(self, prefix: str, f: int) -> List[str]:
        if self.is_qr(f):
            return [f"{prefix}.{f}.weight_q", f"{prefix}.{f}.weight_r"]
        return [f"{prefix}.{f}.weight"]

    def state_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """Names and shapes of the reference state_dict for this configuration."""
        F, K, N = self.field_size, self.embedding_size, self.deep_nodes
        out: Dict[str, Tuple[int, ...]] = {}
        shallow = self.use_fm or self.use_fwfm
        if shallow:
            out["bias"] = (1,)
            for prefix, width, present in (
                ("fm_1st_embeddings", 1, not self.use_fwlw),
                ("fm_2nd_embeddings", K, True),
            ):
                if not present:
                    continue
                for f in range(F):
                    if self.is_qr(f):
                        nq, nr = self.qr_rows(f)
                        out[f"{prefix}.{f}.weight_q"] = (nq, width)
                        out[f"{prefix}.{f}.weight_r"] = (nr, width)
                    else:
                        out[f"{prefix}.{f}.weight"] = (self.feature_sizes[f], width)
            if self.use_lw:
                out["fm_1st.weight"] = (1, F)
            if self.use_fwlw:
                out["fwfm_linear.weight"] = (F, K)
            if self.use_fwfm:
                out["field_cov.weight"] = (F, F)
        if self.use_deep:
            if not shallow:
                raise ValueError("deep-only model is outside the hot path (SURVEY section 2 row 2)")
            widths = [F * K] + [N] * self.h_depth
            for l in range(1, self.h_depth + 1):
                out[f"net_1_linear_{l}.weight"] = (widths[l], widths[l - 1])
                out[f"net_1_linear_{l}.bias"] = (widths[l],)
            out["net_1_fc.weight"] = (1, N)
        return out

    def to_json(self) -> dict:
        return asdict(self)

    @staticmethod
    def from_json(d: dict) -> "PathConfig":
        return PathConfig(**d)
