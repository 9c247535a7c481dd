"""Input side of the hot path: the reference's loaders with array outputs and a binary cache.

The reference parses its CSVs line by line into Python lists (``utils/data_preprocess.py:54-72``: ``read_data`` splits every
line, builds ``[int(item) ...]`` / ``[float(item) ...]`` per row) and every caller then converts those lists to numpy, reshapes
and slices them per batch (``model/DeepFMs.py:532-539, 619-620, 766-770``).  With the forward at ~30 us per 4096 samples that
host work is the end-to-end cost, so the same functions are provided here with the same names, arguments and result keys, but

* ``index`` is one ``(N, C)`` int32 array (every Criteo / Twitter cardinality fits; ``DFW_XI_INT32`` consumes it as is, and
  ``index_dtype="int64"`` modules get ``.astype(np.int64)``), ``value`` one ``(N, num)`` float32 array, ``label`` an int8 array;
  ``np.array(result['index'])`` in the reference's callers works unchanged;
* parsing goes through the C CSV reader (pandas) instead of a Python loop;
* ``write_cache`` / ``read_cache`` keep the parsed table as three ``.npy`` columns + ``meta.json`` that are memory-mapped on load:
  a second run feeds ``predict_proba_host`` / ``eval_by_batch`` straight from the page cache, no parsing at all.

Column convention of the reference files (``read_data``): column 0 is the label, columns listed in ``num_list`` are numeric
values, every other column is a categorical index; ``feature_sizes = [1] * len(num_list) + [len(vocab_f) + 1 ...]`` from the
``category_emb`` file (lines ``field,token,index``; ``load_category_index``, ``utils/data_preprocess.py:18-27``).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

__all__ = ["load_category_index", "feature_sizes_from_index", "read_data", "read_data_twitter", "write_cache", "read_cache"]


def load_category_index(file_path: str, feature_dim_start: int = 0, dim: int = 39) -> List[Dict[str, int]]:
    """utils/data_preprocess.py:18-27: one {token: index} dict per field from lines ``field,token,index``."""
    cate_dict: List[Dict[str, int]] = [{} for _ in range(dim)]
    with open(file_path, "r") as f:
        for line in f:
            datas = line.strip().split(",")
            cate_dict[int(datas[0]) - feature_dim_start][datas[1]] = int(datas[2])
    return cate_dict


def feature_sizes_from_index(cate_dict: Sequence[Dict[str, int]], num_list: Sequence[int]) -> List[int]:
    """utils/data_preprocess.py:58-61: numeric fields first (one row each), then ``len(vocab) + 1`` per categorical field."""
    sizes = [1] * len(num_list)
    nums = set(num_list)
    for num, item in enumerate(cate_dict):
        if num + 1 not in nums:
            sizes.append(len(item) + 1)
    return sizes


def _split_columns(table: np.ndarray, num_list: Sequence[int]):
    ncol = table.shape[1]
    nums = [c for c in range(ncol) if c in set(num_list)]              # ascending column order, like the reference's enumerate
    cats = [c for c in range(1, ncol) if c not in set(num_list)]
    return nums, cats


def read_data(file_path: str, emb_file: Optional[str], num_list: Sequence[int], feature_dim_start: int = 0,
              dim: int = 39) -> dict:
    """utils/data_preprocess.py:54-72 with array outputs.  ``emb_file=None`` skips ``feature_sizes`` (inference on a
    checkpoint already knows them)."""
    import pandas as pd
    result = {"label": None, "value": None, "index": None, "feature_sizes": []}
    if emb_file is not None:
        result["feature_sizes"] = feature_sizes_from_index(load_category_index(emb_file, feature_dim_start, dim), num_list)
    df = pd.read_csv(file_path, header=None, dtype=np.float64, engine="c", float_precision="round_trip")   # == float(item)
    table = df.to_numpy()
    nums, cats = _split_columns(table, num_list)
    lab = table[:, 0]
    idx = table[:, cats]
    if not (np.all(lab == np.rint(lab)) and np.all(idx == np.rint(idx))):
        raise ValueError(f"{file_path}: label / categorical columns must hold integers (the reference int()s them)")
    if idx.size and (idx.min() < 0 or idx.max() >= 2 ** 31):
        raise ValueError(f"{file_path}: categorical index outside [0, 2^31)")
    result["label"] = lab.astype(np.int8) if lab.size and np.abs(lab).max() < 128 else lab.astype(np.int64)
    result["index"] = np.ascontiguousarray(idx.astype(np.int32))
    result["value"] = np.ascontiguousarray(table[:, nums].astype(np.float32))
    return result


def read_data_twitter(file_path: str, emb_file: Optional[str], num_list: Sequence[int], feature_dim_start: int = 0,
                      dim: int = 39, twitter_category: Optional[str] = None) -> dict:
    """utils/data_preprocess.py:30-51: parquet with the four engagement labels first; keeps ``twitter_category`` as the label,
    columns 1..len(num_list) as values, the rest as indices."""
    import pandas as pd
    result = {"label": None, "value": None, "index": None, "feature_sizes": []}
    if emb_file is not None:
        result["feature_sizes"] = feature_sizes_from_index(load_category_index(emb_file, feature_dim_start, dim), num_list)
    data = pd.read_parquet(file_path)
    for label in ["reply", "retweet", "retweet_comment", "like"]:
        if label != twitter_category and label in data.columns:
            data = data.drop(columns=[label])
    n_num = len(num_list)
    result["label"] = data[twitter_category].to_numpy().astype(np.int8)
    result["index"] = np.ascontiguousarray(data.iloc[:, n_num + 1:].to_numpy().astype(np.int32))
    result["value"] = np.ascontiguousarray(data.iloc[:, 1:n_num + 1].to_numpy().astype(np.float32))
    return result


# --------------------------------------------------------------------------------------------- binary columnar cache
_COLUMNS = (("index", "index.i32.npy"), ("value", "value.f32.npy"), ("label", "label.npy"))


def write_cache(result: dict, path: str) -> str:
    """Store a parsed table under directory ``path``: one ``.npy`` file per column (C order, fixed dtype) + ``meta.json``."""
    os.makedirs(path, exist_ok=True)
    n = len(result["label"])
    idx = np.ascontiguousarray(np.asarray(result["index"], dtype=np.int32).reshape(n, -1))
    val = np.ascontiguousarray(np.asarray(result["value"], dtype=np.float32).reshape(n, -1))
    lab = np.ascontiguousarray(np.asarray(result["label"]))
    for arr, (_, fname) in zip((idx, val, lab), _COLUMNS):
        np.save(os.path.join(path, fname), arr)
    meta = dict(rows=int(n), categorical=int(idx.shape[1]), numerical=int(val.shape[1]),
                feature_sizes=[int(x) for x in result.get("feature_sizes", [])], format=1)
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f)
    return path


def read_cache(path: str, mmap: bool = True) -> dict:
    """Inverse of ``write_cache``; with ``mmap`` the columns are memory-mapped (read-only), so slicing a batch touches only
    its pages."""
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    if meta.get("format") != 1:
        raise ValueError(f"{path}: unknown cache format {meta.get('format')!r}")
    out = {"feature_sizes": meta["feature_sizes"]}
    for key, fname in _COLUMNS:
        out[key] = np.load(os.path.join(path, fname), mmap_mode="r" if mmap else None)
    if out["index"].shape != (meta["rows"], meta["categorical"]) or out["value"].shape != (meta["rows"], meta["numerical"]):
        raise ValueError(f"{path}: column shapes do not match meta.json")
    return out
