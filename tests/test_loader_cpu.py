"""Host-side loaders against the reference's own utils/data_preprocess.read_data: the fixture
tests/golden/loader_parity.npz was produced by the unmodified reference functions (tests/golden/make_loader_golden.py)."""
import os

import numpy as np
import pytest

from golden_util import GOLDEN
from xsdeepfwfm_deprecated_b200.utils import data_preprocess as dp

NUM_LIST = list(range(1, 14))


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "loader_parity.npz"))


def write_vocab(path, sizes):
    with open(path, "w") as f:
        for field, n in enumerate(sizes, start=1):
            if field in NUM_LIST:
                f.write(f"{field},0,0\n")
            else:
                for t in range(int(n)):
                    f.write(f"{field},tok{t},{t}\n")


@pytest.mark.parametrize("tag", ["int", "frac"])
def test_read_data_matches_the_reference_loader(fx, tmp_path, tag):
    csv, emb = tmp_path / "in.csv", tmp_path / "category_emb"
    csv.write_text(str(fx[f"{tag}::csv"]))
    write_vocab(emb, fx["vocab_sizes"])
    r = dp.read_data(str(csv), str(emb), NUM_LIST, feature_dim_start=1, dim=39)
    assert r["index"].dtype == np.int32 and r["value"].dtype == np.float32 and r["index"].flags.c_contiguous
    assert np.array_equal(r["label"], fx[f"{tag}::label"])
    assert np.array_equal(r["index"], fx[f"{tag}::index"])
    # the reference keeps Python floats and its callers make them float32 (model/DeepFMs.py:620 -> FloatTensor)
    assert np.array_equal(r["value"], fx[f"{tag}::value"].astype(np.float32))
    assert r["feature_sizes"] == fx[f"{tag}::feature_sizes"].tolist()
    assert [len(d) for d in dp.load_category_index(str(emb), 1, 39)] == fx["vocab_sizes"].tolist()
    # what the reference's callers do with the lists works on the arrays (main_all.py / model/DeepFMs.py:532-539)
    Xi = np.array(r["index"]).reshape((-1, 26, 1))
    assert Xi.shape == (400, 26, 1) and np.array(r["value"]).shape == (400, 13)


def test_cache_round_trip_is_memory_mapped_and_exact(fx, tmp_path):
    csv = tmp_path / "in.csv"
    csv.write_text(str(fx["frac::csv"]))
    r = dp.read_data(str(csv), None, NUM_LIST)
    assert r["feature_sizes"] == []
    r["feature_sizes"] = fx["frac::feature_sizes"].tolist()
    path = dp.write_cache(r, str(tmp_path / "cache"))
    c = dp.read_cache(path)
    assert isinstance(c["index"], np.memmap) and not c["index"].flags.writeable
    for k in ("index", "value", "label"):
        assert c[k].dtype == r[k].dtype and np.array_equal(c[k], r[k]), k
    assert c["feature_sizes"] == r["feature_sizes"]
    # a batch slice of the mapped columns is what predict_proba_host copies into its pinned staging buffers
    sl = np.ascontiguousarray(c["index"][128:256])
    assert sl.shape == (128, 26) and np.array_equal(sl, r["index"][128:256])
    eager = dp.read_cache(path, mmap=False)
    assert not isinstance(eager["value"], np.memmap) and np.array_equal(eager["value"], r["value"])


def test_loader_refuses_what_the_reference_would_crash_on(tmp_path):
    bad = tmp_path / "bad.csv"
    bad.write_text("0,1.0,2.5\n1,3.0,4\n")           # a fractional categorical index: the reference's int(item) raises
    with pytest.raises(ValueError):
        dp.read_data(str(bad), None, [1])
    neg = tmp_path / "neg.csv"
    neg.write_text("0,1.0,-3\n")
    with pytest.raises(ValueError):
        dp.read_data(str(neg), None, [1])
    (tmp_path / "c").mkdir()
    (tmp_path / "c" / "meta.json").write_text('{"format": 9}')
    with pytest.raises(ValueError):
        dp.read_cache(str(tmp_path / "c"))


def test_twitter_parquet_layout(tmp_path):
    """utils/data_preprocess.py:30-51: four engagement labels first, then the numeric block, then the categorical block."""
    import pandas as pd
    rng = np.random.default_rng(3)
    n, n_num, n_cat = 50, 11, 36
    cols = {"reply": rng.integers(0, 2, n), "retweet": rng.integers(0, 2, n), "retweet_comment": rng.integers(0, 2, n),
            "like": rng.integers(0, 2, n)}
    for i in range(n_num):
        cols[f"num{i}"] = rng.random(n).astype(np.float32)
    for i in range(n_cat):
        cols[f"cat{i}"] = rng.integers(0, 1000, n)
    df = pd.DataFrame(cols)
    p = tmp_path / "t.parquet"
    df.to_parquet(p)
    r = dp.read_data_twitter(str(p), None, list(range(1, n_num + 1)), twitter_category="like")
    assert np.array_equal(r["label"], df["like"].to_numpy())
    assert r["value"].shape == (n, n_num) and np.array_equal(r["value"][:, 0], df["num0"].to_numpy())
    assert r["index"].shape == (n, n_cat) and np.array_equal(r["index"][:, -1], df["cat35"].to_numpy())
