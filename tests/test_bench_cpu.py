"""bench.py contract pieces that need no GPU: the reference arm's JSON line and the workload constants."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "DeepFwFM inference samples/sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["value"] > 0
    cb = line["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "model", "DeepFMs.py"))
    assert cb["kind"] == ("reference" if have_ref else "port")       # the unmodified reference wherever build() installed it
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_constants_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    try:
        bench.set_workload("criteo")
        assert (bench.FIELD, bench.NUM, bench.CATS) == (39, 13, 26)
        assert bench.ALG_BYTES_PER_SAMPLE == 1304 and bench.ALG_BYTES_PER_BATCH == 1918564 and bench.MLP_FLOPS_PER_SAMPLE == 952800
        assert sum(bench.SIZES) == 1326055          # rows of all 39 tables (latency/criteo_latency.cpp:38-39 + 13 numeric)
        bench.set_workload("twitter")
        assert (bench.FIELD, bench.NUM, bench.CATS) == (47, 11, 36) and bench.ALG_BYTES_PER_SAMPLE == 1776
        assert bench.XV_UNIT is True
        bench.set_workload("criteo_qr")
        assert bench.MODEL_KW["qr_flag"] == 1 and bench.MODEL_KW["qr_collisions"] == 4 and sum(bench.SIZES[13:]) == 33762577
        assert bench.XV_UNIT is False
        bench.set_workload("criteo_pruned")
        assert bench.PRUNED is True and bench.ALG_BYTES_PER_SAMPLE == 1304
    finally:
        bench.set_workload("criteo")


def test_product_workload_constants_equal_the_oracle_copies():
    """bench.py sizes its tables from the package's own constants (nothing of oracle/ on the product side); the test
    infrastructure keeps its copies -- they must stay the same lists."""
    sys.path.insert(0, ROOT)
    from oracle import synth
    from xsdeepfwfm_deprecated_b200.utils import workloads
    for name in ("CRITEO_PAPER", "CRITEO_KAGGLE", "TWITTER_SYNTH"):
        assert getattr(workloads, name) == getattr(synth, name), name


def test_graph_segments_cover_every_step():
    """the timed K steps must all replay from CUDA graphs (G-step segments + one tail graph)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "tail_graph(k - done).replay()" in src and "steps_on_lanes(0, k - done)" not in src
