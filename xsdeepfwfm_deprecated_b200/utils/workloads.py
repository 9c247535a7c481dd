"""Field layouts and table cardinalities of the workloads BASELINE.json names.

Product-side constants: bench.py and the examples size their synthetic tables from these.  (oracle/synth.py keeps its
own copies for the test infrastructure; tests/test_bench_cpu.py checks the two agree.)
"""

# Paper Criteo cardinalities: /root/reference/latency/criteo_latency.cpp:38-39 (13 one-row numeric tables in front)
CRITEO_PAPER = [1] * 13 + [1458, 556, 245197, 166166, 306, 20, 12055, 634, 4, 46330, 5229, 243454,
                           3177, 27, 11745, 225322, 11, 4727, 2058, 5, 238640, 18, 16, 67856, 89, 50942]
# Un-thresholded Kaggle display-advertising cardinalities (not in the reference)
CRITEO_KAGGLE = [1] * 13 + [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593,
                            3194, 27, 14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105,
                            142572]
# Synthetic Twitter RecSys2020 shape: 11 dense + 36 sparse (/root/reference/model/Datasets.py:41-42)
TWITTER_SYNTH = [1] * 11 + ([3, 3, 3, 16777216, 67, 4, 16, 16777216, 8388608, 16777216, 1048576, 2097152,
                             1048576, 4, 32, 8, 25] + [4096] * 7 + [64, 512, 128, 1048576, 1048576]
                            + [262144] * 4 + [1048576] * 3)
