"""Turn the ncu outputs under gpurun_out/<tag>/ into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py r1      # reads gpurun_out/r1/*, writes profiles/r1_*.md|csv
"""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
           "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            try:
                d.setdefault(r[ki].split("(")[0][-60:], []).append(float(r[vi].replace(",", "")) / 1000.0)
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write("kernel,launches,avg_us,min_us,share_of_listed_time\n")
        for k, v in d.items():
            f.write(f"{k},{len(v)},{sum(v)/len(v):.2f},{min(v):.2f},{sum(v)/tot:.3f}\n")
    return d


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    name_i = h.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {os.path.basename(rep)} (units in row 2 of the raw page)\n\n")
        for r in rows[2:]:
            f.write(f"## {r[name_i].split('(')[0]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in METRICS:
                if m in h:
                    f.write(f"| {m} | {r[h.index(m)]} | {rows[1][h.index(m)]} |\n")
            f.write("\n")


for prec in ("fp32", "bf16", "bf16x3", "b65536_bf16x3", "b65536_bf16"):
    p = os.path.join(src, f"launches_{prec}.csv")
    if os.path.exists(p):
        launches(p, os.path.join(dst, f"{tag}_launches_{prec}.csv"))
    p = os.path.join(src, f"full_{prec}.ncu-rep")
    if os.path.exists(p):
        full(p, os.path.join(dst, f"{tag}_ncu_full_{prec}.md"))
print("written:", sorted(os.listdir(dst)))
