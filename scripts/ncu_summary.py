"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, warp-stall breakdown and instruction mix of the first
profiled kernel.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page("raw")
hdr, units = raw[0], raw[1]
print(f"# {rep}")
for row in raw[2:]:
    print("kernel:", row[hdr.index("Kernel Name")][:110])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:70s} {row[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
src = page("source")
hi = [i for i, r in enumerate(src) if "Source" in r and "# Samples" in r][0]
h = src[hi]
ix = {n: i for i, n in enumerate(h)}
data, seen = [], set()
for r in src[hi + 1:]:
    if len(r) <= 10 or not r[ix["# Samples"]].isdigit():
        continue
    if r[ix["Address"]] in seen:
        break
    seen.add(r[ix["Address"]])
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
toti = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print(f"\nfirst kernel: {len(data)} SASS instructions, {tot} warp samples, {toti} warp instructions executed")
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print("warp stall samples:")
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v / tot:6.3f}")
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = r[ix["Source"]].split()
    name = (m[1] if m[0].startswith("@") else m[0]).split(".")[0]
    op[name] += int(r[ix["Instructions Executed"]])
    ops[name] += int(r[ix["# Samples"]])
print("instruction mix (share of executed warp instructions | share of stall samples):")
for k, v in op.most_common(18):
    print(f"  {k:12s} {v / toti:6.3f} | {ops[k] / tot:6.3f}")
