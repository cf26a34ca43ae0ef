"""Per-launch time / DRAM-byte table of ONE timed step from an ncu launch list (read here, no GPU), plus per-family sums and
profiles/ncu_traffic.json (what bench.py's `roofline.traffic` / `dram_by_family` read).
Usage: python scripts/launch_table.py gpurun_out/launches.csv profiles/ncu_traffic_<tag>.txt [--json]"""
import collections, csv, json, os, re, sys

src, dst = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
by = collections.OrderedDict()
for r in rows[1:]:
    by.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
ids = list(by)


def short(n):
    n = re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    return n.replace("void ", "").replace("rlvd::", "")


def family(n):
    if "neighbor" in n or "struct_scan" in n or "row_ptr" in n:
        return "neighbor"
    if "edge_pack" in n:
        return "edge_pack"
    if "edge_stream" in n:
        return "edge"
    if "readout" in n:
        return "readout"
    if any(k in n for k in ("tc_linear", "mix_update", "embed", "axpy")):
        return "node"
    return "other"


starts = [n for n, i in enumerate(ids) if "neighbor_kernel<(int)0>" in by[i]["name"] or "neighbor_kernel<0>" in by[i]["name"]]
seq = ids[starts[1]:starts[2]]          # the second scoring step of the run (the first one is warm-up)
fam = collections.OrderedDict()
out = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, bench.py config 2 "
       "(128 replicas), every kernel of ONE timed step (cold cache, serialised launches)"]
es = []
for i in seq:
    d = by[i]
    n = short(d["name"])
    t = d["gpu__time_duration.sum"] / 1e6
    rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
    out.append(f"{n[:44]:44s} {t:8.3f} ms  rd {rd / 1e9:7.3f} GB  wr {wr / 1e9:7.3f} GB")
    a = fam.setdefault(family(n), [0.0, 0.0])
    a[0] += t
    a[1] += rd + wr
    if "edge_stream_kernel<(bool)1>" in d["name"] or "edge_stream_kernel<1>" in n:
        es.append((rd, wr))
out.append("")
for f, (t, b) in fam.items():
    out.append(f"family {f:10s} {t:8.3f} ms  {b / 1e9:8.3f} GB")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
if "--json" in sys.argv:
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tj = json.load(open(p))
    rel = os.path.relpath(dst, ROOT)
    tj["source"] = f"{rel} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, python bench.py --steps 2 --warmup 1, 128 replicas per GPU)"
    tj["dram_bytes_read_per_launch"], tj["dram_bytes_write_per_launch"] = int(es[0][0]), int(es[0][1])
    tj["edge_stream_bytes_per_launch"] = int(es[0][0] + es[0][1])
    tj["dram_bytes_per_step_by_family"] = {"source": f"{rel}: dram__bytes_read.sum + dram__bytes_write.sum of every launch of one step, summed per family",
                                           **{f: int(b) for f, (t, b) in fam.items() if f != "other"}}
    json.dump(tj, open(p, "w"), indent=1)
