"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (read here, no GPU).
Joins the SASS page of the report with nvdisasm's line table of the built library.
Usage: python scripts/ncu_lines.py <rep> <cubin-stem e.g. edge_stream> <kernel-substring> [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, stem, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
nth = int(sys.argv[5]) if len(sys.argv) > 5 else 0          # which matching launch of the report
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "rlvacdiffsim_b200", "librlvd_b200.so")], cwd=tmp, capture_output=True)
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f"{stem}.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
names = [r[raw[0].index("Kernel Name")] for r in raw[2:]]
his = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
which = [i for i, n in enumerate(names) if kern in n][nth]
h = rows[his[which]]
ix = {n: i for i, n in enumerate(h)}
end = his[which + 1] if which + 1 < len(his) else len(rows)
data = [r for r in rows[his[which] + 1:end] if len(r) > 10 and r[ix["# Samples"]].isdigit()]
# the function's line table: pick the .text section whose name contains kern and whose size matches best
secs = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern.split("<")[0] in l]
best = None
for i0 in secs:
    line_of, cur = {}, None
    for l in txt[i0 + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    if best is None or abs(len(line_of) - len(data)) < abs(len(best) - len(data)):
        best = line_of
addr = lambda a: int(a, 16) if a.startswith("0x") else int(a)
base = addr(data[0][ix["Address"]])
agg = collections.defaultdict(lambda: [0, 0])
for r in data:
    ln = best.get(addr(r[ix["Address"]]) - base)
    agg[ln][0] += int(r[ix["# Samples"]])
    agg[ln][1] += int(r[ix["Instructions Executed"]])
ts, ti = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
sagg = {st: sum(int(r[ix[st]]) for r in data) for st in stalls}
print("stalls: " + "  ".join(f"{k[6:]} {v / ts:.3f}" for k, v in sorted(sagg.items(), key=lambda kv: -kv[1])[:8]))
srcs = {}
print(f"# {names[which][:100]}: {len(data)} SASS instructions ({len(best)} in the line table), {ti} warp instructions, {ts} samples")
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = str(ln)
    if ln and os.path.exists(os.path.join(root, "rlvacdiffsim_b200", "csrc", ln[0])):
        srcs.setdefault(ln[0], open(os.path.join(root, "rlvacdiffsim_b200", "csrc", ln[0])).read().splitlines())
        text = f"{ln[0]}:{ln[1]}: " + srcs[ln[0]][ln[1] - 1].strip()[:100]
    print(f"{100 * v[1] / ti:5.1f}% instr {100 * v[0] / ts:5.1f}% samples  {text}")
