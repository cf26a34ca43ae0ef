"""Per-kernel SASS opcode histogram of librlvd_b200.so (cuobjdump -sass): the mnemonics that prove tcgen05 / TMEM / bulk-TMA /
mbarrier use (B200_PROFILING.md): UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk),
SYNCS (mbarrier), UTCBAR (tcgen05.commit), plus FFMA2 (packed fp32).  Usage: python scripts/sass_opcodes.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rlvacdiffsim_b200", "librlvd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FMUL2", "IDP", "HMMA", "LDGSTS", "BAR.SYNC", "NANOSLEEP")
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("rlvd::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                hist[kern][w] += 1
print("# cuobjdump -sass rlvacdiffsim_b200/librlvd_b200.so (sm_100a): instruction counts per kernel; columns = mnemonic prefixes")
print(f"{'kernel':58s} {'total':>7s} " + " ".join(f"{w:>8s}" for w in WATCH))
for k, c in hist.items():
    print(f"{k[:58]:58s} {c['total']:7d} " + " ".join(f"{c[w]:8d}" for w in WATCH))
