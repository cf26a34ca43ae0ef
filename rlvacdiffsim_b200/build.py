"""In-tree build of ``librlvd_b200.so`` (sm_100a only) with plain ``nvcc`` — no torch headers, no JIT cache.

The built library sits next to this file so that it travels to the GPU box with the source snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librlvd_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["engine.cu", "neighbor.cu", "node.cu", "tc_linear.cu", "edge.cu", "tc_edge.cu", "edge_stream.cu", "readout.cu", "actions.cu", "train.cu", "train_full.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("RLVD_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as fh:
                    h.update(name.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` for sm_100a and link ``librlvd_b200.so``.  Returns the library path."""
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == fp:
                return LIB
    nvcc = _nvcc()
    objs = [os.path.join(CSRC, s.replace(".cu", ".o")) for s in SOURCES]

    def compile_one(pair):
        src, obj = pair
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, res

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        results = list(pool.map(compile_one, zip(SOURCES, objs)))
    log = []
    for src, res in results:
        log.append(f"== {src}\n{res.stderr}")
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    with open(os.path.join(CSRC, ".ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(STAMP, "w") as fh:
        fh.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
