"""GPU: print the parity table (CUDA path vs fp64 oracle, next to the fp32 eager oracle's own deviation)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rlvacdiffsim_b200 as rb  # noqa: E402
from test_gpu_parity import CASES, _score_batch  # noqa: E402
from conftest import GOLD, load_gold  # noqa: E402

model = rb.registry.get_model_class("dqn_v2").load(os.path.join(GOLD, "model_trained")).to("cuda").eval()
print(f"{'case':28s} {'key':8s} {'cuda vs f64':>12s} {'eager32 vs f64':>15s} {'max|ref|':>10s}")
for name, tag in CASES:
    g = load_gold(name)
    out = _score_batch(model, g, tag)
    for k in ("E_R", "delta_e", "q0", "q1", "rl_q"):
        ref = g[f"{tag}.f64.{k}"]
        scale = np.maximum(np.abs(ref), 1.0)
        err = np.max(np.abs(out[k] - ref) / scale)
        eager = np.max(np.abs(g[f"{tag}.f32.{k}"] - ref) / scale)
        print(f"{name + '/' + tag:28s} {k:8s} {err:12.2e} {eager:15.2e} {np.abs(ref).max():10.3f}")
    ok = int(np.argmax(out["rl_q"])) == int(g[f"{tag}.argmax"])
    print(f"{name + '/' + tag:28s} argmax   {'equal' if ok else 'DIFFERENT'}")
