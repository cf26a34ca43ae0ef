"""GPU: representation (q, mu) and Q errors vs the fp64 oracle for each edge-stage implementation."""
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
eng = model._engine_for(torch.device("cuda", 0))
IMPLS = {"stream": {"tensor_cores": 1, "sliced_edge": 1, "layer0_sums": 1},
         "stream_l0edge": {"tensor_cores": 1, "sliced_edge": 1, "layer0_sums": 0},
         "tc_general": {"tensor_cores": 1, "sliced_edge": 0},
         "simt": {"tensor_cores": 0}}
only = sys.argv[1:] or list(IMPLS)
print("# per output: el = max |cuda - f64| / max(|f64|, 1) per element; b = max |cuda - f64| / max(1, max_batch |f64|)  "
      "(delta_e's b is relative to max(|E_R|, |E_P|)); eager32 columns: the fp32 eager oracle under the same two norms")
print(f"{'impl':12s} {'case':26s} " + " ".join(f"{k + ':el':>9s} {k + ':b':>9s}" for k in ("delta_e", "q0", "q1", "rl_q")) + "   | eager32: same keys")
for impl in only:
    for k, v in IMPLS[impl].items():
        eng.set_option(k, v)
    for name, tag in CASES:
        g = load_gold(name)
        out = _score_batch(model, g, tag)
        errs, eag = [], []
        for k in ("delta_e", "q0", "q1", "rl_q"):
            ref = g[f"{tag}.f64.{k}"]
            scale = np.maximum(np.abs(ref), 1.0)
            sb = max(1.0, np.abs(ref).max()) if k != "delta_e" else max(1.0, np.abs(g[f"{tag}.f64.E_R"]).max(), np.abs(g[f"{tag}.f64.E_P"]).max())
            errs += [np.max(np.abs(out[k] - ref) / scale), np.abs(out[k] - ref).max() / sb]
            eag += [np.max(np.abs(g[f"{tag}.f32.{k}"] - ref) / scale), np.abs(g[f"{tag}.f32.{k}"] - ref).max() / sb]
        print(f"{impl:12s} {name + '/' + tag:26s} " + " ".join(f"{e:9.2e}" for e in errs) + "   " + " ".join(f"{e:9.2e}" for e in eag))
