"""One scoring call for a single 255-atom replica (12 candidates) -- run under ncu for the per-kernel launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import rlvacdiffsim_b200 as rb
from rlvacdiffsim_b200.simulator import ReplicaScorer
ck = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "model_trained")
model = rb.registry.get_model_class("dqn_v2").load(ck).to("cuda:0").eval()
atoms = rb.make_crconi_supercell(4, seed=0)
actions = rb.get_action_space(atoms, 3.528)
qp = {"alpha": 0.0, "beta": 1.0, "dqn": True, "temperature": 700.0}
sc = ReplicaScorer(model, 0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    q = sc.score([atoms], [actions], qp)[0]
torch.cuda.synchronize()
print(q[:3])
