"""End-to-end demo of the device-resident deployment loop (the shape of `rl-deploy`, rlsim/drl/deploy.py:75-131, without
relaxation / energy logging): N independent 255-atom CrCoNi replicas, `horizon` LSS or TKS steps each, all stepping
together on one B200; per-replica XDATCAR / q_values.json / actions.json (/ times.json) through the buffered writer.

    python scripts/deploy_demo.py --replicas 64 --horizon 50 --mode tks --temperature 900 --out /tmp/deploy_demo
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rlvacdiffsim_b200 as rb  # noqa: E402
from rlvacdiffsim_b200.simulator import DeviceLSS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=16)
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--mode", default="lss", choices=["lss", "tks"])
    ap.add_argument("--temperature", type=float, default=700.0)
    ap.add_argument("--out", default="/tmp/rlvd_deploy_demo")
    ap.add_argument("--checkpoint", default=os.path.join(ROOT, "tests", "golden", "model_trained"))
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    model = rb.registry.get_model_class("dqn_v2").load(args.checkpoint).to("cuda").eval()
    reps = [rb.make_crconi_supercell(4, seed=args.seed + r) for r in range(args.replicas)]
    # deploy_drl.toml uses {alpha: 0, beta: 1, dqn: true}; deploy_drl_tks.toml {alpha: 1, beta: 0, dqn: false}
    qp = {"alpha": 0.0, "beta": 1.0, "dqn": True} if args.mode == "lss" else {"alpha": 1.0, "beta": 0.0, "dqn": False}
    lss = DeviceLSS(model, reps, 0, 3.528, 1, qp)
    writer = rb.TrajectoryWriter(args.out, np.stack([r.get_atomic_numbers() for r in reps]),
                                 np.stack([r.get_cell() for r in reps]), flush_every=32)
    np.random.seed(args.seed)
    lss.step(args.temperature, apply=False)                      # warm-up (weights upload, workspaces)
    t0 = time.perf_counter()
    res = lss.run(args.horizon, args.temperature, mode=args.mode, writer=writer)
    dt = time.perf_counter() - t0
    writer.close()
    n_q = sum(len(q) for r in res["Q"] for q in r)
    print(f"{args.replicas} replicas x {args.horizon} {args.mode.upper()} steps: {dt:.2f} s wall "
          f"({1e3 * dt / args.horizon:.2f} ms per step of all replicas, {n_q / dt:,.0f} Q-evals/s incl. host bookkeeping + writers)")
    if args.mode == "tks":
        print("simulated time per replica (s):", np.round([t[-1] for t in res["time"]][:8], 12).tolist(), "...")
    print("outputs:", args.out, sorted(os.listdir(args.out))[:6], "...")


if __name__ == "__main__":
    main()
