#!/usr/bin/env python
"""bench.py — action Q-evals/sec on 256-atom CrCoNi replicas (BASELINE.json metric), one rank per GPU.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU oracle port of the reference path, all host threads

A "step" scores every candidate hop of this rank's shard of BASELINE config 2 (1024 replicas of a 4x4x4
CrCoNi supercell with one vacancy, 12 candidates each, split over 8 GPUs = 128 replicas per GPU; weak
scaling).  Like-for-like with the reference every candidate runs TWO PaiNN passes (reactant + product).
`value` is timed with the inputs already in HBM; `e2e` goes through rlvd_score_reactions_host with pinned
host buffers (H2D of every input and D2H of every Q inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CKPT = os.path.join(ROOT, "tests", "golden", "model_trained")
Q_PARAMS = {"alpha": 0.0, "beta": 1.0, "dqn": True, "temperature": 700.0}   # demo/config.toml:9-12, T_end
REPLICAS_PER_GPU = 128        # 1024 replicas / 8 GPUs (BASELINE config 2)
N_CAND = 12
F, NRBF, L = 128, 20, 3


def make_shard(rank: int, n_rep: int, jitter: float = 0.0):
    import rlvacdiffsim_b200 as rb
    reps, spaces = [], []
    for r in range(n_rep):
        s = rb.make_crconi_supercell(4, seed=rank * n_rep + r, jitter=jitter)
        reps.append(s)
        spaces.append(rb.get_action_space(s, 3.528))
    return reps, spaces


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    self.rows.append([c.strip() for c in ln.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(2)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(n_rep: int, threads: int, repeats: int = 1):
    """Q-evals/s of the eager-PyTorch CPU restatement (oracle/) on `n_rep` replicas x 12 candidates."""
    import torch
    from oracle.painn_oracle import PaiNNOracle, score_actions
    torch.set_num_threads(threads)
    model = PaiNNOracle.load(CKPT, dtype=torch.float32)
    reps, spaces = make_shard(0, n_rep)
    score_actions(model, reps[0].numbers, reps[0].positions, reps[0].cell, spaces[0][:4], Q_PARAMS)   # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for s, a in zip(reps, spaces):
            score_actions(model, s.numbers, s.positions, s.cell, a, Q_PARAMS, batch_size=16)   # simulator.py:52
        best = min(best, time.perf_counter() - t0)
    return n_rep * N_CAND / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_rep = 2
    for _ in range(args.warmup):
        cpu_oracle_rate(1, threads)
    times = []
    for _ in range(args.steps):
        rate, dt = cpu_oracle_rate(n_rep, threads)
        times.append(dt)
    total = sum(times)
    value = args.steps * n_rep * N_CAND / total
    line = {"impl": "reference", "metric": "action Q-evals/sec", "value": value, "unit": "Q-evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 2 (1024 independent 256-site CrCoNi replicas, 255 atoms, 12 "
                                   "candidate hops each, LSS scoring) sharded 128 replicas per GPU",
                       "sample": f"{n_rep} replicas x 12 candidates per step (bounded CPU sample of the same workload)",
                       "q_params": Q_PARAMS},
            "cpu_baseline": {"value": value, "unit": "Q-evals/s", "cores": threads, "kind": "port",
                             "sample": f"{n_rep} replicas x 12 candidates x {args.steps} steps, eager PyTorch fp32 oracle"},
            "e2e": {"value": value, "unit": "Q-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas-per-gpu", type=int, default=REPLICAS_PER_GPU)
    ap.add_argument("--jitter", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (experiments), e.g. --opt fuse_update=1")
    ap.add_argument("--dedup", action="store_true", help="share the reactant pass between a replica's candidates")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import rlvacdiffsim_b200 as rb
    from rlvacdiffsim_b200.engine import cell_geometry
    from rlvacdiffsim_b200.simulator import ReplicaScorer, pack_replicas

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = rb.registry.get_model_class("dqn_v2").load(CKPT).to(dev).eval()
    eng = model._engine_for(dev)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng_sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    n_rep = args.replicas_per_gpu
    reps, spaces = make_shard(rank, n_rep, args.jitter)
    pk = pack_replicas(reps, spaces, dedup_reactants=args.dedup)
    n_q = len(pk["react_struct"])
    n_struct = len(pk["node_ptr"]) - 1
    M = int(pk["node_ptr"][-1])
    inv, img = cell_geometry(pk["cells"], eng.cutoff)
    qp = eng.q_params(Q_PARAMS, model.head_spec)
    d = {"Z": torch.from_numpy(pk["Z"].astype(np.int32)).to(dev),
         "pos": torch.from_numpy(pk["pos"]).to(dev),
         "cell": torch.from_numpy(pk["cells"].astype(np.float32)).to(dev),
         "inv": torch.from_numpy(inv).to(dev), "img": torch.from_numpy(img).to(dev),
         "ptr": torch.from_numpy(pk["node_ptr"]).to(dev),
         "rs": torch.from_numpy(pk["react_struct"]).to(dev), "ps": torch.from_numpy(pk["prod_struct"]).to(dev)}
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 512 MB > 126 MB L2

    def step_device():
        return eng.score_device(d["Z"], d["pos"], d["cell"], d["inv"], d["img"], d["ptr"], d["rs"], d["ps"], qp,
                                max_struct_nodes=int(np.diff(pk["node_ptr"]).max()))

    def step_host():
        return eng.score_host(pk["Z"], pk["pos"], pk["cells"], pk["node_ptr"], pk["react_struct"], pk["prod_struct"], qp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps."""
        total_ms, res = 0.0, None
        for _ in range(steps):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            a.record()
            res = fn()
            b.record()
            torch.cuda.synchronize(dev)
            wall = (time.perf_counter() - t0) * 1e3
            total_ms += max(a.elapsed_time(b), wall if fn is step_host else 0.0)
        return total_ms, res

    for _ in range(args.warmup):
        res = step_device()
    step_host()
    n_edges = int(res["n_edges"])

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    launches0 = eng.launch_count()
    ms_dev, res = timed(step_device, args.steps)
    launches = eng.launch_count() - launches0
    barrier()
    ms_e2e, res_h = timed(step_host, args.steps)
    barrier()
    clocks = sampler.stop() if sampler else None
    assert np.array_equal(res["rl_q"].cpu().numpy(), res_h["rl_q"])

    # per-kernel-family timing (CUDA events around each launch on the launching stream), separate pass
    eng.profile(True)
    eng.profile_reset()
    prof_steps = min(args.steps, 3)
    for _ in range(prof_steps):
        flush.fill_(1.0)
        step_device()
    fam = {k: eng.profile_read(k) for k in ("neighbor", "edge_pack", "edge", "node", "readout")}
    eng.profile(False)

    # whole LSS step on the device (enumerate -> expand -> K1..K5 -> select -> apply), next to the reference-shaped host
    # front end (get_action_space + pack per replica) it replaces; reported as an extra object, never as the headline
    from rlvacdiffsim_b200.simulator import DeviceLSS
    lss = DeviceLSS(model, reps, dev, 3.528, 1, {k: Q_PARAMS[k] for k in ("alpha", "beta", "dqn")})
    lss_rng = np.random.RandomState(1234 + rank)
    for _ in range(2):
        lss.step(Q_PARAMS["temperature"], uniforms=lss_rng.random_sample(n_rep))
    torch.cuda.synchronize(dev)
    lss_steps = min(args.steps, 3)
    t0 = time.perf_counter()
    for _ in range(lss_steps):
        lss.step(Q_PARAMS["temperature"], uniforms=lss_rng.random_sample(n_rep))
    torch.cuda.synchronize(dev)
    ms_lss = (time.perf_counter() - t0) * 1e3 / lss_steps
    t0 = time.perf_counter()
    n_front = min(n_rep, 8)
    front = make_shard(rank, n_front, args.jitter)
    pack_replicas(front[0], front[1])
    ms_host_front = (time.perf_counter() - t0) * 1e3 / n_front * n_rep      # supercell build + enumeration + packing, scaled

    # the same Q-values with each replica's reactant pass computed once instead of once per candidate (13 instead of 24
    # PaiNN passes per replica): an extra object for users, never the headline (the reference recomputes the reactant)
    dedup_obj = None
    if not args.dedup:
        pk2 = pack_replicas(reps, spaces, dedup_reactants=True)
        inv2, img2 = cell_geometry(pk2["cells"], eng.cutoff)
        d2 = {"Z": torch.from_numpy(pk2["Z"].astype(np.int32)).to(dev), "pos": torch.from_numpy(pk2["pos"]).to(dev),
              "cell": torch.from_numpy(pk2["cells"].astype(np.float32)).to(dev), "inv": torch.from_numpy(inv2).to(dev),
              "img": torch.from_numpy(img2).to(dev), "ptr": torch.from_numpy(pk2["node_ptr"]).to(dev),
              "rs": torch.from_numpy(pk2["react_struct"]).to(dev), "ps": torch.from_numpy(pk2["prod_struct"]).to(dev)}
        mx2 = int(np.diff(pk2["node_ptr"]).max())

        def step_dedup():
            return eng.score_device(d2["Z"], d2["pos"], d2["cell"], d2["inv"], d2["img"], d2["ptr"], d2["rs"], d2["ps"], qp,
                                    max_struct_nodes=mx2)
        for _ in range(2):
            step_dedup()
        ms_dd, res_dd = timed(step_dedup, min(args.steps, 3))
        ms_dd /= min(args.steps, 3)
        rel = float(np.abs(res_dd["rl_q"].cpu().numpy() - res["rl_q"].cpu().numpy()).max() /
                    np.abs(res["rl_q"].cpu().numpy()).max())
        dedup_obj = {"what": "same Q-values, reactant pass shared by a replica's candidates (--dedup)", "ms_per_step": ms_dd,
                     "value": world * n_q / (ms_dd * 1e-3), "unit": "Q-evals/s (rank 0 device time)",
                     "structures_per_step_per_gpu": len(pk2["node_ptr"]) - 1, "max_rel_diff_vs_headline_rl_q": rel}
        del d2

    # latency regime: ONE replica's 12 candidates per call — the reference's own deployment shape (deploy.py:75-94 steps
    # a single replica serially); small batches take the split pack / readout kernels.  Extra object, never the headline.
    one = pack_replicas(reps[:1], spaces[:1])
    inv1, img1 = cell_geometry(one["cells"], eng.cutoff)
    d1 = {"Z": torch.from_numpy(one["Z"].astype(np.int32)).to(dev), "pos": torch.from_numpy(one["pos"]).to(dev),
          "cell": torch.from_numpy(one["cells"].astype(np.float32)).to(dev), "inv": torch.from_numpy(inv1).to(dev),
          "img": torch.from_numpy(img1).to(dev), "ptr": torch.from_numpy(one["node_ptr"]).to(dev),
          "rs": torch.from_numpy(one["react_struct"]).to(dev), "ps": torch.from_numpy(one["prod_struct"]).to(dev)}
    mx1 = int(np.diff(one["node_ptr"]).max())

    def step_one():
        return eng.score_device(d1["Z"], d1["pos"], d1["cell"], d1["inv"], d1["img"], d1["ptr"], d1["rs"], d1["ps"], qp,
                                max_struct_nodes=mx1)
    for _ in range(5):
        step_one()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(20):
        step_one()["rl_q"].cpu()
    ms_one = (time.perf_counter() - t0) * 1e3 / 20

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel = the fused edge kernel: algorithmic bytes per launch:
        #   read x[N,3F] + mu[N,3F], write q, mu[N,3F]  → 40*F*N per structure-layer, + 7 B per edge (col + shift)
        edge_ms, edge_n = fam["edge"]
        n_nodes_total, edge_bytes = M, 40.0 * F * M + 7.0 * n_edges
        edge_avg_ms = edge_ms / max(edge_n, 1)
        achieved = edge_bytes / (edge_avg_ms * 1e-3) / 1e9 if edge_avg_ms > 0 else 0.0
        # the same kernel against the resource that actually bounds it: every edge slot gathers five
        # 128-byte table rows per 32-channel slice through the shared-memory crossbar, 128 B/clk/SM (layer 0 does not
        # run this kernel any more: it is per-node GEMMs over species-resolved radial sums)
        sm_clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        smem_peak = 128.0 * eng_sm_count * sm_clock_hz / 1e9
        smem_bytes = n_edges * 4 * 128.0 * 5
        smem_avg_ms = edge_avg_ms
        smem_achieved = smem_bytes / (smem_avg_ms * 1e-3) / 1e9 if smem_avg_ms > 0 else 0.0
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("edge_stream_bytes_per_launch")
        except Exception:
            pass
        # measured DRAM traffic of each kernel family (ncu, committed under profiles/) over its live time this run: which
        # families sit on the HBM roofline and which are bound elsewhere (edge: shared-memory crossbar; readout: FMA)
        dram_by_family = None
        try:
            tb = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["dram_bytes_per_step_by_family"]
            dram_by_family = {}
            for k in ("neighbor", "edge_pack", "edge", "node", "readout"):
                ms_k = fam[k][0] / max(prof_steps, 1)
                if ms_k > 0 and n_rep == REPLICAS_PER_GPU:
                    gbs = tb[k] / (ms_k * 1e-3) / 1e9
                    dram_by_family[k] = {"GB_per_step": tb[k] / 1e9, "GB/s": gbs, "frac_of_peak": gbs / hbm_peak}
        except Exception:
            pass
        total_prof = sum(v[0] for v in fam.values())
        value = world * n_q * args.steps / (ms_dev * 1e-3)
        e2e = world * n_q * args.steps / (ms_e2e * 1e-3)
        h2d = sum(int(np.asarray(pk[k]).nbytes) for k in ("Z", "pos", "node_ptr", "react_struct", "prod_struct")) + \
            int(pk["cells"].astype(np.float32).nbytes) + int(inv.nbytes) + int(img.nbytes)
        line = {
            "metric": "action Q-evals/sec", "value": value, "unit": "Q-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 2 (1024 independent 256-site CrCoNi replicas, 255 atoms, 12 "
                                   "candidate hops each, LSS scoring) sharded 128 replicas per GPU",
                       "replicas_per_gpu": n_rep, "q_evals_per_step_per_gpu": n_q, "structures_per_step_per_gpu": n_struct,
                       "nodes": M, "edges": n_edges, "reactant_dedup": bool(args.dedup), "jitter": args.jitter,
                       "l2": "512 MB flush buffer written between timed steps",
                       "q_params": Q_PARAMS, "parallelism": f"replica-sharded x{world}, no data-path collective"},
            "e2e": {"value": e2e, "unit": "Q-evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 6 * 4 * n_q + 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "edge_stream_kernel (K3a+K3c+K4 fused; one launch per interaction layer 1, 2)", "achieved": achieved,
                         "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": edge_bytes,
                         "avg_launch_ms": edge_avg_ms, "launches_timed": edge_n,
                         "smem": {"bound": "shared-memory crossbar (what ncu shows this kernel saturating)",
                                  "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s",
                                  "frac": smem_achieved / smem_peak if smem_peak else None,
                                  "gather_bytes_per_launch": smem_bytes,
                                  "peak_source": "128 B/clk/SM x SMs x sampled SM clock"}},
            "dram_by_family": dram_by_family,
            "dedup": dedup_obj,
            "single_replica": {"what": "one 255-atom replica, 12 candidate hops per call (24 structures), device-resident inputs, "
                                       "Q-values read back to the host every call (rank 0 wall clock)",
                               "ms_per_call": ms_one, "value": len(one["react_struct"]) / (ms_one * 1e-3), "unit": "Q-evals/s"},
            "device_lss": {"what": "one full LSS step of every replica on the device: rlvd_action_space -> expand -> K1..K5 -> "
                                   "softmax/draw -> apply, fp64 state resident in HBM, one 8-byte D2H per step",
                           "ms_per_step": ms_lss, "value": world * n_q / (ms_lss * 1e-3), "unit": "Q-evals/s (rank 0 wall clock)",
                           "host_front_end_ms_per_step": ms_host_front,
                           "host_front_end": "reference-shaped per-replica numpy enumeration + packing for the same shard "
                                             f"(measured on {n_front} replicas, scaled)"},
            "kernel_share": {k: (v[0] / total_prof if total_prof > 0 else None) for k, v in fam.items()},
            "kernel_ms_per_step": {k: v[0] / prof_steps for k, v in fam.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, dt = cpu_oracle_rate(12, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "Q-evals/s", "cores": threads, "kind": "port",
                                    "sample": f"12 replicas x 12 candidates ({dt:.1f} s), eager PyTorch fp32 oracle, "
                                              "DataLoader batches of 16 as simulator.py:52"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
