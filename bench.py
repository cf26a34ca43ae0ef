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
            self._halt.wait(0.05)

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



# ------------------------------------------------------------------------------------------------------------------
# The other BASELINE configs (1, 3, 4, 5): extra objects on the default line, or the whole line with --config K.
# ------------------------------------------------------------------------------------------------------------------
LSS_1200 = {"alpha": 0.0, "beta": 1.0, "dqn": True, "temperature": 1200.0}      # demo/config.toml:9-12, T_start
TKS_900 = {"alpha": 1.0, "beta": 0.0, "dqn": False, "temperature": 900.0}       # configs/deploy_drl_tks.toml:9-12,21


def _device_batch(pk, eng, dev):
    import torch
    from rlvacdiffsim_b200.engine import cell_geometry
    inv, img = cell_geometry(pk["cells"], eng.cutoff)
    return {"Z": torch.from_numpy(pk["Z"].astype(np.int32)).to(dev), "pos": torch.from_numpy(pk["pos"]).to(dev),
            "cell": torch.from_numpy(pk["cells"].astype(np.float32)).to(dev), "inv": torch.from_numpy(inv).to(dev),
            "img": torch.from_numpy(img).to(dev), "ptr": torch.from_numpy(pk["node_ptr"]).to(dev),
            "rs": torch.from_numpy(pk["react_struct"]).to(dev), "ps": torch.from_numpy(pk["prod_struct"]).to(dev),
            "max_nodes": int(np.diff(pk["node_ptr"]).max())}


def _time_calls(fn, n, dev):
    """Mean device milliseconds of n calls (CUDA events on the launching stream around the whole run)."""
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b) / n, out


def bench_config1(model, eng, dev, cpu: bool = True) -> dict:
    """BASELINE config 1: every candidate hop of the reference's demo POSCAR_0 (215 atoms, rhombohedral cell) with the bundled
    DQN model — the reference's own CPU-runnable case and its deployment shape (one replica, 12 candidates per call)."""
    import torch
    import rlvacdiffsim_b200 as rb
    from rlvacdiffsim_b200.simulator import pack_replicas
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_demo.npz"))
    demo = rb.read_poscar(os.path.join(ROOT, "tests", "golden", "POSCAR_0"))
    acts = rb.get_action_space_legacy(demo, 3.528)
    pk = pack_replicas([demo], [acts])
    d = _device_batch(pk, eng, dev)
    qp = eng.q_params(LSS_1200, model.head_spec)
    fn = lambda: eng.score_device(d["Z"], d["pos"], d["cell"], d["inv"], d["img"], d["ptr"], d["rs"], d["ps"], qp, max_struct_nodes=d["max_nodes"])
    for _ in range(5):
        fn()
    ms_dev, res = _time_calls(fn, 30, dev)
    host = lambda: eng.score_host(pk["Z"], pk["pos"], pk["cells"], pk["node_ptr"], pk["react_struct"], pk["prod_struct"], qp)
    host()
    t0 = time.perf_counter()
    for _ in range(20):
        res_h = host()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / 20
    ref = gold["lss1200.f64.rl_q"]
    err = float(np.abs(res_h["rl_q"] - ref).max() / max(1.0, np.abs(ref).max()))
    out = {"what": "BASELINE config 1: demo/poscars/POSCAR_0 (215 atoms, rhombohedral), 12 candidates (legacy enumerator, the one the "
                   "checkpoint was trained under), T = 1200 K, one call per step like deploy.py:75-94",
           "q_evals_per_call": len(acts), "ms_per_call_device": ms_dev, "ms_per_call_e2e": ms_e2e,
           "value": len(acts) / (ms_dev * 1e-3), "e2e_value": len(acts) / (ms_e2e * 1e-3), "unit": "Q-evals/s",
           "parity_rl_q_vs_fp64_oracle": err, "argmax_equal": bool(int(np.argmax(res_h["rl_q"])) == int(gold["lss1200.argmax"]))}
    if cpu:
        from oracle.painn_oracle import PaiNNOracle, score_actions
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        m32 = PaiNNOracle.load(CKPT, dtype=torch.float32)
        score_actions(m32, demo.numbers, demo.positions, demo.cell, acts[:2], LSS_1200)
        t0 = time.perf_counter()
        score_actions(m32, demo.numbers, demo.positions, demo.cell, acts, LSS_1200, batch_size=16)
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"value": len(acts) / dt, "unit": "Q-evals/s", "cores": threads, "seconds": dt}
    return out


def bench_config3(model, eng, dev, cpu: bool = False) -> dict:
    """BASELINE config 3: 2048-site supercell, 8 vacancies (2040 atoms, ~109.7k edges per structure), full enumeration."""
    import torch
    import rlvacdiffsim_b200 as rb
    from rlvacdiffsim_b200.simulator import pack_replicas
    s = rb.make_crconi_supercell(8, n_vac=8, seed=0)
    spaces = rb.get_action_space_device(eng, [s], 3.528, max_vacancies=8)[0]
    pk = pack_replicas([s], [spaces])
    d = _device_batch(pk, eng, dev)
    qp = eng.q_params(Q_PARAMS, model.head_spec)
    fn = lambda: eng.score_device(d["Z"], d["pos"], d["cell"], d["inv"], d["img"], d["ptr"], d["rs"], d["ps"], qp, max_struct_nodes=d["max_nodes"])
    for _ in range(3):
        res = fn()
    eng.profile(True)
    eng.profile_reset()
    ms_dev, res = _time_calls(fn, 5, dev)
    fam = {k: eng.profile_read(k)[0] / 5 for k in ("neighbor", "edge_pack", "edge", "node", "readout")}
    eng.profile(False)
    host = lambda: eng.score_host(pk["Z"], pk["pos"], pk["cells"], pk["node_ptr"], pk["react_struct"], pk["prod_struct"], qp)
    host()
    t0 = time.perf_counter()
    for _ in range(3):
        host()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / 3
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_syn2040_v8.npz"))
    ref = gold["lss700.f64.rl_q"]
    got = res["rl_q"].cpu().numpy()[: len(ref)]
    out = {"what": "BASELINE config 3: 8x8x8 cubic CrCoNi, 8 vacancies (2040 atoms), every candidate hop of every vacancy, T = 700 K; "
                   "device enumeration (rlvd_action_space, 8 vacancies)",
           "candidates": len(spaces), "structures": len(pk["node_ptr"]) - 1, "nodes": int(pk["node_ptr"][-1]), "edges": int(res["n_edges"]),
           "ms_per_step_device": ms_dev, "ms_per_step_e2e": ms_e2e, "value": len(spaces) / (ms_dev * 1e-3),
           "e2e_value": len(spaces) / (ms_e2e * 1e-3), "unit": "Q-evals/s", "kernel_ms": fam,
           "edge_path": "edge_stream (structure-resident)" if d["max_nodes"] <= 256 else "tc_edge (general tcgen05 gather kernel: structures above 256 atoms)",
           "parity_rl_q_vs_fp64_oracle_first6": float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))}
    if cpu:
        from oracle.painn_oracle import PaiNNOracle, score_actions
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        m32 = PaiNNOracle.load(CKPT, dtype=torch.float32)
        t0 = time.perf_counter()
        score_actions(m32, s.numbers, s.positions, s.cell, spaces[:2], Q_PARAMS, batch_size=16)
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"value": 2 / dt, "unit": "Q-evals/s", "cores": threads, "seconds": dt, "sample": "2 of the 96 candidates"}
    return out


def bench_config4(model, eng, dev, rank: int, n_rep: int, steps: int) -> dict:
    """BASELINE config 4: TKS deploy — Q scoring + residence-time clock every step (simulator.py:238-287) and the t_net time
    estimator on every frame (estimate_time.py:38-66; PaiNN representation of the checkpoint + a seeded head: no t_net weights
    are bundled), over a bounded sample of the 10k-step trajectory, queued without host synchronisation."""
    import torch
    import rlvacdiffsim_b200 as rb
    from rlvacdiffsim_b200.simulator import DeviceLSS
    reps = [rb.make_crconi_supercell(4, seed=5000 + rank * n_rep + r) for r in range(n_rep)]
    lss = DeviceLSS(model, reps, dev, 3.528, 1, {k: TKS_900[k] for k in ("alpha", "beta", "dqn")}, seed=2024, replica_id0=rank * n_rep)
    torch.manual_seed(0)
    tnet = rb.registry.get_model_class("t_net").load_representation(model.reaction_model, n_feat=64, pooling="mean", dropout_rate=0.1,
                                                                    tau0=1.0, D_scaler=1 / 256, T_scaler_m=7918.951990135471).to(dev).eval()
    N = lss.n_atoms
    cell32 = lss.cell.to(torch.float32)
    inv32 = torch.diag_embed(1.0 / torch.diagonal(lss.cell, dim1=1, dim2=2)).to(torch.float32)
    t_ms = [0.0]
    n_frames = [0]
    times = []

    def on_frames(frames, first_step):                      # frames[k, R, N, 3] fp64 on the device
        k = frames.shape[0]
        B = k * n_rep
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pos = frames.reshape(B * N, 3).to(torch.float32)
        Z = lss.Z.repeat(k)
        ptr = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=dev)
        T = torch.full((B,), 900.0, dtype=torch.float32, device=dev)
        defect = torch.full((B,), 1.0 / 256, dtype=torch.float32, device=dev)
        img = torch.zeros((B, 3), dtype=torch.int32, device=dev)
        t = tnet.time_device(Z, pos, cell32.repeat(k, 1, 1), inv32.repeat(k, 1, 1), img, ptr, N, T, defect, edge_capacity=B * N * 64)
        b.record()
        times.append((a, b, t))
        n_frames[0] += B

    for _ in range(2):                                      # warm-up with the timed run's flush size (sizes every workspace)
        lss.run(8, 900.0, mode="tks", sync_free=True, flush_every=8, on_frames=on_frames)
    times.clear(); n_frames[0] = 0
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    res = lss.run(steps, 900.0, mode="tks", sync_free=True, flush_every=8, on_frames=on_frames)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    tnet_ms = sum(a.elapsed_time(b) for a, b, _ in times)
    t_est = torch.cat([t for _, _, t in times]).cpu().numpy()
    n_q = n_rep * 12
    return {"what": "BASELINE config 4: TKS (alpha 1, beta 0, dqn false, T = 900 K) of 255-atom replicas, sync-free device loop "
                    "(enumerate -> expand -> K1..K5 -> Philox draw -> apply -> clock), t_net on every frame at each flush of 8 steps",
            "replicas": n_rep, "steps_timed": steps, "ms_per_step": wall * 1e3 / steps, "value": n_q * steps / wall,
            "unit": "Q-evals/s (wall clock, t_net included)", "tnet_frames": n_frames[0], "tnet_ms_per_step": tnet_ms / steps,
            "tnet_frames_per_s": n_frames[0] / (tnet_ms * 1e-3) if tnet_ms > 0 else None,
            "host_fraction_of_wall": res["host_seconds"] / res["wall_seconds"],
            "projected_seconds_for_10k_steps": wall / steps * 10000,
            "sim_time_after_sample_s": float(np.mean([res["time"][r][-1] for r in range(n_rep)])),
            "tnet_estimate_mean": float(t_est.mean()), "finite": bool(np.isfinite(t_est).all())}


def bench_config5(model_cpu, dev, rank: int, world: int, steps: int, warmup: int) -> dict:
    """BASELINE config 5: DQN replay minibatch (8 reaction graphs of 255-atom replicas = 16 PaiNN passes) forward + backward
    of all 614,447 parameters + clip_grad_norm_(0.8) + Adam (trainer.py:163-177), gradient all-reduce over NCCL at N > 1."""
    import torch
    import torch.distributed as dist
    import rlvacdiffsim_b200 as rb
    from rlvacdiffsim_b200.simulator import convert_to_graph_list
    graphs = []
    rng = np.random.RandomState(100 + rank)
    for r in range(8):
        s = rb.make_crconi_supercell(4, seed=9000 + rank * 8 + r)
        acts = rb.get_action_space(s, 3.528)
        g = convert_to_graph_list(s, [acts[int(rng.randint(12))]])[0]
        g.rl_q = torch.tensor(float(rng.normal(0.0, 0.1)))
        graphs.append(g)
    model = rb.registry.get_model_class("dqn_v2").load(CKPT)
    st = model.new_train_state(dev)
    eng = model._engine_for(dev)
    qp_train = dict(Q_PARAMS, beta=0.5)                                  # configs/train_dqn.toml:9-12
    host_batch = rb.Batch.from_data_list(graphs)
    batch = rb.batch_to(host_batch, dev)
    target = batch["rl_q"].reshape(-1).float()

    def fb():
        return model.train_forward_backward(batch, qp_train, st, target, mode="dqn")

    def sync_grads():
        if world > 1:
            dist.all_reduce(st["grads"], op=dist.ReduceOp.SUM)
            st["grads"].mul_(1.0 / world)

    def step():
        out = fb()
        sync_grads()
        model.optimizer_step(st, 1e-5, 0.8)
        return out

    for _ in range(max(warmup, 3)):
        step()
    l0 = eng.launch_count()
    ms_step, out = _time_calls(step, steps, dev)
    launches = (eng.launch_count() - l0) // steps
    ms_fb, _ = _time_calls(fb, steps, dev)
    ms_ar = 0.0
    if world > 1:
        ms_ar, _ = _time_calls(sync_grads, 50, dev)
    ms_adam, _ = _time_calls(lambda: model.optimizer_step(st, 0.0, 0.8), steps, dev)
    # the same step (fixed minibatch shape) as ONE CUDA graph: what is left when the 160 launches cost nothing on the host
    ms_graph = None
    try:
        prep = model.prepare_train_batch(batch, dev)
        cap = prep["n_nodes"] * 64

        def gstep():
            o = model.train_forward_backward_prepared(prep, qp_train, st, target, mode="dqn", edge_capacity=cap)
            sync_grads()
            model.optimizer_step(st, 1e-5, 0.8)
            return o
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                gstep()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if world == 1:                                  # NCCL collectives are left out of the capture
            step_saved = st["step"]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gout = gstep()
            st["step"] = step_saved + 1
            ms_graph, _ = _time_calls(lambda: g.replay(), steps, dev)
            assert eng.edge_overflow() == 0 and bool(torch.isfinite(gout["loss"]).item())
    except Exception as exc:
        ms_graph = f"{type(exc).__name__}: {exc}"

    def e2e_step():                                                     # batch_to (H2D of the minibatch) + step + loss.item() (D2H)
        b = rb.batch_to(host_batch, dev)
        o = model.train_forward_backward(b, qp_train, st, b["rl_q"].reshape(-1).float(), mode="dqn")
        sync_grads()
        model.optimizer_step(st, 1e-5, 0.8)
        return float(o["loss"].item())
    e2e_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = e2e_step()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([ms_step, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e = t.tolist()
    n_param = int(st["params"].numel())
    return {"what": "BASELINE config 5: replay minibatch of 8 reaction graphs (16 PaiNN passes, 4080 nodes) per rank: forward + backward "
                    "of all 614,447 parameters + clip_grad_norm_(0.8) + Adam(lr 1e-5), dropout off; gradients averaged with one NCCL "
                    "all-reduce of the flat 2.46 MB bucket per step when N > 1",
            "graphs_per_step_per_rank": 8, "ms_per_step": ms_step, "value": world * 8 / (ms_step * 1e-3), "unit": "reaction graphs/s (fwd+bwd+update)",
            "e2e_ms_per_step": ms_e2e, "e2e_value": world * 8 / (ms_e2e * 1e-3), "ms_forward_backward": ms_fb, "ms_allreduce": ms_ar,
            "ms_clip_adam": ms_adam, "ms_per_step_cuda_graph": ms_graph, "allreduce_bytes": 4 * n_param, "kernel_launches_per_step": int(launches), "loss": loss,
            "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in (host_batch["pos"], host_batch["pos_product"], host_batch["elems"], host_batch["cell"]))),
            "world": world}


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


def run_other_config(args, model, eng, dev, rank, world, local):
    """`--config 1|3|4|5`: that config's numbers as the JSON line (same contract keys as the headline line)."""
    import torch
    import torch.distributed as dist
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    l0 = eng.launch_count()
    if args.config == 1:
        obj = bench_config1(model, eng, dev, cpu=not args.no_cpu_baseline)
        value, e2e, ms, metric, unit = obj["value"], obj["e2e_value"], obj["ms_per_call_device"], "action Q-evals/sec (config 1: demo POSCAR_0, one call)", "Q-evals/s"
    elif args.config == 3:
        obj = bench_config3(model, eng, dev, cpu=not args.no_cpu_baseline)
        value, e2e, ms, metric, unit = obj["value"], obj["e2e_value"], obj["ms_per_step_device"], "action Q-evals/sec (config 3: 2040 atoms, 8 vacancies)", "Q-evals/s"
    elif args.config == 4:
        obj = bench_config4(model, eng, dev, rank, args.replicas_per_gpu, max(args.steps, 8))
        t = torch.tensor([obj["ms_per_step"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        value = world * args.replicas_per_gpu * 12 / (ms * 1e-3)
        e2e, metric, unit = value, "action Q-evals/sec (config 4: TKS steps + t_net on every frame)", "Q-evals/s"
    else:
        obj = bench_config5(model, dev, rank, world, max(args.steps, 10), args.warmup)
        value, e2e, ms, metric, unit = obj["value"], obj["e2e_value"], obj["ms_per_step"], "replay-minibatch reaction graphs/sec (config 5: fwd + bwd + clip + Adam, 8 graphs per rank)", "graphs/s"
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (fp16x3 split operands on tcgen05, fp32 accumulate; fp32 SIMT elsewhere)" if args.config != 5 else "f32",
                "data": "synthetic", "config": {"workload": obj["what"], "baseline_config": args.config},
                "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": obj.get("h2d_bytes_per_step"), "d2h_bytes_per_step": obj.get("d2h_bytes_per_step", 4)},
                "gpu_launches": int(launches + (obj.get("kernel_launches_per_step", 0) * args.steps if args.config == 5 else 0)),
                "clocks": clocks, "detail": obj}
        print(json.dumps(line))
    return 0


def _edge_ncu_note():
    """What one committed `ncu --set full` capture of the dominant kernel says about its shared-memory pipe (profiles/ncu_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("edge_stream_ncu")
    except (OSError, ValueError):
        return None


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
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE config whose numbers form the JSON line (default 2 = the headline; its line also carries "
                         "short runs of 1, 3, 4, 5 as extra objects)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config 1/3/4/5 objects of the default line")
    ap.add_argument("--streams", type=int, default=1, help="experiment: score the shard as this many independent sub-batches on "
                                                            "concurrent CUDA streams (one engine each), sync-free calls")
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
    if args.config != 2:
        rc = run_other_config(args, model, eng, dev, rank, world, local)
        if world > 1:
            dist.destroy_process_group()
        return rc
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

    # A5 + the per-step exchange of SURVEY.md §8e inside the timed step: softmax / draw per replica (rlvd_select_apply with
    # fixed uniforms, nothing applied: the benchmark state stays put), one fp32 record per replica packed on the device and,
    # at N > 1, ONE all_gather_into_tensor of those records over NCCL (57 KB at 1024 replicas)
    from rlvacdiffsim_b200.sharding import gather_step_device
    sel = None
    if not args.dedup:
        sel = {"first": torch.arange(0, (n_rep + 1) * N_CAND, N_CAND, dtype=torch.int32, device=dev),
               "rep_ptr": torch.arange(0, n_rep + 1, dtype=torch.int32, device=dev),
               "T": torch.full((n_rep,), Q_PARAMS["temperature"], dtype=torch.float32, device=dev),
               "u": torch.from_numpy(np.random.RandomState(7 + rank).random_sample(n_rep)).to(dev),
               "actions": torch.zeros((n_rep, N_CAND, 4), dtype=torch.float64, device=dev),
               "pos": torch.zeros((n_rep, 3), dtype=torch.float64, device=dev)}
    gather_events = []

    def step_device():
        out = eng.score_device(d["Z"], d["pos"], d["cell"], d["inv"], d["img"], d["ptr"], d["rs"], d["ps"], qp,
                               max_struct_nodes=int(np.diff(pk["node_ptr"]).max()))
        if sel is not None:
            chosen, _, _, _ = eng.select_apply(sel["rep_ptr"], sel["first"], out["rl_q"], sel["T"], sel["u"], sel["actions"],
                                               sel["pos"], False)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out["records"] = gather_step_device(eng, sel["first"], out["rl_q"], chosen, out["E_R"], rank, world, N_CAND)
            b.record()
            gather_events.append((a, b))
        return out

    # e2e inputs live in page-locked host memory before the timed region (as the contract asks): the C call DMA-copies them
    # straight to the device; cells go through cell_geometry (cached while the boxes do not change, as in LSS / TKS)
    hp = eng.pin_arrays({"Z": pk["Z"].astype(np.int32), "pos": pk["pos"].astype(np.float32), "cells": pk["cells"],
                         "node_ptr": pk["node_ptr"].astype(np.int32), "rs": pk["react_struct"].astype(np.int32),
                         "ps": pk["prod_struct"].astype(np.int32)})

    def step_host():
        return eng.score_host(hp["Z"], hp["pos"], hp["cells"], hp["node_ptr"], hp["rs"], hp["ps"], qp)

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

    if args.streams > 1:
        # ---- experiment: the shard as `streams` independent sub-batches, each on its own stream and engine, queued without
        # host synchronisation — lets the hardware run one sub-batch's (issue / LSU-bound) edge kernel next to another's
        # (HBM-bound) node kernels
        from rlvacdiffsim_b200.engine import Engine
        ns = args.streams
        engines = [eng] + [Engine(dev) for _ in range(ns - 1)]
        for e2 in engines[1:]:
            sd, hp_ = model._engine_state()
            e2.load_state_dict(sd, hp_)
        streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
        per = n_rep // ns
        subs = []
        for k in range(ns):
            pk_k = pack_replicas(reps[k * per:(k + 1) * per], spaces[k * per:(k + 1) * per])
            subs.append(_device_batch(pk_k, eng, dev))
        caps = [int(sb["pos"].shape[0]) * 64 for sb in subs]
        qps = [e2.q_params(Q_PARAMS, model.head_spec) for e2 in engines]

        def step_streams():
            outs = []
            cur = torch.cuda.current_stream(dev)
            for k in range(ns):
                streams[k].wait_stream(cur)
                with torch.cuda.stream(streams[k]):
                    sb = subs[k]
                    outs.append(engines[k].score_device(sb["Z"], sb["pos"], sb["cell"], sb["inv"], sb["img"], sb["ptr"], sb["rs"],
                                                        sb["ps"], qps[k], max_struct_nodes=sb["max_nodes"], edge_capacity=caps[k]))
            for k in range(ns):
                cur.wait_stream(streams[k])
            return outs
        for _ in range(3):
            outs = step_streams()
        torch.cuda.synchronize(dev)
        ms_s, outs = timed(step_streams, args.steps)
        q_cat = torch.cat([o["rl_q"] for o in outs]).cpu().numpy()
        same = bool(np.array_equal(q_cat, res["rl_q"].cpu().numpy()[: len(q_cat)]))
        print(json.dumps({"experiment": "streams", "streams": ns, "ms_per_step": ms_s / args.steps, "rl_q_equal_to_single_stream": same,
                          "overflow": [e2.edge_overflow() for e2 in engines]}))

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    launches0 = eng.launch_count()
    gather_events.clear()
    ms_dev, res = timed(step_device, args.steps)
    launches = eng.launch_count() - launches0
    gather_us = 1e3 * float(np.mean([a.elapsed_time(b) for a, b in gather_events])) if gather_events else None
    if sel is not None:                                       # every rank holds every replica's record after the gather
        rec = res["records"].cpu().numpy()
        assert rec.shape == (world * n_rep, 4 + N_CAND) and sorted(rec[:, 0].astype(int).tolist()) == list(range(world * n_rep))
        mine = rec[rec[:, 0].astype(int) % world == rank]
        assert np.array_equal(mine[np.argsort(mine[:, 0]), 4:].reshape(-1), res["rl_q"].cpu().numpy())
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

    # the same call frozen into a CUDA graph (Engine.capture_score): one launch of ~30 kernel nodes instead of 30 launches
    ms_one_graph = graph_err = None
    try:
        T1 = torch.full((len(one["react_struct"]),), Q_PARAMS["temperature"], dtype=torch.float32, device=dev)
        gs = eng.capture_score(d1["Z"], d1["pos"], d1["cell"], d1["inv"], d1["img"], d1["ptr"], d1["rs"], d1["ps"], qp, T1, mx1)
        for _ in range(5):
            gs()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(50):
            q_g = gs(pos=d1["pos"])["rl_q"].cpu()
        ms_one_graph = (time.perf_counter() - t0) * 1e3 / 50
        graph_err = float((q_g - step_one()["rl_q"].cpu()).abs().max())
        assert eng.edge_overflow() == 0
    except Exception as exc:
        graph_err = f"{type(exc).__name__}: {exc}"

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    # short runs of the other BASELINE configs (N = 1 only; `--config K` gives each its own full line, also at N > 1)
    extra_cfg = {}
    if world == 1 and not args.no_extra_configs and not args.dedup:
        for name, fn in (("1", lambda: bench_config1(model, eng, dev, cpu=not args.no_cpu_baseline)),
                         ("3", lambda: bench_config3(model, eng, dev, cpu=False)),
                         ("4", lambda: bench_config4(model, eng, dev, rank, n_rep, 16)),
                         ("5", lambda: bench_config5(model, dev, rank, world, 20, 3))):
            try:
                extra_cfg[name] = fn()
            except Exception as exc:                      # an extra object must never take the headline down
                extra_cfg[name] = {"error": f"{type(exc).__name__}: {exc}"}

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
        # per-kernel-family rulers (north_star: HBM GB/s for neighbor-list / embed / scatter, tensor-pipe for channel mixing).
        # Algorithmic bytes per SURVEY.md §8(d): K1 12N + 4(N+1) + 7E; fused edge kernel 40 F N + 7E per structure-layer
        # (two launches: layers 1, 2; layer 0 runs as per-node GEMMs inside the node family); node family per layer: context
        # net 16 F N (read q, write x) + mixing 32 F N (read + write q, mu); readout 2 N F 4 per reaction.
        ms_k = {k: fam[k][0] / max(prof_steps, 1) for k in fam}
        bf16_peak = float(peaks.get("bf16_tflops_sustained", 1383.0))
        alg = {"neighbor": 12.0 * M + 4.0 * (M + n_struct) + 7.0 * n_edges,
               "edge_pack": 7.0 * n_edges + 12.0 * M,
               "edge": 2 * edge_bytes,
               "node": L * 48.0 * F * M,
               "readout": 2.0 * 4 * F * M}
        node_flops = (100e6 + 150e6 + 125e6) / 255.0 * M          # interatomic + mu_channel_mix + intra-atomic GEMMs
        rooflines = {}
        for k in alg:
            if ms_k[k] > 0:
                gbs = alg[k] / (ms_k[k] * 1e-3) / 1e9
                rooflines[k] = {"bound": "hbm", "algorithmic_GB_per_step": alg[k] / 1e9, "ms_per_step": ms_k[k], "achieved_GB/s": gbs,
                                "frac": gbs / hbm_peak}
        if ms_k["node"] > 0:
            tf = node_flops / (ms_k["node"] * 1e-3) / 1e12
            rooflines["node_k3b_tensor"] = {"bound": "tensor", "what": "channel-mixing GEMMs (K3b) over the whole node family's time",
                                            "useful_TFLOP_per_step": node_flops / 1e12, "achieved_TFLOP/s": tf,
                                            "dispatched_TFLOP/s": 3 * tf, "split_factor": 3, "peak": bf16_peak,
                                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (fp16 and bf16 share the rate)",
                                            "frac_useful": tf / bf16_peak, "frac_dispatched": 3 * tf / bf16_peak}
        if ms_k["edge"] > 0:
            ef = 2.0 * 2 * 3 * F * (NRBF + 1) * n_edges                   # filter evaluation, layers 1 and 2
            tf = ef / (ms_k["edge"] * 1e-3) / 1e12
            rooflines["edge_k3a_tensor"] = {"bound": "tensor", "what": "filter_net evaluation on tcgen05 inside edge_stream_kernel",
                                            "achieved_TFLOP/s": tf, "dispatched_TFLOP/s": 3 * tf * 64.0 / 63.0, "peak": bf16_peak,
                                            "frac_useful": tf / bf16_peak}
        value = world * n_q * args.steps / (ms_dev * 1e-3)
        e2e = world * n_q * args.steps / (ms_e2e * 1e-3)
        h2d = sum(int(np.asarray(pk[k]).nbytes) for k in ("Z", "pos", "node_ptr", "react_struct", "prod_struct")) + \
            int(pk["cells"].astype(np.float32).nbytes) + int(inv.nbytes) + int(img.nbytes)
        line = {
            "metric": "action Q-evals/sec", "value": value, "unit": "Q-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16x3 split operands on tcgen05, fp32 accumulate; fp32 SIMT elsewhere)", "data": "synthetic",
            "config": {"workload": "BASELINE config 2 (1024 independent 256-site CrCoNi replicas, 255 atoms, 12 "
                                   "candidate hops each, LSS scoring) sharded 128 replicas per GPU",
                       "replicas_per_gpu": n_rep, "q_evals_per_step_per_gpu": n_q, "structures_per_step_per_gpu": n_struct,
                       "nodes": M, "edges": n_edges, "reactant_dedup": bool(args.dedup), "jitter": args.jitter,
                       "l2": "512 MB flush buffer written between timed steps",
                       "q_params": Q_PARAMS, "parallelism": f"replica-sharded x{world}; no data-path collective, one all_gather of per-replica records per step"},
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
                                  "peak_source": "128 B/clk/SM x SMs x sampled SM clock",
                                  "ncu": _edge_ncu_note()}},
            "dram_by_family": dram_by_family,
            "dedup": dedup_obj,
            "single_replica": {"what": "one 255-atom replica, 12 candidate hops per call (24 structures), device-resident inputs, "
                                       "Q-values read back to the host every call (rank 0 wall clock)",
                               "ms_per_call": ms_one, "value": len(one["react_struct"]) / (ms_one * 1e-3), "unit": "Q-evals/s",
                               "cuda_graph": {"what": "the same call replayed as ONE CUDA graph (Engine.capture_score), positions copied "
                                                      "in and Q read back every call", "ms_per_call": ms_one_graph,
                                              "value": (len(one["react_struct"]) / (ms_one_graph * 1e-3)) if ms_one_graph else None,
                                              "max_abs_diff_vs_eager_rl_q": graph_err}},
            "device_lss": {"what": "one full LSS step of every replica on the device: rlvd_action_space -> expand -> K1..K5 -> "
                                   "softmax/draw -> apply, fp64 state resident in HBM, one 8-byte D2H per step",
                           "ms_per_step": ms_lss, "value": world * n_q / (ms_lss * 1e-3), "unit": "Q-evals/s (rank 0 wall clock)",
                           "host_front_end_ms_per_step": ms_host_front,
                           "host_front_end": "reference-shaped per-replica numpy enumeration + packing for the same shard "
                                             f"(measured on {n_front} replicas, scaled)"},
            "kernel_share": {k: (v[0] / total_prof if total_prof > 0 else None) for k, v in fam.items()},
            "kernel_ms_per_step": {k: v[0] / prof_steps for k, v in fam.items()},
            "gather": {"what": "per-step exchange inside the timed step: rlvd_pack_step_records + (N > 1) one NCCL "
                               "all_gather_into_tensor of [replica id, n_actions, chosen, E_R, Q[12]] per replica",
                       "us_per_step": gather_us, "bytes_per_rank": 4 * (4 + N_CAND) * n_rep, "world": world},
            "rooflines": rooflines,
        }
        if extra_cfg:
            line["configs"] = extra_cfg
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
