"""CPU: the host-side mirror of the reference interface — no GPU compute."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT
import rlvacdiffsim_b200 as rb
from rlvacdiffsim_b200 import _lib
from rlvacdiffsim_b200.engine import cell_geometry
from rlvacdiffsim_b200.simulator import (LatticeEnv, RLSimulator, ThermalAnnealing, convert_to_graph_list,
                                         pack_replicas)
from rlvacdiffsim_b200.structures import append_xdatcar, apply_action, make_crconi_supercell, read_poscar, write_poscar
from oracle.painn_oracle import image_range, inverse_cell_f32


def test_registry_load_save_roundtrip(tmp_path, checkpoint_path):
    model = rb.registry.get_model_class("dqn_v2").load(checkpoint_path)           # deploy.py:18
    ck = torch.load(checkpoint_path, weights_only=True, map_location="cpu")
    assert set(model.state_dict().keys()) == set(ck["state_dict"].keys())
    assert all(torch.equal(model.state_dict()[k], v) for k, v in ck["state_dict"].items())
    assert sum(p.numel() for p in model.parameters()) + 21 == 614447              # SURVEY §2 (incl. freqs+cutoff)
    assert any("reaction_model" in n for n, _ in model.named_parameters())        # trainer.py:36-39
    out = tmp_path / "m.pth.tar"
    model.save(str(out))
    again = rb.registry.get_model_class("dqn_v2").load(str(out))
    assert all(torch.equal(again.state_dict()[k], v) for k, v in ck["state_dict"].items())
    assert rb.registry.get_reaction_model_class("painn") is type(model.reaction_model)
    with pytest.raises(KeyError):
        rb.registry.get_model_class("nope")


def test_forward_refuses_cpu(checkpoint_path):
    model = rb.registry.get_model_class("dqn_v2").load(checkpoint_path).eval()
    s = make_crconi_supercell(3, seed=0)
    acts = rb.get_action_space(s, 3.528)
    batch = next(iter(rb.DataLoader(rb.ReactionDataset(convert_to_graph_list(s, acts)), batch_size=16)))
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        model(batch, q_params={"alpha": 0.0, "beta": 1.0, "dqn": True, "temperature": 700})


def test_graph_containers_and_loader():
    s = make_crconi_supercell(3, seed=1)
    acts = rb.get_action_space(s, 3.528)
    graphs = convert_to_graph_list(s, acts)
    assert len(graphs) == 12 and graphs[0]["elems"].dtype == torch.int64
    moved = int(acts[0][0])
    diff = (graphs[0]["pos_product"] - graphs[0]["pos"]).abs().sum(1)
    assert torch.nonzero(diff).flatten().tolist() == [moved]
    graphs[0].rl_q = torch.tensor(0.5)                                           # trainer.py:157
    for g in graphs[1:]:
        g.rl_q = torch.tensor(0.0)
    loader = rb.DataLoader(rb.ReactionDataset(graphs), batch_size=8)              # trainer.py:160
    batches = list(loader)
    assert [b.num_graphs for b in batches] == [8, 4] and len(loader) == 2
    b = rb.batch_to(batches[0], "cpu")
    assert b["pos"].shape == (8 * 107, 3) and b["batch"].max().item() == 7
    assert b["ptr"].tolist() == [107 * i for i in range(9)]
    assert b["rl_q"].shape == (8,) and b["rl_q"][0].item() == 0.5
    ag = rb.AtomsGraph.from_ase(s, 5.0, read_properties=False, neighborlist_backend="ase", add_batch=True)
    ag.T = torch.as_tensor(900.0)                                                 # estimate_time.py:47-52
    ag.defect = torch.as_tensor(1 / 256)
    tb = next(iter(rb.DataLoader(rb.AtomsDataset([ag, ag]), batch_size=2)))
    assert tb["T"].tolist() == [900.0, 900.0]


def test_cell_geometry_matches_oracle_definitions():
    for atoms in (read_poscar(os.path.join(GOLD, "POSCAR_0")), make_crconi_supercell(2, seed=0)):
        inv, img = cell_geometry(atoms.cell[None], 5.0)
        assert np.array_equal(inv[0], inverse_cell_f32(atoms.cell))
        assert np.array_equal(img[0], image_range(atoms.cell, 5.0))


def test_pack_replicas_layouts():
    reps = [make_crconi_supercell(3, seed=s) for s in (0, 1)]
    spaces = [rb.get_action_space(r, 3.528) for r in reps]
    full = pack_replicas(reps, spaces, dedup_reactants=False)
    assert len(full["node_ptr"]) == 2 * 24 + 1 and len(full["react_struct"]) == 24
    assert np.array_equal(full["prod_struct"], full["react_struct"] + 1)
    dd = pack_replicas(reps, spaces, dedup_reactants=True)
    assert len(dd["node_ptr"]) == 2 * 13 + 1
    assert sorted(set(dd["react_struct"].tolist())) == [0, 13]
    a = spaces[1][5]
    p = dd["prod_struct"][12 + 5]
    lo = dd["node_ptr"][p]
    np.testing.assert_allclose(dd["pos"][lo + int(a[0])], (reps[1].positions[int(a[0])] + np.array(a[1:])).astype(np.float32))


def test_poscar_and_xdatcar_roundtrip(tmp_path):
    s = make_crconi_supercell(3, seed=4, jitter=0.02)
    p = tmp_path / "POSCAR"
    write_poscar(str(p), s)
    back = read_poscar(str(p))
    first = list(dict.fromkeys(s.numbers.tolist()))                    # species blocks in order of first appearance
    order = np.concatenate([np.nonzero(s.numbers == z)[0] for z in first])
    np.testing.assert_allclose(back.positions, s.positions[order], atol=1e-12)
    assert np.array_equal(back.numbers, s.numbers[order])
    x = tmp_path / "XDATCAR"
    append_xdatcar(str(x), s, 0, append=False)
    append_xdatcar(str(x), apply_action(s, [0, 0.1, 0.0, 0.0]), 1)
    assert open(x).read().count("Direct configuration=") == 2


def test_simulator_surface(checkpoint_path):
    model = rb.registry.get_model_class("dqn_v2").load(checkpoint_path)
    env = LatticeEnv(make_crconi_supercell(3, seed=0), {"device": "cuda"})
    sim = RLSimulator(env, model, q_params={"alpha": 0.0, "beta": 1.0, "dqn": True}, lattice_parameter=3.528)
    sim.update_q_params(temperature=700)
    assert sim.q_params["temperature"] == 700 and sim.kb == pytest.approx(8.617e-5)
    sched = ThermalAnnealing(total_horizon=20, annealing_time=10, T_start=1200, T_end=700)   # demo/config.toml
    assert sched.get_temperature(0) == 1200 and sched.get_temperature(10) == 700 and sched.get_temperature(19) == 700
    before = env.atoms.positions.copy()
    act = sim._action_space()[0]
    env.step(act)
    assert np.abs(env.atoms.positions - before).sum() == pytest.approx(np.abs(np.array(act[1:])).sum())


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rlvd_b200.h")).read()
    declared = set(re.findall(r"\b(rlvd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()                      # raises if the .so is missing; no compute without a GPU
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rlvd_abi_version() == 2


def test_t_net_surface_and_oracle_head(checkpoint_path):
    """rlsim/time/train.py:44-48: registry.get_model_class("t_net").load_representation(dqn.reaction_model, **cfg)."""
    import torch
    from oracle.painn_oracle import PaiNNOracle, TIME_HEAD_SPEC_V0
    dqn = rb.registry.get_model_class("dqn_v2").load(checkpoint_path)
    tnet = rb.registry.get_model_class("t_net").load_representation(dqn.reaction_model, n_feat=32, tau0=12.0)
    assert tnet.tau == 12.0 and tnet.tau0 == 12.0 and tnet.cutoff == 5.0
    assert tnet.representation is dqn.reaction_model.representation          # shared weights, same kernels
    assert {"time_output.layers.0.weight", "time_output.layers.6.bias"} <= set(tnet.state_dict())
    g = rb.AtomsGraph.from_ase(make_crconi_supercell(2, seed=1), 5.0)
    g.T, g.defect = torch.tensor(900.0), torch.tensor(1 / 32)
    batch = next(iter(rb.DataLoader(rb.AtomsDataset([g, g]), batch_size=8)))
    assert batch["T"].shape == (2,) and batch.num_graphs == 2
    with pytest.raises(RuntimeError):
        tnet.eval()(batch, inference=True)                                    # CPU batch: no fallback
    # the oracle head is monotone in its logit and bounded by tau
    s = make_crconi_supercell(2, seed=1)
    m = PaiNNOracle.load(checkpoint_path, dtype=torch.float64)
    head = {"w0": np.zeros((4, 130)), "b0": np.zeros(4), "w1": np.zeros((4, 4)), "b1": np.zeros(4),
            "w2": np.zeros((1, 4)), "b2": np.array([0.0])}
    t = m.time_forward(s.numbers, s.positions.astype(np.float32), np.asarray(s.cell)[None], np.array([0, len(s)]),
                       [900.0], [1 / 32], head, tau=12.0)
    assert TIME_HEAD_SPEC_V0.squash == "tau * sigmoid" and float(t[0]) == pytest.approx(6.0)


def test_sro_host_mirror_properties():
    """rlsim/utils/sro.py:49-83 restated in numpy: symmetric, near zero for a random equiatomic alloy, and the
    pair bookkeeping sums to the 12-neighbour first shell."""
    from rlvacdiffsim_b200.sro import _sro_from_counts, get_sro_from_atoms
    s = make_crconi_supercell(4, seed=0, jitter=0.02)
    m = get_sro_from_atoms(s)
    assert m.shape == (3, 3) and np.allclose(m, m.T) and np.abs(m).max() < 0.1
    # a perfectly ordered toy: only unlike first-shell pairs -> alpha_AA = 1, alpha_AB = 1 - 1/(2 * 0.5 * 0.5) = -1
    counts = np.array([[0.0, 48.0], [48.0, 0.0]])
    o = _sro_from_counts(counts, np.array([4.0, 4.0]))
    assert o[0, 0] == 1.0 and o[1, 1] == 1.0 and o[0, 1] == pytest.approx(-1.0)


def test_trajectory_writer_matches_per_step_appends(tmp_path):
    """Buffered, batched XDATCAR / JSON writers produce byte-for-byte what the reference-shaped per-step
    appends do (simulator.py:175,215), across flush boundaries."""
    import json

    from rlvacdiffsim_b200.trajectory import TrajectoryWriter
    rng = np.random.default_rng(3)
    reps = [make_crconi_supercell(2, seed=s) for s in range(3)]
    numbers = np.stack([r.get_atomic_numbers() for r in reps])
    cells = np.stack([r.get_cell() for r in reps])
    w = TrajectoryWriter(str(tmp_path / "buf"), numbers, cells, flush_every=4)
    T = 10
    qs, acts = [], []
    for t in range(T):
        for r in reps:
            r.set_positions(r.get_positions() + rng.normal(0, 0.05, r.get_positions().shape))
        pos = np.stack([r.get_positions() for r in reps])
        q = [rng.normal(size=rng.integers(3, 9)) for _ in reps]
        a = [int(rng.integers(0, len(x))) for x in q]
        qs.append(q)
        acts.append(a)
        w.append(pos, q_values=q, chosen=a)
        for i, r in enumerate(reps):
            append_xdatcar(str(tmp_path / f"ref{i}"), r, t, append=t > 0)
    w.close()
    for i in range(3):
        assert open(w.xdatcar_path(i)).read() == open(tmp_path / f"ref{i}").read()
    ql = json.load(open(tmp_path / "buf" / "q_values.json"))
    assert len(ql) == 3 and len(ql[0]) == T and ql[1][4] == qs[4][1].tolist()
    assert json.load(open(tmp_path / "buf" / "actions.json"))[2] == [acts[t][2] for t in range(T)]


def test_trainer_surface_and_memory_roundtrip(tmp_path, checkpoint_path):
    """rlsim/drl/trainer.py / rlsim/memory.py mirror: constructor semantics (train_all freezes reaction_model), the JSON
    layout of Memory.save / load, and a loud failure without a GPU."""
    from rlvacdiffsim_b200.structures import apply_action
    from rlvacdiffsim_b200.trainer import Memory, Trainer
    model = rb.registry.get_model_class("dqn_v2").load(checkpoint_path)
    tr = Trainer(model, {"alpha": 0.0, "beta": 1.0, "dqn": True}, train_all=False)
    frozen = [n for n, p in model.named_parameters() if not p.requires_grad]
    assert frozen and all("reaction_model" in n for n in frozen)
    assert all(p.requires_grad for n, p in model.named_parameters() if "reaction_model" not in n)
    s = make_crconi_supercell(2, seed=1)
    space = rb.get_action_space(s, 3.528)
    mem = Memory(alpha=1.0, beta=0.0, T=900.0)
    mem.add({"state": s, "act": 1, "act_space": space, "act_probs": [0.1], "fail": False, "next": apply_action(s, space[1]),
             "log_freq": 2.0, "E_s": 1.5, "E_min": 1.0, "E_next": 0.9})
    assert abs(mem.rewards[0] - (-(0.5) + mem.kb * 900.0 * 2.0)) < 1e-12          # memory.py:58-62
    mem.save(str(tmp_path / "traj0"))
    back = Memory(alpha=0, beta=0)
    back.load(str(tmp_path / "traj0.json"))
    assert back.actions == [1] and back.rewards == mem.rewards and len(back.act_space[0]) == len(space)
    np.testing.assert_array_equal(back.states[0].get_positions(), s.get_positions())
    with pytest.raises(RuntimeError):                  # no CUDA device here: the whole-model update fails loudly, no CPU path
        Trainer(model, {"alpha": 0.0, "beta": 1.0, "dqn": True}, train_all=True).dqn_update([mem, mem], gamma=0.8, episode_size=1,
                                                                                      num_epoch=1, device="cuda")


def test_train_layout_is_the_checkpoint_order(checkpoint_path):
    """rlvd_train_layout (no GPU needed): the flat parameter vector is every floating-point tensor of the checkpoint in
    state_dict order, 614,447 floats, buffers marked non-trainable."""
    import torch
    from rlvacdiffsim_b200.engine import Engine
    ck = torch.load(checkpoint_path, weights_only=True, map_location="cpu")["state_dict"]
    want = [(k, v.numel()) for k, v in ck.items() if v.is_floating_point()]
    tab = Engine.train_layout()
    assert [(n, c) for n, _, c, _ in tab] == want
    off = 0
    for name, o, c, tr in tab:
        assert o == off
        off += c
        assert tr == (not name.endswith(("cutoff_fn.cutoff", "radial_basis.freqs")))
    assert off == 614447


def test_time_estimator_glue(tmp_path, checkpoint_path):
    """rlsim/cli/scripts/estimate_time.py mirror (no GPU): timer_converter's formula and error, the XDATCAR reader against the
    writer, and the t_net checkpoint round trip in the reference's {"state_dict", "hyper_parameters"} format."""
    import torch
    from rlvacdiffsim_b200.structures import _species_blocks, append_xdatcar, apply_action
    from rlvacdiffsim_b200.timeest import read_xdatcar, timer_converter
    t = torch.tensor([0.0, 0.25, 0.9])
    np.testing.assert_allclose(timer_converter(t, 1.0).numpy(), -np.log(1 - t.numpy()), rtol=1e-6)
    with pytest.raises(ValueError, match="exceeds the tau"):
        timer_converter(torch.tensor([0.5, 1.5]), 1.0)                                   # estimate_time.py:26-29
    s = make_crconi_supercell(2, seed=3)
    frames = [s]
    for k in range(4):
        frames.append(apply_action(frames[-1], rb.get_action_space(frames[-1], 3.528)[k]))
    path = str(tmp_path / "XDATCAR0")
    for k, f in enumerate(frames):
        append_xdatcar(path, f, k, append=k > 0)
    back = read_xdatcar(path)
    assert len(back) == 5 and len(read_xdatcar(path, skip=2)) == 3
    for f, b in zip(frames, back):
        order = _species_blocks(f)[2]                                                   # XDATCAR groups atoms by species, first seen first
        assert np.array_equal(b.get_atomic_numbers(), np.asarray(f.get_atomic_numbers())[order])
        d = b.get_positions() - (f.get_positions()[order] % np.diag(f.get_cell()))
        d -= np.round(d / np.diag(f.get_cell())) * np.diag(f.get_cell())
        assert np.abs(d).max() < 2e-7 * 7.1                                             # 8 decimals of fractional coordinates
    dqn = rb.registry.get_model_class("dqn_v2").load(checkpoint_path)
    torch.manual_seed(0)
    tnet = rb.registry.get_model_class("t_net").load_representation(dqn.reaction_model, n_feat=64, pooling="mean", tau0=1.0)
    tnet.tau = 2.5
    tnet.save(str(tmp_path / "checkpoint10.pth.tar"))
    again = rb.registry.get_model_class("t_net").load(str(tmp_path / "checkpoint10.pth.tar"))
    assert again.tau == 2.5 and again.hyper["n_feat"] == 64
    for (k1, v1), (k2, v2) in zip(tnet.state_dict().items(), again.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    # t_net_binary (time/train.py:42-48): same constructor, a second head for the "goal" logit, same checkpoint format
    tb = rb.registry.get_model_class("t_net_binary").load_representation(dqn.reaction_model, n_feat=32, pooling="sum", tau0=2.0)
    assert tb.HAS_GOAL and any(k.startswith("goal_output.") for k in tb.state_dict())
    tb.save(str(tmp_path / "tb.pth.tar"))
    tb2 = rb.registry.get_model_class("t_net_binary").load(str(tmp_path / "tb.pth.tar"))
    assert tb2.hyper["@name"] == "t_net_binary" and tb2.hyper["pooling"] == "sum"
    for (k1, v1), (k2, v2) in zip(tb.state_dict().items(), tb2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
