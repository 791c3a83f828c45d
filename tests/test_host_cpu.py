"""CPU: host-side logic — C-ABI surface, loader behaviour, module tree / state_dict contract, batching."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import fvgn_b200 as fv
from oracle import fvgn_oracle as O
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fvgn_b200.h")).read()
    declared = sorted(set(re.findall(r"FVGN_API\s+[\w\s\*]+?\b(fvgn_\w+)\s*\(", hdr)))
    assert declared == fv._lib.declared_symbols(), "ctypes table and header disagree"
    assert os.path.exists(fv._lib.LIB_PATH), "build the library first: python __graft_entry__.py"
    L = ctypes.CDLL(fv._lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert L.fvgn_abi_version() == 11
    # size queries are pure host code and must work without a device
    L.fvgn_norm_workspace_floats.restype = ctypes.c_size_t
    assert L.fvgn_norm_workspace_floats(14) > 0


def test_ops_reject_cpu_tensors_loudly():
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        fv.ops.build_csr(torch.zeros(3, dtype=torch.int32), 2)
    m = fv.build_mlp(14, 128, 128, drop_out=False)
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 14))


def test_state_dict_contract_and_seeded_init():
    torch.manual_seed(0)
    m = fv.FVGN(device="cpu", params=H.params())
    sd = m.state_dict()
    ref = H.torch_reference_state_dict(0)
    assert len(sd) == 283 and sum(p.numel() for p in m.parameters()) == 2210695
    assert sorted(sd.keys()) == sorted(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype and torch.equal(sd[k], ref[k]), k
    assert sd["_cell_normalizer.acc_sum"].tolist() == [1.0, 1.0]  # buffers start at ONE (App. D.1)


def test_checkpoint_roundtrip(tmp_path):
    m = fv.FVGN(device="cpu", params=H.params())
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    path = str(tmp_path / "0.state")
    m.save_checkpoint(path=path, optimizer=opt, scheduler=None, trian_time_steps=np.arange(3))
    d = torch.load(path, weights_only=False)
    assert set(d.keys()) >= {"model", "trian_time_steps", "optimizer0"}
    m2 = fv.FVGN(device="cpu", params=H.params())
    steps = m2.load_checkpoint(optimizer=torch.optim.AdamW(m2.parameters(), lr=1e-3), ckpdir=path, device="cpu")
    assert np.array_equal(steps, np.arange(3))
    for a, b in zip(m.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)


def test_unsupported_variants_raise():
    with pytest.raises(NotImplementedError):
        fv.build_mlp(4, 128, 128, drop_out=False, nn_act="ReLU")
    with pytest.raises(NotImplementedError):
        fv.EdgeBlock(custom_func=None, dual_edge=True)
    with pytest.raises(ValueError):
        fv.CellBlock(128, 128, MultiHead=3)


def test_batching_matches_pyg_rule():
    ms = [fv.synthetic.make("tiny", seed=s, T=3) for s in (1, 2)]
    tabs = [O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"]) for m in ms]
    mine = fv.dataset.collate([fv.dataset.build_graphs(t, m["velocity"], m["pressure"], training_pair=0) for t, m in zip(tabs, ms)])
    theirs = [O.batch_graphs([O.build_graphs(t, m["velocity"], m["pressure"], training_pair=0)[k] for t, m in zip(tabs, ms)]) for k in range(3)]
    for a, b in zip(mine, theirs):
        for k in b.keys:
            va, vb = getattr(a, k), getattr(b, k)
            if torch.is_tensor(vb):
                assert torch.equal(va.reshape(vb.shape).to(vb.dtype), vb), k


def test_synthetic_meshes_shapes():
    cyl = fv.synthetic.make("cyl", seed=0)
    assert 1700 < cyl["mesh_pos"].shape[0] < 2100 and 3200 < cyl["cells"].shape[0] < 4000
    t = O.build_connectivity(cyl["mesh_pos"], cyl["cells"], cyl["node_type"])
    assert cyl["mesh_pos"].shape[0] - t["face"].shape[1] + cyl["cells"].shape[0] == 0  # one hole
    assert (t["cells_area"] > 0).all()


def test_fused_optimizer_step_invalidates_packed_weights():
    """torch's fused optimizers update parameters without bumping Tensor._version; the pack key of the tensor-core weight tiles must
    change anyway (process-wide optimizer-step hook), or a training loop would run its forward on stale packed weights"""
    import torch

    from fvgn_b200 import ops

    ws = [torch.nn.Parameter(torch.ones(4, 4)) for _ in range(6)]
    for w in ws:
        w.grad = torch.ones_like(w)
    k0 = ops._weights_key(ws)
    try:
        opt = torch.optim.AdamW(ws, lr=1e-3, fused=True)
    except (RuntimeError, TypeError):
        opt = torch.optim.AdamW(ws, lr=1e-3)
    opt.step()
    assert ops._weights_key(ws) != k0
    k1 = ops._weights_key(ws)
    with torch.no_grad():
        ws[0].mul_(2.0)  # ordinary in-place update: caught by the version counter
    assert ops._weights_key(ws) != k1


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: the unmodified reference via oracle/_ref where oracle/make_ref.py has placed it, else the
    oracle port) prints ONE JSON line with the contract's keys, and its `config` is the native arm's"""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_sec" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0 and d["config"]["workload"] == "tiny"
    from oracle import ref_loader as RL

    assert d["cpu_baseline"]["kind"] == ("reference" if RL.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    sys.path.insert(0, root)
    import bench

    c = d["config"]
    assert c == bench.config_of("tiny", c["cells_per_rank"], c["faces_per_rank"], c["nodes_per_rank"], c["graphs_per_rank"], 1)
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_sass_of_the_tensor_core_kernels_is_blackwell_native():
    """cuobjdump -sass of the built library: the fused MLP kernels issue tcgen05.mma (UTCHMMA), read / write TMEM (LDTM / STTM), stream their
    weight tiles with cp.async.bulk (UBLKCP) and use mbarriers (SYNCS); no legacy HMMA anywhere; the strict path is FFMA.  (The mnemonics of
    /opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel"; skipped where cuobjdump is not installed.)"""
    import re
    import shutil
    import subprocess

    import fvgn_b200 as fv

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        sass = subprocess.run([exe, "-sass", fv._lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    if "Function :" not in sass:
        pytest.skip("cuobjdump printed no SASS")
    per, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = set()
        elif cur:
            per[cur].update(re.findall(r"\b(UTCHMMA\.2CTA|UTCBAR\.2CTA\.MULTICAST|UCGABAR_ARV|UTCHMMA|LDTM|STTM|UBLKCP|LDGSTS|SYNCS|HMMA|FFMA)\b", line))
            per[cur].update(re.findall(r"\.(A_KEEP|A_REUSE)\b", line))
    assert not any("HMMA" in v for v in per.values()), "legacy mma.sync / wmma path found"
    want = {"fused_mlp_tc_kernel": {"UTCHMMA", "LDTM", "STTM", "UBLKCP", "SYNCS"}, "fused_mlp_tc3_kernel": {"UTCHMMA", "LDTM", "STTM", "UBLKCP", "SYNCS"},
            "dgrad_tc3_kernel": {"UTCHMMA", "LDTM", "STTM", "UBLKCP", "SYNCS"}, "wgrad_tc3_kernel": {"UTCHMMA", "LDTM", "SYNCS"},
            "fused_mlp_simt_kernel": {"FFMA"}}
    for name, need in want.items():
        ks = [k for k in per if name in k]
        assert ks, name
        for k in ks:
            assert need <= per[k], (k, need - per[k])
    # round 2: the CTA-pair inference kernels issue the 2-CTA MMA (tcgen05.mma.cta_group::2), commit with a cluster multicast and synchronise
    # the pair with the cluster barrier
    pair = [k for k in per if "fused_mlp_tc3_pair_kernel" in k]
    assert len(pair) == 2, pair
    for k in pair:
        assert {"UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UCGABAR_ARV", "LDTM", "STTM", "UBLKCP", "SYNCS"} <= per[k], (k, per[k])
    # the fp32-equivalent forward reuses the A operand of two consecutive split products through the A collector (shared-memory A in layer 1,
    # TMEM A in layers 2 / 3: one TMEM read less per K slice next to the epilogue's tcgen05.ld traffic)
    for k in [k for k in per if "fused_mlp_tc3_kernel" in k]:
        assert {"A_KEEP", "A_REUSE"} <= per[k], (k, per[k])
    # the two-term backward variants move their rows with cp.async (the staging rings of this round)
    assert any("wgrad_tc3_kernelILb0" in k and "LDGSTS" in v for k, v in per.items())
    assert any("dgrad_tc3_kernelILb0" in k and "LDGSTS" in v for k, v in per.items())
