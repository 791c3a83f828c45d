"""CPU: the dataset layer — host logic only, no kernels.

  * fvgn_b200.dataset.Data_Pool.get / _build_graph over H5-layout groups == the UNMODIFIED reference's Data_Pool on the same groups
    (every attribute of the three bags, torch.equal) — live where /root/reference or oracle/_ref exists
  * MeshStore (the H5 key layout in a numpy archive) round trip; Data_Pool over the reopened store
  * PinnedBatcher collation == PyG-style Batch.from_data_list; staging buffers are reused across batches of a size bucket
  * Batch.to_data_list inverts from_data_list
  * the DataLoader over Data_Pool yields what the reference's PyG DataLoader yields
  * converters: TFRecord framing / tf.train.Example parser round trip; the COMSOL reader against the reference's Cosmol_manager
"""
import io
import contextlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from oracle import ref_loader as RL
from tests import helpers as H

import fvgn_b200 as fv

needs_ref = pytest.mark.skipif(not RL.available(), reason="reference (or its oracle/_ref copy) not present")


def _groups(n=3, T=4, kind="grid"):
    out, meshes = [], []
    for s in range(n):
        m = fv.synthetic.make(kind, seed=20 + s, nx=9 + s, ny=7, T=T)
        tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
        tab["mesh_pos"], tab["node_type"] = m["mesh_pos"], m["node_type"]
        out.append(RL.h5_group_from_tables(tab, m["velocity"], m["pressure"]))
        meshes.append(m)
    return out, meshes


def _same_bag(a, b, what):
    ka = sorted(k for k in a.keys if k not in ("ptr", "_slices"))
    kb = sorted(k for k in b.keys if k not in ("ptr", "_slices"))
    assert ka == kb, (what, ka, kb)
    for k in ka:
        va, vb = getattr(a, k), getattr(b, k)
        if torch.is_tensor(va):
            assert va.dtype == vb.dtype and va.shape == vb.shape and torch.equal(va, vb), (what, k, va.dtype, vb.dtype, va.shape, vb.shape)
        else:
            assert va == vb, (what, k)


@needs_ref
@pytest.mark.parametrize("training", [True, False])
def test_get_and_build_graph_match_the_reference(training):
    groups, _ = _groups()
    p = H.params(train_traj_length=3, dataset_size=len(groups))
    ref = RL.ref_data_pool(groups, RL.default_params(train_traj_length=3, dataset_size=len(groups)), is_training=training)
    ref._set_status(is_training=training)
    mine = fv.dataset.Data_Pool(params=p, is_traning=training, device="cpu")
    mine.pool = {str(i): g for i, g in enumerate(groups)}
    assert mine.len() == ref.len() == 9
    for idx in (0, 2, 4, 8):
        for a, b, nm in zip(mine.get(idx), ref.get(idx), ("node", "edge", "cell")):
            _same_bag(a, b, f"get({idx}).{nm}")
    for i in range(len(groups)):
        mine._build_graph(None, i)
        ref._build_graph(ref.pool[str(i)], i)
    for attr in ("mbatch_graph_node", "mbatch_graph_edge", "mbatch_graph_cell"):
        for a, b in zip(getattr(mine, attr), getattr(ref, attr)):
            _same_bag(a, b, attr)
    # the block-diagonal collation of the rollout batch (PinnedBatcher) == the reference's Batch.from_data_list
    refL = RL.load()
    for name, attr in (("node", "mbatch_graph_node"), ("edge", "mbatch_graph_edge"), ("cell", "mbatch_graph_cell")):
        got = mine._batcher.collate(name, getattr(mine, attr))
        want = refL.Batch.from_data_list(getattr(ref, attr))
        for k in want.keys:
            if k in ("_slices",):
                continue
            va, vb = getattr(got, k), getattr(want, k)
            if torch.is_tensor(vb):
                assert torch.equal(va, vb), (name, k)
    mine.reset()
    assert not mine.mbatch_graph_node and not mine.mbatch_graph_cell


def test_mesh_store_round_trip_and_data_pool_over_it(tmp_path):
    groups, _ = _groups()
    w = fv.dataset.MeshStore(str(tmp_path / "train.fvgn.npz"), "w")
    for i, g in enumerate(groups):
        grp = w.create_group(str(i))
        for k, v in g.items():
            grp.create_dataset(k, data=v)
    w.close()
    p = H.params(train_traj_length=3, dataset_size=len(groups), dataset_dir_h5=str(tmp_path))
    dp = fv.dataset.Data_Pool(params=p, is_traning=True, device="cpu")
    assert dp.load_mesh_to_cpu(mode="h5", split="train") == str(tmp_path)
    assert len(dp.pool) == len(groups) and dp.len() == 9
    for i, g in enumerate(groups):
        r = dp.pool[str(i)]
        assert sorted(r.keys()) == sorted(g.keys())
        for k, v in g.items():
            got = np.asarray(r[k])
            assert got.dtype == np.asarray(v).dtype and np.array_equal(got, np.asarray(v)), (i, k)
        assert float(r["reynolds_num"][()]) == float(g["reynolds_num"]) and r["cells_type"][0].shape == g["cells_type"][0].shape
    mem = fv.dataset.Data_Pool(params=p, is_traning=True, device="cpu")
    mem.pool = {str(i): g for i, g in enumerate(groups)}
    for idx in (1, 5):
        for a, b in zip(dp.get(idx), mem.get(idx)):
            _same_bag(a, b, "store vs dict")
    with pytest.raises(NotImplementedError):
        dp.load_mesh_to_cpu(mode="tf", split="train")


def test_pinned_batcher_matches_from_data_list_and_reuses_its_buffers():
    groups, _ = _groups(n=4)
    p = H.params(train_traj_length=3, dataset_size=len(groups))
    dp = fv.dataset.Data_Pool(params=p, is_traning=True, device="cpu")
    dp.pool = {str(i): g for i, g in enumerate(groups)}
    pb = fv.dataset.PinnedBatcher(pin=False)
    for sel in ([0, 1, 2], [3, 1, 0], [2, 3, 1]):
        samples = [dp.get(3 * i) for i in sel]
        for k, name in enumerate(("node", "edge", "cell")):
            got = pb.collate(name, [s[k] for s in samples])
            want = fv.Batch.from_data_list([s[k] for s in samples])
            for key in want.keys:
                va, vb = getattr(got, key), getattr(want, key)
                if torch.is_tensor(vb):
                    assert va.dtype == vb.dtype and torch.equal(va, vb), (name, key)
            assert got.num_graphs == 3
    first = pb.allocations
    samples = [dp.get(3 * i + 1) for i in (1, 2, 3)]
    for k, name in enumerate(("node", "edge", "cell")):
        pb.collate(name, [s[k] for s in samples])
    assert pb.allocations == first  # same size bucket: no new staging memory
    assert fv.dataset.PinnedBatcher.bucket(5000) == 6144 and fv.dataset.PinnedBatcher.bucket(4096) == 4096 and fv.dataset.PinnedBatcher.bucket(9000) == 12288


def test_batch_to_data_list_inverts_from_data_list():
    groups, _ = _groups()
    p = H.params(train_traj_length=3, dataset_size=len(groups))
    dp = fv.dataset.Data_Pool(params=p, is_traning=False, device="cpu")
    dp.pool = {str(i): g for i, g in enumerate(groups)}
    samples = [dp.get(3 * i) for i in range(3)]
    for k in (0, 2):  # graph_node and graph_cell: their index tables point into their own rows
        b = fv.Batch.from_data_list([s[k] for s in samples])
        back = b.to_data_list()
        assert len(back) == 3
        for a, s in zip(back, samples):
            for key in s[k].keys:
                va, vb = getattr(a, key), getattr(s[k], key)
                if torch.is_tensor(vb) and vb.dim() > 0 and key not in ("face",):
                    assert torch.equal(va, vb), key


@needs_ref
def test_data_loader_yields_the_reference_collation():
    groups, _ = _groups()
    p = H.params(train_traj_length=3, dataset_size=len(groups))
    dp = fv.dataset.Data_Pool(params=p, is_traning=True, device="cpu")
    dp.pool = {str(i): g for i, g in enumerate(groups)}
    loader = fv.dataset.make_data_loader(dp, batch_size=4, shuffle=False)
    batches = list(loader)
    assert len(batches) == 3 and all(len(b) == 3 for b in batches)
    ref = RL.ref_data_pool(groups, RL.default_params(train_traj_length=3, dataset_size=len(groups)), is_training=True)
    ref._set_status(is_training=True)
    refL = RL.load()
    want = [refL.Batch.from_data_list([ref.get(i)[k] for i in range(4)]) for k in range(3)]
    for got, w in zip(batches[0], want):
        for key in w.keys:
            if key == "_slices":
                continue
            vb = getattr(w, key)
            if torch.is_tensor(vb):
                assert torch.equal(getattr(got, key), vb), key


# ------------------------------------------------------------------------------------------------ converters (host side)
def _tf_meta(T):
    return {"trajectory_length": T, "field_names": ["cells", "mesh_pos", "node_type", "velocity", "pressure"],
            "features": {"cells": {"type": "static", "shape": [1, -1, 3], "dtype": "int32"},
                         "mesh_pos": {"type": "static", "shape": [1, -1, 2], "dtype": "float32"},
                         "node_type": {"type": "static", "shape": [1, -1, 1], "dtype": "int32"},
                         "velocity": {"type": "dynamic", "shape": [T, -1, 2], "dtype": "float32"},
                         "pressure": {"type": "dynamic", "shape": [T, -1, 1], "dtype": "float32"}}}


def write_synthetic_tfrecord(dirname, meshes, split="train"):
    from fvgn_b200 import convert as CV

    T = meshes[0]["velocity"].shape[0]
    meta = _tf_meta(T)
    json.dump(meta, open(os.path.join(dirname, "meta.json"), "w"))
    trs = [{"cells": m["cells"][None], "mesh_pos": m["mesh_pos"][None], "node_type": m["node_type"].reshape(1, -1, 1), "velocity": m["velocity"],
            "pressure": m["pressure"]} for m in meshes]
    CV.write_tfrecord(os.path.join(dirname, split + ".tfrecord"), trs, meta)
    return meta


def test_tfrecord_reader_round_trip(tmp_path):
    from fvgn_b200 import convert as CV

    meshes = [fv.synthetic.make("grid", seed=s, nx=9, ny=7 + s, T=4) for s in range(3)]
    write_synthetic_tfrecord(str(tmp_path), meshes)
    out = list(CV.load_dataset(str(tmp_path), "train"))
    assert len(out) == 3
    for m, o in zip(meshes, out):
        assert o["cells"].shape == (4,) + m["cells"].shape and o["cells"].dtype == np.int32  # 'static' fields tiled like the reference's _parse
        assert np.array_equal(o["cells"][0], m["cells"]) and np.array_equal(o["cells"][3], m["cells"])
        assert np.array_equal(o["mesh_pos"][0], m["mesh_pos"]) and np.array_equal(o["node_type"][0, :, 0], m["node_type"].reshape(-1))
        assert np.array_equal(o["velocity"], m["velocity"]) and np.array_equal(o["pressure"], m["pressure"])
    assert len(list(CV.tfrecord_records(str(tmp_path / "train.tfrecord"), verify=True))) == 3  # CRC-32C framing is self-consistent
    assert CV._crc32c(b"123456789") == 0xE3069283  # the CRC-32C check value
    raw = open(tmp_path / "train.tfrecord", "rb").read()
    open(tmp_path / "bad.tfrecord", "wb").write(raw[:-3])
    with pytest.raises(IOError):
        list(CV.tfrecord_records(str(tmp_path / "bad.tfrecord")))


def write_synthetic_comsol(dirname, m, name="cyl"):
    """a .mphtxt mesh + a COMSOL data export of a synthetic mesh, data rows in a SHUFFLED vertex order (as COMSOL exports them)"""
    pos, cells = m["mesh_pos"].astype(np.float64), m["cells"]
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]])
    key = np.sort(e, axis=1)
    uniq, cnt = np.unique(key, axis=0, return_counts=True)
    boundary = uniq[cnt == 1]
    mesh_file = os.path.join(dirname, f"{name}_mesh1.mphtxt")
    with open(mesh_file, "w") as f:
        f.write("# Created by COMSOL Multiphysics.\n\n2 # sdim\n%d # number of mesh vertices\n0 # lowest mesh vertex index\n\n# Mesh vertex coordinates\n" % pos.shape[0])
        for p in pos:
            f.write("%.17g %.17g \n" % (p[0], p[1]))
        f.write("\n3 # number of element types\n\n# Type #0\n\n3 vtx # type name\n\n\n1 # number of vertices per element\n0 # number of elements\n# Elements\n\n")
        f.write("# Type #1\n\n3 edg # type name\n\n\n2 # number of vertices per element\n%d # number of elements\n# Elements\n" % boundary.shape[0])
        for b in boundary:
            f.write("%d %d \n" % (b[0], b[1]))
        f.write("\n%d # number of geometric entity indices\n# Geometric entity indices\n" % boundary.shape[0])
        f.write("".join("0 \n" for _ in boundary))
        f.write("\n# Type #2\n\n3 tri # type name\n\n\n3 # number of vertices per element\n%d # number of elements\n# Elements\n" % cells.shape[0])
        for c in cells:
            f.write("%d %d %d \n" % (c[0], c[1], c[2]))
        f.write("\n%d # number of geometric entity indices\n# Geometric entity indices\n" % cells.shape[0])
        f.write("".join("1 \n" for _ in cells))
    perm = np.random.default_rng(3).permutation(pos.shape[0])
    data_file = os.path.join(dirname, f"{name}_data1.txt")
    row = lambda a: "   ".join("%.17g" % v for v in a) + "\n"
    with open(data_file, "w") as f:
        f.write("% Model: synthetic\n% x\n" + row(pos[perm, 0].astype(np.float32)))
        f.write("% y\n" + row(pos[perm, 1].astype(np.float32)))
        for t in range(m["velocity"].shape[0]):
            f.write("%% u (m/s) @ t=%g, U_mean=1.25, aoa=0.5\n" % (0.01 * t) + row(m["velocity"][t, perm, 0]))
            f.write("%% v (m/s) @ t=%g, U_mean=1.25, aoa=0.5\n" % (0.01 * t) + row(m["velocity"][t, perm, 1]))
            f.write("%% p (Pa) @ t=%g, U_mean=1.25, aoa=0.5\n" % (0.01 * t) + row(m["pressure"][t, perm, 0]))
    return mesh_file, data_file


def test_comsol_reader(tmp_path):
    from fvgn_b200 import convert as CV

    m = fv.synthetic.make("cyl", seed=4, T=3, n_target=400) if "n_target" in fv.synthetic.make.__code__.co_varnames else fv.synthetic.make("cyl", seed=4, T=3)
    mesh_file, data_file = write_synthetic_comsol(str(tmp_path), m)
    tr = CV.ComsolCase(mesh_file, data_file).trajectory()
    assert np.array_equal(tr["mesh_pos"][0], m["mesh_pos"].astype(np.float32)) and np.array_equal(tr["cells"][0], m["cells"])
    assert np.array_equal(tr["velocity"], m["velocity"]) and np.array_equal(tr["pressure"], m["pressure"])  # rows re-ordered to the mesh's vertex order
    assert tr["U_mean"] == 1.25 and tr["aoa"] == 0.5
    nt = tr["node_type"].reshape(-1)
    pos = tr["mesh_pos"][0]
    assert set(np.unique(nt)) <= {0, 4, 5, 6}
    assert np.all(nt[(pos[:, 0] == pos[:, 0].min()) & (pos[:, 1] > pos[:, 1].min()) & (pos[:, 1] < pos[:, 1].max())] == 4)
    assert np.all(nt[(pos[:, 1] == pos[:, 1].max())] == 6) and np.all(nt[pos[:, 1] == pos[:, 1].min()] == 6)


@needs_ref
def test_comsol_reader_matches_the_reference_manager(tmp_path):
    """the reference's Cosmol_manager (parse_comsol.py:21-332) on the same files: its raw `mesh` dict == ComsolCase.trajectory()"""
    import importlib

    from fvgn_b200 import convert as CV

    RL.load()
    PC = importlib.import_module("Extract_mesh.parse_comsol")
    m = fv.synthetic.make("grid", seed=6, nx=11, ny=8, T=3)
    mesh_file, data_file = write_synthetic_comsol(str(tmp_path), m, name="rect")
    path = {"saving_h5": False, "rho": 1.0, "mu": 1e-3, "dt": 0.01, "name": "comsol"}
    with contextlib.redirect_stdout(io.StringIO()):
        mgr = PC.Cosmol_manager(mesh_file, data_file, h5_writer=None, path=path)
        # (extract_mesh ends in extract_mesh_state; with saving_h5 False it only computes)
        cuda_saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            ref_mesh = mgr.extract_mesh(plot=False, data_index=0)
        finally:
            torch.Tensor.cuda = cuda_saved
    tr = CV.ComsolCase(mesh_file, data_file).trajectory()
    assert np.array_equal(ref_mesh["mesh_pos"][0], tr["mesh_pos"][0]) and np.array_equal(ref_mesh["cells"][0], tr["cells"][0])
    assert np.array_equal(ref_mesh["node_type"][0], tr["node_type"][0])
    assert np.array_equal(ref_mesh["boundary"][0], tr["boundary"][0])
    T = tr["velocity"].shape[0]
    assert np.array_equal(ref_mesh["velocity"][:T], tr["velocity"]) and np.array_equal(ref_mesh["pressure"][:T], tr["pressure"])
    assert ref_mesh["U_mean"] == tr["U_mean"] and ref_mesh["aoa"] == tr["aoa"] and ref_mesh["R"] == tr["R"]
