"""Offline converters: raw trajectories -> the H5-layout mesh store Data_Pool reads.

  TFRecord (+ meta.json)  ->  <out>/<split>.h5 | <split>.fvgn.npz      Extract_mesh/parse_tfrecord_refactor.py:79-107, 3019-3084
  COMSOL .mphtxt + data   ->  the same                                  Extract_mesh/parse_comsol.py:21-332

What the reference does per trajectory in Python dict loops over cells and faces (`extract_mesh_state`, parse_tfrecord_refactor.py:419-971:
seconds per 3.6k-cell mesh) is here ONE call of the device connectivity builder (`ops.build_connectivity` -> fvgn_connectivity_faces /
fvgn_connectivity_tables: bit-exact tables, tests/test_gpu_parity.py) plus a few small torch reductions for the flow scalars
(`cal_mean_u_and_cd`, :1045-1094).  Readers are self-contained: the TFRecord framing and the tf.train.Example wire format are parsed
here (tensorflow is not a dependency), the COMSOL text formats with numpy.

    python -m fvgn_b200.convert tfrecord <dataset_dir> <out_dir> [--splits train valid test] [--name incompressible]
    python -m fvgn_b200.convert comsol <mesh.mphtxt> <data.txt> <out_dir> --split train
"""
from __future__ import annotations

import json
import os
import re
import struct

import numpy as np
import torch

from . import ops
from .dataset.mesh_store import MeshStore, open_store
from .utils.utilities import NodeType

# ------------------------------------------------------------------------------------------------ TFRecord framing + tf.train.Example
_CRC_TABLE = None


def _crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli), table driven — only the writer (tests / tools) and `verify=True` reads use it"""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = np.array(t, dtype=np.uint32)
    crc = 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = _crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def tfrecord_records(path, verify=False):
    """yields the payload of every record of a TFRecord file: uint64 length | uint32 masked crc(length) | data | uint32 masked crc(data)"""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError(f"{path}: truncated record header")
            (n,) = struct.unpack("<Q", head[:8])
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise IOError(f"{path}: truncated record")
            if verify and (struct.unpack("<I", head[8:])[0] != _masked_crc(head[:8]) or struct.unpack("<I", tail)[0] != _masked_crc(data)):
                raise IOError(f"{path}: CRC mismatch")
            yield data


def _varint(buf, i):
    v = s = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << s
        if not b & 0x80:
            return v, i
        s += 7


def _fields(buf):
    """(field number, wire type, value) of one protobuf message; length-delimited values as memoryview slices"""
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fn, wt = key >> 3, key & 7
        if wt == 2:
            ln, i = _varint(buf, i)
            yield fn, wt, buf[i:i + ln]
            i += ln
        elif wt == 0:
            v, i = _varint(buf, i)
            yield fn, wt, v
        elif wt == 5:
            yield fn, wt, buf[i:i + 4]
            i += 4
        elif wt == 1:
            yield fn, wt, buf[i:i + 8]
            i += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")


def parse_example_bytes(record: bytes) -> dict:
    """tf.train.Example -> {feature name: [bytes, ...]} for its BytesList features (Example.features = 1; Features.feature = 1 (map entry:
    key = 1, value = 2); Feature.bytes_list = 1; BytesList.value = 1)"""
    out = {}
    mv = memoryview(record)
    for fn, _, features in _fields(mv):
        if fn != 1:
            continue
        for fn2, _, entry in _fields(features):
            if fn2 != 1:
                continue
            name, vals = None, []
            for fn3, _, v in _fields(entry):
                if fn3 == 1:
                    name = bytes(v).decode("utf-8")
                elif fn3 == 2:
                    for fn4, _, lst in _fields(v):
                        if fn4 == 1:  # bytes_list
                            vals += [bytes(b) for fn5, _, b in _fields(lst) if fn5 == 1]
            if name is not None:
                out[name] = vals
    return out


_TF_DTYPES = {"float32": np.float32, "float64": np.float64, "int32": np.int32, "int64": np.int64, "uint8": np.uint8}


def parse_trajectory(record: bytes, meta: dict) -> dict:
    """parse_tfrecord_refactor.py:79-95 (`_parse`): decode_raw + reshape per meta['features']; 'static' fields are tiled to the trajectory
    length exactly as the reference does"""
    feats = parse_example_bytes(record)
    out = {}
    for key, field in meta["features"].items():
        a = np.frombuffer(b"".join(feats[key]), dtype=_TF_DTYPES[field["dtype"]]).reshape(field["shape"])
        if field["type"] == "static":
            a = np.tile(a, (meta["trajectory_length"], 1, 1))
        elif field["type"] != "dynamic":
            raise ValueError("invalid data format")  # ('dynamic_varlen' does not occur in the cylinder_flow / airfoil sets)
        out[key] = a
    return out


def load_dataset(path, split):
    """parse_tfrecord_refactor.py:98-107: iterator over the trajectories of <path>/<split>.tfrecord"""
    with open(os.path.join(path, "meta.json")) as fp:
        meta = json.load(fp)
    for rec in tfrecord_records(os.path.join(path, split + ".tfrecord")):
        yield parse_trajectory(rec, meta)


def write_tfrecord(path, trajectories, meta):
    """inverse of load_dataset (tests, synthetic data sets): one tf.train.Example per trajectory, every field one BytesList value"""
    def ld(fn, payload):
        out, n = bytearray([(fn << 3) | 2]), len(payload)
        while True:
            b = n & 0x7F
            n >>= 7
            out.append(b | (0x80 if n else 0))
            if not n:
                break
        return bytes(out) + payload

    with open(path, "wb") as f:
        for tr in trajectories:
            entries = b""
            for key, field in meta["features"].items():
                a = np.asarray(tr[key])
                if field["type"] == "static":
                    a = a[:1]
                raw = np.ascontiguousarray(a.astype(_TF_DTYPES[field["dtype"]])).tobytes()
                feature = ld(1, ld(1, raw))                       # Feature{bytes_list = BytesList{value = raw}}
                entries += ld(1, ld(1, key.encode()) + ld(2, feature))  # map entry
            data = ld(1, entries)                                 # Example{features = Features{...}}
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", _masked_crc(head)) + data + struct.pack("<I", _masked_crc(data)))


# ------------------------------------------------------------------------------------------------ extract_mesh_state on the device
def flow_scalars(tab, velocity0, pressure0, rho, mu):
    """cal_mean_u_and_cd (parse_tfrecord_refactor.py:1045-1094): mean inflow velocity, angle of attack, characteristic length (largest
    distance between obstacle-wall nodes), Reynolds number"""
    face = tab["face"].long()
    ft = tab["face_type"].view(-1)
    uvp = torch.cat((velocity0, pressure0), dim=1)
    on_edge = (uvp[face[0]] + uvp[face[1]]) / 2.0
    inflow = ft == int(NodeType.INFLOW)
    pos = tab["mesh_pos"]
    top, bottom = pos[:, 1].max(), pos[:, 1].min()
    mean_u = ((on_edge[inflow, 0] * tab["face_length"].view(-1)[inflow]).sum() / (top - bottom)).to(torch.float32)
    aoa = torch.rad2deg(torch.atan2(on_edge[inflow, 1], on_edge[inflow, 0])).mean()
    wall = pos[tab["node_type"].view(-1) == int(NodeType.WALL_BOUNDARY)]
    cyl = wall[(wall[:, 1] > bottom) & (wall[:, 1] < top)]
    L0 = torch.cdist(cyl, cyl, p=2).max() if cyl.shape[0] > 1 else torch.zeros((), device=pos.device)
    mean_u_np, L0_np = mean_u.cpu().numpy(), L0.cpu().numpy()
    return {"aoa": aoa.cpu().numpy(), "mean_u": mean_u_np, "charac_scale": L0_np, "rho": np.array(rho), "mu": np.array(mu),
            "reynolds_num": (mean_u_np * L0_np) / np.array(mu)}


def extract_mesh_state(dataset, h5_writer, index, solving_params=None, face_rule=None, device="cuda", renumber=False):
    """parse_tfrecord_refactor.py:419-971 on the device.  dataset: mesh_pos [T|1,N,2], cells [T|1,C,3], node_type [T|1,N,1], velocity
    [T,N,2], pressure [T,N,1] (numpy).  Writes group str(index) with the reference's keys / shapes / dtypes into `h5_writer` (h5py.File
    or MeshStore) and returns it as a dict.  face_rule None follows the reference's branch selection — `"compressible" in name` (:491;
    the shipped name "incompressible" CONTAINS it, Appendix B) -> rule "A", else "B".
    renumber="cells" (or True) / "all": cells (and nodes) are first sorted along a Z-order curve (renumber.py: gathers then hit L2 whatever
    numbering the source had; "cells" leaves every result unchanged up to that permutation, "all" also reorders the faces but changes the
    orientation conventions the reference derives from node numbers); the group additionally holds `renumber|node_perm` /
    `renumber|cell_perm` (new -> old), everything else is an ordinary group."""
    sp = solving_params or {"name": "incompressible", "rho": 1.0, "mu": 1e-3, "dt": 0.01}
    perms = None
    if renumber:
        from .renumber import renumber_mesh

        m1, node_perm, cell_perm = renumber_mesh({"mesh_pos": np.asarray(dataset["mesh_pos"][0]), "cells": np.asarray(dataset["cells"][0]),
                                                  "node_type": np.asarray(dataset["node_type"][0]).reshape(-1),
                                                  "velocity": np.asarray(dataset["velocity"]), "pressure": np.asarray(dataset["pressure"])},
                                                 nodes=(renumber == "all"))
        dataset = dict(dataset, mesh_pos=m1["mesh_pos"][None], cells=m1["cells"][None], node_type=m1["node_type"].reshape(1, -1, 1),
                       velocity=m1["velocity"], pressure=m1["pressure"])
        perms = {"renumber|node_perm": node_perm, "renumber|cell_perm": cell_perm}
    if face_rule is None:
        face_rule = "A" if "compressible" in sp.get("name", "") else "B"
    dev = torch.device(device)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev)
    with torch.cuda.device(dev):
        tab = ops.build_connectivity(t(dataset["cells"][0], torch.int32), t(dataset["mesh_pos"][0], torch.float32),
                                     t(np.asarray(dataset["node_type"][0]).reshape(-1), torch.int32), face_rule)
        vel, prs = np.asarray(dataset["velocity"], dtype=np.float32), np.asarray(dataset["pressure"], dtype=np.float32)
        scal = flow_scalars(tab, t(vel[0], torch.float32), t(prs[0], torch.float32), sp.get("rho", 1.0), sp.get("mu", 1e-3))
        mesh = {k: tab[k].cpu().numpy()[None] for k in ("cells_node", "centroid", "face", "face_length", "face_type", "cells_face", "cells_type",
                                                        "unit_norm_v", "neighbour_cell", "cell_factor", "cells_area")}
    mesh["mesh_pos"] = np.asarray(dataset["mesh_pos"][0], dtype=np.float32)[None]
    mesh["node_type"] = np.asarray(dataset["node_type"][0]).reshape(1, -1, 1).astype(np.int32)
    mesh["target|velocity_on_node"], mesh["target|pressure_on_node"] = vel, prs
    mesh.update(scal)
    if perms is not None:
        mesh.update(perms)
    if h5_writer is not None:
        grp = h5_writer.create_group(str(index))
        for k, v in mesh.items():
            grp.create_dataset(k, data=v)
    return mesh


def convert_tfrecord(tf_dataset_path, out_dir, splits=("valid", "train", "test"), solving_params=None, limit=None, device="cuda", renumber=False):
    """the __main__ of parse_tfrecord_refactor.py:3019-3084 (saving_h5 branch)"""
    os.makedirs(out_dir, exist_ok=True)
    counts = {}
    for split in splits:
        if not os.path.exists(os.path.join(tf_dataset_path, split + ".tfrecord")):
            continue
        w = _writer(out_dir, split)
        n = 0
        for n, traj in enumerate(load_dataset(tf_dataset_path, split), 1):
            extract_mesh_state(traj, w, n - 1, solving_params, device=device, renumber=renumber)
            if limit is not None and n >= limit:
                break
        w.close()
        counts[split] = n
    return counts


def _writer(out_dir, split):
    try:
        import types

        import h5py

        if not isinstance(h5py, types.ModuleType):  # (a test double registered in sys.modules is not h5py)
            raise ImportError("h5py is not installed")
        return h5py.File(os.path.join(out_dir, f"{split}.h5"), "w")
    except ImportError:
        return MeshStore(os.path.join(out_dir, f"{split}.fvgn.npz"), "w")


# ------------------------------------------------------------------------------------------------ COMSOL
_NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"


class ComsolCase:
    """parse_comsol.py:21-332 (`Cosmol_manager`): a .mphtxt mesh (vertices, `edg` boundary elements, `tri` elements) + a COMSOL data
    export (x, y rows and u, v, p rows per time step) -> the raw trajectory dict extract_mesh_state takes.  Node typing
    (:213-262): x == x_min strictly inside the walls -> INFLOW; y at / beyond the top or bottom wall -> WALL_BOUNDARY; x == x_max
    inside -> OUTFLOW; other boundary-element nodes inside the channel -> WALL_BOUNDARY (the obstacle); else NORMAL.  The data rows are
    re-ordered to the mesh's vertex order by matching float32 coordinates (:199-263)."""

    def __init__(self, mesh_file, data_file):
        self.case_info = {"U_mean": 0.0, "R": 0.3 if "rect" in mesh_file else 0.5 if "NACA0012" in mesh_file else 0.0, "aoa": 0.0}
        self._read_mesh(mesh_file)
        self._read_data(data_file)

    def _read_mesh(self, path):
        lines = open(path).read().split("\n")
        i, n = 0, len(lines)
        self.mesh_pos = self.boundary = self.cells = None
        nv = 0
        while i < n:
            row = lines[i]
            if row.endswith("# number of mesh vertices"):
                nv = int(row.split(" ")[0])
            elif row == "# Mesh vertex coordinates":
                self.mesh_pos = np.array([[float(v) for v in lines[i + 1 + k].split()[:2]] for k in range(nv)], dtype=np.float64)
                i += nv
            elif row in ("3 edg # type name", "3 tri # type name"):
                kind = row[2:5]
                ne = 0
                while i < n and lines[i] != "# Elements":
                    if lines[i].endswith("# number of elements"):
                        ne = int(lines[i].split(" ")[0])
                    i += 1
                elems = np.array([[int(v) for v in lines[i + 1 + k].split()] for k in range(ne)], dtype=np.int64)
                if kind == "edg":
                    self.boundary = elems[:, :2]
                else:
                    self.cells = elems[:, :3]
                i += ne
            i += 1
        if self.mesh_pos is None or self.cells is None or self.boundary is None:
            raise ValueError(f"{path}: vertices / edg / tri sections not found")

    def _read_data(self, path):
        X, Y, U, V, P = [], [], [], [], []
        lines = open(path).read().split("\n")
        row_of = lambda s: np.array([float(v) for v in re.findall(_NUM, s)], dtype=np.float64)
        for i, row in enumerate(lines[:-1]):
            if "x (m) @ t" in row or row.startswith("% x"):
                X.append(row_of(lines[i + 1]))
            elif "y (m) @ t" in row or row.startswith("% y"):
                Y.append(row_of(lines[i + 1]))
            elif "u (m/s) @" in row:
                for para in row.split(","):
                    if "U_mean" in para:
                        self.case_info["U_mean"] = float(para.split("=")[1])
                    elif "aoa" in para:
                        self.case_info["aoa"] = float(para.split("=")[1])
                    elif "R" in para and "=" in para:
                        self.case_info["R"] = float(para.split("=")[1])
                U.append(row_of(lines[i + 1]))
            elif "v (m/s) @" in row:
                V.append(row_of(lines[i + 1]))
            elif "p (Pa) @" in row:
                P.append(row_of(lines[i + 1]))
        if not X:
            raise ValueError(f"data {path} read failed")
        self.data_pos = np.stack((X[0], Y[0]), axis=-1)
        self.velocity = np.stack((np.array(U), np.array(V)), axis=-1)
        self.pressure = np.array(P)[..., None]

    def trajectory(self, max_steps=600):
        pos = self.mesh_pos.astype(np.float32)
        top, bottom, outlet, inlet = pos[:, 1].max(), pos[:, 1].min(), pos[:, 0].max(), pos[:, 0].min()
        inside_y = (pos[:, 1] > bottom) & (pos[:, 1] < top)
        on_boundary = np.zeros(pos.shape[0], dtype=bool)
        on_boundary[self.boundary.reshape(-1)] = True
        nt = np.full(pos.shape[0], int(NodeType.NORMAL), dtype=np.int32)
        obstacle = on_boundary & (pos[:, 0] > 0) & (pos[:, 0] < outlet) & (pos[:, 1] > 0) & (pos[:, 1] < top)
        nt[obstacle] = int(NodeType.WALL_BOUNDARY)                     # lowest priority of the elif chain (:245-255) ...
        nt[(pos[:, 0] == outlet) & inside_y] = int(NodeType.OUTFLOW)   # ... :236-244
        nt[~inside_y] = int(NodeType.WALL_BOUNDARY)                    # ... :231-235 (y >= top or y <= bottom)
        nt[(pos[:, 0] == inlet) & inside_y] = int(NodeType.INFLOW)     # ... :223-230 (highest)
        # data rows -> mesh vertex order by exact float32 coordinate match
        key = lambda a: a.astype(np.float32).view(np.uint32).astype(np.uint64)
        dk = (key(self.data_pos[:, 0]) << np.uint64(32)) | key(self.data_pos[:, 1])
        mk = (key(pos[:, 0]) << np.uint64(32)) | key(pos[:, 1])
        order = np.argsort(dk, kind="stable")
        where = np.searchsorted(dk[order], mk)
        if np.any(where >= dk.shape[0]) or np.any(dk[order][np.minimum(where, dk.shape[0] - 1)] != mk):
            raise ValueError("data coordinates do not match the mesh vertices")
        idx = order[where]
        T = min(max_steps, self.velocity.shape[0])
        return {"mesh_pos": pos[None], "boundary": self.boundary.astype(np.int32)[None], "cells": self.cells.astype(np.int32)[None],
                "node_type": nt.reshape(1, -1, 1), "velocity": self.velocity[:T, idx].astype(np.float32),
                "pressure": self.pressure[:T, idx].astype(np.float32), "U_mean": self.case_info["U_mean"], "R": self.case_info["R"],
                "aoa": self.case_info["aoa"]}


def convert_comsol(mesh_file, data_file, h5_writer, index, solving_params=None, device="cuda", renumber=False):
    """Cosmol_manager.extract_mesh (parse_comsol.py:188-332) -> extract_mesh_state.  The reference passes its `path` dict as
    solving_params (:323-330), which has no "name" with "compressible" in it: face-typing rule B."""
    sp = dict(solving_params or {"name": "comsol", "rho": 1.0, "mu": 1e-3, "dt": 0.01})
    return extract_mesh_state(ComsolCase(mesh_file, data_file).trajectory(), h5_writer, index, sp, device=device, renumber=renumber)


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(prog="fvgn_b200.convert")
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("tfrecord")
    a.add_argument("dataset_dir"), a.add_argument("out_dir")
    a.add_argument("--splits", nargs="+", default=["valid", "train", "test"])
    a.add_argument("--name", default="incompressible")
    a.add_argument("--limit", type=int, default=None)
    a.add_argument("--renumber", choices=["cells", "all"], default=None, help="sort cells (and nodes) along a Z-order curve first (renumber.py)")
    b = sub.add_parser("comsol")
    b.add_argument("mesh_file"), b.add_argument("data_file"), b.add_argument("out_dir")
    b.add_argument("--split", default="train")
    b.add_argument("--index", type=int, default=0)
    b.add_argument("--renumber", choices=["cells", "all"], default=None)
    args = ap.parse_args(argv)
    if args.cmd == "tfrecord":
        print(convert_tfrecord(args.dataset_dir, args.out_dir, args.splits, {"name": args.name, "rho": 1.0, "mu": 1e-3, "dt": 0.01}, args.limit,
                               renumber=args.renumber))
    else:
        os.makedirs(args.out_dir, exist_ok=True)
        w = _writer(args.out_dir, args.split)
        convert_comsol(args.mesh_file, args.data_file, w, args.index, renumber=args.renumber)
        w.close()


if __name__ == "__main__":
    main()
