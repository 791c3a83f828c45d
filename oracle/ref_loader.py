"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference (FVGN-src/main) on CPU.

The reference is pure Python but needs torch_geometric / torch_scatter / h5py / matplotlib /
tensorflow / tfrecord / circle_fit, none of which are in this image.  `oracle/shims/` holds
attribute-bag stand-ins for the two PyG/torch_scatter pieces that carry arithmetic
(`Data`, `Batch.from_data_list`, `global_mean_pool`, `scatter_add`); everything else is a
`MagicMock`.  Nothing is copied from the reference: its modules are imported from where they lie.

Available where /root/reference exists (the build container) or where `oracle/make_ref.py` has placed its
byte-for-byte copy under `oracle/_ref/` (git-ignored; it travels to the GPU box with the snapshot, where bench.py's
reference arm / torch-eager CUDA leg and the drop-in tests use it).  The committed fixtures in tests/golden/ that
`oracle/make_golden.py` produced with this loader cover the boxes that have neither.
"""
import os
import sys
import types
import importlib
import importlib.machinery
from unittest import mock

def _find_root():
    """$FVGN_REFERENCE_ROOT, else /root/reference (the build container), else oracle/_ref (the byte-for-byte copy oracle/make_ref.py
    placed there; git-ignored, travels to the GPU box)"""
    cands = [os.environ.get("FVGN_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "FVGN-src", "main", "FVGN_model", "FVGN.py")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()
REF_MAIN = os.path.join(REF_ROOT, "FVGN-src", "main")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_loaded = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_MAIN, "FVGN_model", "FVGN.py"))


def default_params(**over):
    """The hot-path-relevant fields of utils/get_param.py:50-144 with their defaults
    (run_train.sh:17 overrides loss to direct_mean_loss)."""
    p = dict(
        edge_input_size=5, edge_one_hot=9, cell_input_size=2, cell_one_hot=0, dual_edge=False,
        message_passing_num=15, cell_output_size=2, edge_output_size=5, drop_out=False, mp_times=2,
        multihead=1, nn_act="SiLU", loss_cont=0, loss_mom=10, loss_face_uv=1, loss_face_p=1,
        loss="direct_mean_loss", cell_target_input_size=2, face_flux_target_input_size=3,
        dataset_size=1000, batch_size=16, cumulative_length=600, grid_feature_size=1,
        dataset_type="h5", n_epochs=1, train_traj_length=400, Noise_injection_factor=0.02,
        rho=1.0, mu=1e-3, dt=0.01,
    )
    p.update(over)
    return types.SimpleNamespace(**p)


def load():
    """Import the reference modules; returns a namespace with FVGN, EncoderProcesserDecoder, ...
    Safe to call repeatedly."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference not present at {REF_MAIN}")
    import torch

    for name in ["tensorflow", "h5py", "matplotlib", "matplotlib.pyplot", "matplotlib.tri",
                 "matplotlib.animation", "matplotlib.font_manager", "matplotlib.ticker",
                 "matplotlib.cm", "matplotlib.colors", "circle_fit", "pandas_stub_unused",
                 "natsort", "statsmodels"]:
        if name not in sys.modules:
            m = sys.modules[name] = mock.MagicMock()
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)  # importlib.util.find_spec(name) must not raise (torch._dynamo probes it)
    for p in (_SHIMS, REF_MAIN, os.path.join(REF_MAIN, "Extract_mesh")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # CPU-only box: `.cuda()` is the identity (Load_mesh.py:788-790,1117-1125; loss_compute.py:40-75)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    mods = {}
    mods["utilities"] = importlib.import_module("utils.utilities")
    mods["normalization"] = importlib.import_module("utils.normalization")
    mods["loss_compute"] = importlib.import_module("utils.loss_compute")
    mods["GN_blocks"] = importlib.import_module("FVGN_model.GN_blocks")
    mods["EPD"] = importlib.import_module("FVGN_model.EncoderProcesserDecoder")
    mods["FVGN_mod"] = importlib.import_module("FVGN_model.FVGN")
    mods["Load_mesh"] = importlib.import_module("dataset.Load_mesh")
    mods["parse"] = importlib.import_module("Extract_mesh.parse_tfrecord_refactor")
    from torch_geometric.data import Data
    from torch_geometric.data.batch import Batch

    _loaded.update(mods)
    _loaded.update(
        FVGN=mods["FVGN_mod"].FVGN,
        EncoderProcesserDecoder=mods["EPD"].EncoderProcesserDecoder,
        Intergrator=mods["EPD"].Intergrator,
        GnBlock=mods["EPD"].GnBlock,
        build_mlp=mods["EPD"].build_mlp,
        CellBlock=mods["GN_blocks"].CellBlock,
        EdgeBlock=mods["GN_blocks"].EdgeBlock,
        Normalizer=mods["normalization"].Normalizer,
        compute_FVM_loss=mods["loss_compute"].compute_FVM_loss,
        Data_Pool=mods["Load_mesh"].Data_Pool,
        extract_mesh_state=mods["parse"].extract_mesh_state,
        NodeType=mods["utilities"].NodeType,
        Data=Data,
        Batch=Batch,
    )
    return types.SimpleNamespace(**_loaded)


class _DictWriter:
    """dict-backed stand-in for the h5 writer used at parse_tfrecord_refactor.py:962-965."""

    def __init__(self):
        self.groups = {}

    def create_group(self, name):
        g = _DictGroup()
        self.groups[name] = g
        return g


class _DictGroup(dict):
    def create_dataset(self, key, data=None):
        import numpy as np

        self[key] = np.asarray(data)


def ref_extract_mesh_state(mesh_pos, cells, node_type, velocity, pressure, rho=1.0, mu=1e-3, face_rule="B"):
    """Run the reference's extract_mesh_state (parse_tfrecord_refactor.py:419-971) on raw arrays.
    Returns its H5 group as a dict of numpy arrays (leading [1,...] / [T,...] axes kept).
    face_rule "A": solving_params name "incompressible" as shipped (:3032) -- which, containing the substring
    "compressible", takes the compressible typing branch (:491); "B": a name without that substring, which
    takes the else branch (:511-620)."""
    import numpy as np

    ref = load()
    T = velocity.shape[0]
    dataset = {
        "mesh_pos": np.broadcast_to(mesh_pos[None].astype(np.float32), (1,) + mesh_pos.shape).copy(),
        "cells": cells[None].astype(np.int32),
        "node_type": node_type.reshape(1, -1, 1).astype(np.int32),
        "velocity": velocity.astype(np.float32),
        "pressure": pressure.astype(np.float32),
    }
    w = _DictWriter()
    solving = {"name": "incompressible" if face_rule == "A" else "incomp-flow", "rho": rho, "mu": mu, "dt": 0.01}
    path = {"saving_h5": True, "mesh_file_path": "synthetic"}
    import io
    import contextlib

    with contextlib.redirect_stdout(io.StringIO()):
        ref.extract_mesh_state(dataset, None, 0, origin_writer=None, solving_params=solving,
                               h5_writer=w, path=path)
    return dict(w.groups["0"])


def h5_group_from_tables(tab, velocity, pressure, rho=1.0, mu=1e-3):
    """The H5 group extract_mesh_state writes (parse_tfrecord_refactor.py:890-899, 962-965, 1045-1094) assembled from
    already-built tables (oracle.fvgn_oracle.build_connectivity: bit-identical to the reference's, tests/test_oracle_*).
    Static arrays get the leading [1, ...] axis, the fields stay [T, N, .]; scalars as 0-dim arrays."""
    import numpy as np

    g = {k: np.asarray(tab[k])[None] for k in ("cells_node", "centroid", "face", "face_length", "face_type", "cells_face", "cells_type",
                                                "unit_norm_v", "neighbour_cell", "cell_factor", "cells_area")}
    g["mesh_pos"] = np.asarray(tab["mesh_pos"], dtype=np.float32)[None]
    g["node_type"] = np.asarray(tab["node_type"]).reshape(1, -1, 1).astype(np.int32)
    g["target|velocity_on_node"] = np.asarray(velocity, dtype=np.float32)
    g["target|pressure_on_node"] = np.asarray(pressure, dtype=np.float32)
    g["mean_u"] = np.float32(1.0)
    g["charac_scale"] = np.float32(0.1)
    g["aoa"] = np.float32(0.0)
    g["rho"], g["mu"] = np.array(rho), np.array(mu)
    g["reynolds_num"] = np.array(float(g["mean_u"]) * float(g["charac_scale"]) / mu)
    return g


def ref_data_pool(mesh_dicts, params=None, is_training=False, device="cpu"):
    """A reference Data_Pool (Load_mesh.py) whose `pool` is a dict of extract_mesh_state outputs."""
    ref = load()
    params = params or default_params()
    import io
    import contextlib

    with contextlib.redirect_stdout(io.StringIO()):
        dp = ref.Data_Pool(params, is_traning=is_training, device=device)
    dp.pool = {str(i): m for i, m in enumerate(mesh_dicts)}
    return dp
