"""Shared helpers for the parity tests (oracle side only; test infrastructure)."""
import io
import contextlib
import types

import numpy as np
import torch

from oracle import fvgn_oracle as O

RHO, MU, DT = 1.0, 1e-3, 0.01


def params(**over):
    p = dict(edge_input_size=5, edge_one_hot=9, cell_input_size=2, cell_one_hot=0, dual_edge=False, message_passing_num=15,
             cell_output_size=2, edge_output_size=5, drop_out=False, mp_times=2, multihead=1, nn_act="SiLU", loss_cont=0,
             loss_mom=10, loss_face_uv=1, loss_face_p=1, loss="direct_mean_loss", cell_target_input_size=2,
             face_flux_target_input_size=3, dataset_size=1000, batch_size=16, cumulative_length=600, grid_feature_size=1,
             dataset_type="h5", n_epochs=1, train_traj_length=2, Noise_injection_factor=0.02, rho=RHO, mu=MU, dt=DT)
    p.update(over)
    return types.SimpleNamespace(**p)


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def tables_from(g, prefix="tab."):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def torch_reference_state_dict(seed=0):
    """Rebuild the reference-layout weights without the reference: same nn.Module construction order and shapes as
    FVGN.__init__ (FVGN.py:19-81) under the same seed => same tensors (verified against fixture checksums)."""
    import torch.nn as nn

    def mlp(i, o=128, ln=True):
        m = nn.Sequential(nn.Linear(i, 128), nn.SiLU(), nn.Linear(128, 128), nn.SiLU(), nn.Linear(128, o))
        return nn.Sequential(m, nn.LayerNorm(o)) if ln else m

    torch.manual_seed(seed)
    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[prefix + k] = v.detach().clone()

    put("model.encoder.eb_encoder.", mlp(14))
    put("model.encoder.cb_encoder.", mlp(2))
    for l in range(15):
        cb = mlp(192)
        eb = mlp(384)
        put(f"model.processer_list.{l}.eb_module.net.", eb)
        put(f"model.processer_list.{l}.cb_module.net.", cb)
    put("model.decoder.edge_decode_module.", mlp(128, 5, ln=False))
    for name, size in (("_acc_output_normalizer", 2), ("_face_uvp_flux_output_normalizer", 3), ("_edge_normalizer", 14), ("_cell_normalizer", 2)):
        sd[name + ".acc_count"] = torch.tensor(1.0)
        sd[name + ".num_accumulations"] = torch.tensor(1.0)
        sd[name + ".acc_sum"] = torch.ones(size)
        sd[name + ".acc_sum_squared"] = torch.ones(size)
    put("_grid_feature_normalizer.", nn.BatchNorm1d(1))
    return sd


def check_weights(sd, g):
    keys = [str(k) for k in g["wkeys"]]
    assert sorted(sd.keys()) == keys
    s = np.array([float(sd[k].double().sum()) for k in keys])
    a = np.array([float(sd[k].double().abs().sum()) for k in keys])
    np.testing.assert_allclose(s, g["wsum"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(a, g["wabs"], rtol=1e-12, atol=1e-9)


def load_state(sd, g, prefix="state."):
    sd = dict(sd)
    for k, v in g.items():
        if k.startswith(prefix):
            sd[k[len(prefix):]] = torch.as_tensor(v)
    return sd
