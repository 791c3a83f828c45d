"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on synthetic meshes.

Run in the build container only:   python oracle/make_golden.py
The fixtures are what pins the oracle (oracle/fvgn_oracle.py) where /root/reference is absent (GPU box).

Fixtures
  conn_<mesh>_<rule>.npz   inputs (mesh_pos, cells, node_type) + every table of extract_mesh_state
  fwd_tiny.npz             primed normaliser state, model inputs, per-stage intermediates, eval outputs, 3-step rollout
  fwd_cyl.npz              eval outputs after 1 step and a 10-step rollout on a cylinder mesh (BASELINE config 1)
  train_tiny.npz           2-graph training batch: pre-processed inputs, 5-tuple, both losses, gradients
Weights are NOT stored (8.8 MB): they are `torch.manual_seed(0)` + the reference constructor; the fixtures keep
per-tensor checksums so a consumer can prove it rebuilt the same weights.
"""
import os
import sys
import io
import copy
import contextlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "finite-volume-graph-network_b200"))
import synthetic as S  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
RHO, MU, DT = 1.0, 1e-3, 0.01


def squeeze_tables(R):
    return {k: (v[0] if v.ndim > 0 and v.shape[0] == 1 and k not in ("target|velocity_on_node", "target|pressure_on_node") else v)
            for k, v in R.items()}


def weight_checksums(sd):
    keys = sorted(sd.keys())
    sums = np.array([float(sd[k].double().sum()) for k in keys])
    asums = np.array([float(sd[k].double().abs().sum()) for k in keys])
    return np.array(keys), sums, asums


def new_model(ref, params, seed=0):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.FVGN(device="cpu", model_dir=None, params=params)


def train_batch(ref, dp, params, picks, seed):
    lists = [dp.get(case * params.train_traj_length + t) for case, t in picks]
    gl = [ref.Batch.from_data_list([l[k] for l in lists]) for k in range(3)]
    torch.manual_seed(seed)
    C = gl[2].x.shape[0]
    F = gl[1].x.shape[0]
    # replay of the reference's own draws (utils/noise.py:17-23 then Load_mesh.py:442-446) for the consumer
    nu = torch.normal(std=params.Noise_injection_factor, mean=0.0, size=(C, 1))
    nv = torch.normal(std=params.Noise_injection_factor, mean=0.0, size=(C, 1))
    flip = (torch.randint(2, (1, F)) == torch.randint(2, (1, F))).view(-1)
    torch.manual_seed(seed)
    gn, ge, gc, mi, mb = dp.train_datapreprocessing(gl)
    return gn, ge, gc, mi, mb, torch.cat((nu, nv), 1), flip


def prime(ref, model, dp, params, picks_list, seed0=100):
    """Two training-mode forwards so the four normalisers and the BatchNorm hold non-trivial statistics."""
    model.train()
    for i, picks in enumerate(picks_list):
        gn, ge, gc, mi, mb, _, _ = train_batch(ref, dp, params, picks, seed0 + i)
        with torch.no_grad():
            model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0, device="cpu")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = RL.load()

    # ---------------- connectivity ----------------
    for name, kind, seed in [("tiny", "tiny", 3), ("cyl", "cyl", 3), ("foil", "foil", 5)]:
        m = S.make(kind, seed=seed)
        for rule in "AB":
            if name == "foil" and rule == "A":
                continue
            R = squeeze_tables(RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"], face_rule=rule))
            keep = {k: R[k] for k in ["cells_node", "face", "cells_face", "neighbour_cell", "face_type", "cells_type", "face_length",
                                      "unit_norm_v", "centroid", "cells_area", "cell_factor"]}
            np.savez_compressed(os.path.join(OUT, f"conn_{name}_{rule}.npz"), in_mesh_pos=m["mesh_pos"], in_cells=m["cells"],
                                in_node_type=m["node_type"], **keep)
            print("conn", name, rule, {k: v.shape for k, v in keep.items() if k in ("face", "cells_face")})

    params = RL.default_params(train_traj_length=2)

    # ---------------- tiny: forward with intermediates ----------------
    meshes = [S.make("tiny", seed=s, T=5) for s in (1, 2)]
    Rs = [RL.ref_extract_mesh_state(m["mesh_pos"], m["cells"], m["node_type"], m["velocity"], m["pressure"]) for m in meshes]
    dp = RL.ref_data_pool(Rs, params, is_training=True)
    dp._set_status(True)
    model = new_model(ref, params, 0)
    wk, ws, wa = weight_checksums(model.state_dict())
    prime(ref, model, dp, params, [[(0, 0), (1, 1)], [(1, 0), (0, 1)]])
    primed = {k: v.clone() for k, v in model.state_dict().items()}
    state_small = {"state." + k: v.numpy() for k, v in primed.items() if "normalizer" in k}

    dpv = RL.ref_data_pool(Rs, params, is_training=False)
    gn, ge, gc, mi, mb = dpv.require_one_mesh(start_epoch=0, rollout_index=0)
    inputs = dict(in_x=gc.x.numpy().copy(), in_edge_attr=gc.edge_attr.numpy().copy(), in_edge_x=ge.x.numpy().copy(),
                  in_edge_y=ge.y.numpy().copy(), in_cell_y=gc.y.numpy().copy(), mask_face_boundary=mb.numpy())
    cap = {}
    hooks = []
    epd = model.model
    hooks.append(epd.encoder.register_forward_hook(lambda m_, i, o: cap.update(enc_x=o.x.detach().numpy().copy(), enc_e=o.edge_attr.detach().numpy().copy())))
    for l in (0, 1, 14):
        hooks.append(epd.processer_list[l].register_forward_hook(
            lambda m_, i, o, l=l: cap.update({f"x_l{l}": o.x.detach().numpy().copy(), f"e_l{l}": o.edge_attr.detach().numpy().copy()})))
    hooks.append(epd.processer_list[0].cb_module.register_forward_hook(lambda m_, i, o: cap.update(cb0_x=o.x.detach().numpy().copy())))
    hooks.append(epd.processer_list[0].eb_module.register_forward_hook(lambda m_, i, o: cap.update(eb0_e=o.edge_attr.detach().numpy().copy())))
    hooks.append(epd.decoder.register_forward_hook(lambda m_, i, o: cap.update(decoded=o[1].detach().numpy().copy())))
    hooks.append(model._grid_feature_normalizer.register_forward_hook(lambda m_, i, o: cap.update(face_area=o.detach().numpy().copy())))
    hooks.append(model.intergator.register_forward_hook(lambda m_, i, o: cap.update(delta_u=o[1].detach().numpy().copy())))
    model.eval()
    cell_type = gc.x[:, 0:1].clone()
    Re = gc.x[:, 1:2].clone()
    rmp = gc.edge_attr[:, 2:5].clone()
    roll_uv, roll_uvp = [], []
    with torch.no_grad():
        for step in range(3):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0, device="cpu")
            if step == 0:
                for h in hooks:
                    h.remove()
                cap["cell_in"] = gc.x.numpy().copy()
                cap["edge_in"] = gc.edge_attr.numpy().copy()
            roll_uv.append(nuv.numpy().copy())
            roll_uvp.append(uvp.numpy().copy())
            gc = ref.Data_Pool.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)
    T0 = squeeze_tables(Rs[0])
    np.savez_compressed(
        os.path.join(OUT, "fwd_tiny.npz"), wkeys=wk, wsum=ws, wabs=wa,
        mesh_pos=meshes[0]["mesh_pos"], cells=meshes[0]["cells"], node_type=meshes[0]["node_type"],
        velocity=meshes[0]["velocity"], pressure=meshes[0]["pressure"],
        mesh_pos1=meshes[1]["mesh_pos"], cells1=meshes[1]["cells"], node_type1=meshes[1]["node_type"],
        velocity1=meshes[1]["velocity"], pressure1=meshes[1]["pressure"],
        **{"tab." + k: v for k, v in T0.items() if isinstance(v, np.ndarray) and v.ndim > 0},
        **{"tab1." + k: v for k, v in squeeze_tables(Rs[1]).items() if isinstance(v, np.ndarray) and v.ndim > 0},
        **state_small, **inputs, **{"cap." + k: v for k, v in cap.items()},
        roll_uv=np.stack(roll_uv), roll_uvp=np.stack(roll_uvp))
    print("fwd_tiny", {k: v.shape for k, v in cap.items()})

    # ---------------- tiny: training batch ----------------
    out = {}
    for loss_name in ("direct_mean_loss", "global_mean_sum"):
        params.loss = loss_name
        model2 = new_model(ref, params, 0)
        model2.load_state_dict(primed)
        model2.train()
        gn, ge, gc, mi, mb, noise, flip = train_batch(ref, dp, params, [(0, 1), (1, 0)], 123)
        if loss_name == "direct_mean_loss":
            out.update(noise=noise.numpy(), flip=flip.numpy(), pre_x=gc.x.numpy().copy(), pre_edge_attr=gc.edge_attr.numpy().copy(),
                       pre_cell_y=gc.y.numpy().copy(), pre_edge_y=ge.y.numpy().copy(), mask_face_interior=mi.numpy(),
                       cell_batch=gc.batch.numpy(), face_batch=ge.batch.numpy())
        res = model2(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0, device="cpu")
        loss = ref.compute_FVM_loss(params=params, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=res[3],
                                    target_face_uvp_normalized=res[4], predicted_delta_u=res[1], target_delta_u=res[2],
                                    loss_function=torch.nn.MSELoss(reduction="mean"))
        loss[0].backward()
        tag = "dm" if loss_name == "direct_mean_loss" else "gm"
        out[f"{tag}.loss"] = np.array([float(loss[0]), float(torch.as_tensor(loss[2]).mean()), float(torch.as_tensor(loss[3]).mean()), float(torch.as_tensor(loss[4]).mean())])
        named = dict(model2.named_parameters())
        gk = sorted(named.keys())
        out[f"{tag}.gkeys"] = np.array(gk)
        out[f"{tag}.gsum"] = np.array([float(named[k].grad.double().sum()) for k in gk])
        out[f"{tag}.gnorm"] = np.array([float(named[k].grad.double().norm()) for k in gk])
        for k in ["model.decoder.edge_decode_module.4.weight", "model.decoder.edge_decode_module.4.bias",
                  "model.encoder.cb_encoder.0.0.weight", "model.encoder.eb_encoder.0.0.weight",
                  "model.processer_list.7.cb_module.net.0.0.weight", "model.processer_list.14.eb_module.net.1.weight",
                  "_grid_feature_normalizer.weight", "_grid_feature_normalizer.bias"]:
            out[f"{tag}.grad.{k}"] = named[k].grad.numpy().copy()
        if tag == "dm":
            out.update(pred_delta_u=res[1].detach().numpy(), target_delta_u=res[2].detach().numpy(),
                       pred_edge=res[3].detach().numpy(), target_face=res[4].detach().numpy())
            sd_after = model2.state_dict()
            out.update({"after." + k: v.numpy().copy() for k, v in sd_after.items() if "normalizer" in k})
    np.savez_compressed(os.path.join(OUT, "train_tiny.npz"), **out)
    print("train_tiny", out["dm.loss"], out["gm.loss"])

    # ---------------- cylinder: 1 step + 10-step rollout (BASELINE config 1) ----------------
    params = RL.default_params(train_traj_length=2)
    mc = S.make("cyl", seed=11, T=12)
    Rc = RL.ref_extract_mesh_state(mc["mesh_pos"], mc["cells"], mc["node_type"], mc["velocity"], mc["pressure"])
    dpt = RL.ref_data_pool([Rc], params, is_training=True)
    dpt._set_status(True)
    model = new_model(ref, params, 0)
    prime(ref, model, dpt, params, [[(0, 0)], [(0, 1)]])
    state_small = {"state." + k: v.numpy().copy() for k, v in model.state_dict().items() if "normalizer" in k}
    dpv = RL.ref_data_pool([Rc], params, is_training=False)
    gn, ge, gc, mi, mb = dpv.require_one_mesh(start_epoch=0, rollout_index=0)
    model.eval()
    cell_type = gc.x[:, 0:1].clone()
    Re = gc.x[:, 1:2].clone()
    rmp = gc.edge_attr[:, 2:5].clone()
    uv_steps, uvp_steps = [], []
    with torch.no_grad():
        for step in range(10):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0, device="cpu")
            uv_steps.append(nuv.numpy().copy())
            uvp_steps.append(uvp.numpy().copy())
            gc = ref.Data_Pool.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)
    Tc = squeeze_tables(Rc)
    np.savez_compressed(os.path.join(OUT, "fwd_cyl.npz"), mesh_pos=mc["mesh_pos"], cells=mc["cells"], node_type=mc["node_type"],
                        velocity=mc["velocity"], pressure=mc["pressure"],
                        **{"tab." + k: v for k, v in Tc.items() if isinstance(v, np.ndarray) and v.ndim > 0},
                        **state_small, uv_step1=uv_steps[0], uvp_step1=uvp_steps[0], uv_step10=uv_steps[-1], uvp_step10=uvp_steps[-1],
                        uv_norms=np.array([np.linalg.norm(u) for u in uv_steps]))
    print("fwd_cyl", uv_steps[0].shape, uvp_steps[0].shape, [float(np.linalg.norm(u)) for u in uv_steps])
    sz = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden total bytes", sz)


if __name__ == "__main__":
    main()
