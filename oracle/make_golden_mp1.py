"""Test infrastructure (build container only): a small fixture of the UNMODIFIED reference run with mp_times=1 (GN_blocks.py:130-138) on the
mesh, weights (torch.manual_seed(0) constructor init) and primed normaliser / BatchNorm state of tests/golden/fwd_tiny.npz:

    tests/golden/fwd_tiny_mp1.npz    roll_uv [3,C,2], roll_uvp [3,F,3]   (3 autoregressive eval steps; rollout.py:368-425)

so that the oracle's mp_times=1 branch is pinned on machines without /root/reference too (tests/test_oracle_golden.py).
usage: python -m oracle.make_golden_mp1"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle import ref_loader as RL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
RHO, MU, DT = 1.0, 1e-3, 0.01


def main():
    ref = RL.load()
    g = dict(np.load(os.path.join(OUT, "fwd_tiny.npz"), allow_pickle=False))
    params = RL.default_params(train_traj_length=2, mp_times=1)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.FVGN(device="cpu", model_dir=None, params=params)
    sd = model.state_dict()
    for k, v in g.items():
        if k.startswith("state."):
            sd[k[6:]] = torch.as_tensor(v)
    model.load_state_dict(sd)
    model.eval()
    R = RL.ref_extract_mesh_state(g["mesh_pos"], g["cells"], g["node_type"], g["velocity"], g["pressure"])
    dpv = RL.ref_data_pool([R], params, is_training=False)
    gn, ge, gc, mi, mb = dpv.require_one_mesh(start_epoch=0, rollout_index=0)
    assert np.array_equal(gc.x.numpy(), g["in_x"]) and np.array_equal(gc.edge_attr.numpy(), g["in_edge_attr"])
    cell_type, Re, rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
    roll_uv, roll_uvp = [], []
    with torch.no_grad():
        for _ in range(3):
            uvp, nuv = model(graph=gc, graph_edge=ge, graph_node=gn, rho=RHO, mu=MU, dt=DT, edge_one_hot=9, cell_one_hot=0, device="cpu")
            roll_uv.append(nuv.numpy().copy()), roll_uvp.append(uvp.numpy().copy())
            gc = ref.Data_Pool.create_next_graph(gc, ge, mb, nuv, cell_type, Re, rmp)
    np.savez_compressed(os.path.join(OUT, "fwd_tiny_mp1.npz"), roll_uv=np.stack(roll_uv), roll_uvp=np.stack(roll_uvp))
    print("fwd_tiny_mp1", np.stack(roll_uv).shape, np.stack(roll_uvp).shape)


if __name__ == "__main__":
    main()
