# worst relative-L2 gradient error vs the reference fixture for a given FVGN_BWD_PRODUCTS
import sys, numpy as np, torch
sys.path.insert(0, ".")
from tests import helpers as H
from tests.test_gpu_parity import _model, _graphs, dev
import fvgn_b200 as fv
from tests.conftest import GOLDEN
import os
g = dict(np.load(os.path.join(GOLDEN, "train_tiny.npz"))); f = dict(np.load(os.path.join(GOLDEN, "fwd_tiny.npz")))
sys.path.insert(0, "tests")
from tests.test_gpu_training import _oracle_grads
want_loss, want = _oracle_grads(f, g, "global_mean_sum")
model = _model(fv, f).train().set_math_mode("bf16x3")
gl = fv.dataset.collate([_graphs(fv, f, "tab.", "velocity", "pressure", pair=1), _graphs(fv, f, "tab1.", "velocity1", "pressure1", pair=0)])
gn, ge, gc, mi, mb = fv.dataset.train_datapreprocessing(gl, cell_noise=dev(g["noise"]), flip_keep=dev(g["flip"]))
p = H.params(loss="global_mean_sum")
out = model(graph=gc, graph_edge=ge, graph_node=gn, rho=H.RHO, mu=H.MU, dt=H.DT, edge_one_hot=9, cell_one_hot=0)
loss = fv.compute_FVM_loss(params=p, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                           target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
loss[0].backward()
named = dict(model.named_parameters())
errs = sorted(((H.rel_l2(named[k].grad.cpu(), w), k) for k, w in want.items()), reverse=True)
print("BWD_PRODUCTS", fv.ops.BWD_PRODUCTS, "worst", errs[:3], "median", errs[len(errs)//2][0])
