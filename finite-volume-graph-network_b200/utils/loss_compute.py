"""compute_FVM_loss with the reference's keyword interface and 5-tuple return (utils/loss_compute.py:24-115), computed by
two deterministic reduction kernels (fvgn_loss_fwd).  `loss_continuity` is 0.0 exactly as in the reference (App. D.2)."""
import torch

from .. import ops


def compute_FVM_loss(params=None, graph_cell=None, graph_edge=None, mask_face_interior=None, predicted_edge_attr=None,
                     target_face_uvp_normalized=None, predicted_delta_u=None, target_delta_u=None, loss_function=None):
    if params.loss not in ops.LOSS_MODES:
        raise ValueError(f"unknown loss '{params.loss}'")
    mode = ops.LOSS_MODES[params.loss]
    cb = fb = None
    n_graphs = 1
    if mode != 0:
        cb = ops.as_i32(graph_cell.batch)
        fb = ops.as_i32(graph_edge.batch)
        n_graphs = int(getattr(graph_cell, "num_graphs", 0)) or int(graph_cell.batch[-1].item()) + 1
    from ..autograd import LossFunction

    out = LossFunction.apply(predicted_delta_u, target_delta_u, predicted_edge_attr, target_face_uvp_normalized, mask_face_interior, cb, fb,
                             n_graphs, mode, params.loss_mom, params.loss_face_uv, params.loss_face_p)
    return out[0], 0.0, out[2].detach(), out[3].detach(), out[4].detach()
