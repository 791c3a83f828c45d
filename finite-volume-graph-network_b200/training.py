"""One whole training step — train_datapreprocessing (a2) + FVGN.forward (training) + compute_FVM_loss + backward + (gradient
all-reduce) + optimizer step, i.e. the loop body of the reference's train.py:150-225 — captured ONCE into a CUDA graph and replayed.
With the forward and the whole backward on the tensor cores the eager step is bound by ~1200 Python-driven launches (~30 ms at batch
16); the graph replays the same kernels back to back.  Results are those of the eager step (same kernels, same order)."""
import torch

from . import ops
from .dataset import Load_mesh as dataset
from .utils.loss_compute import compute_FVM_loss


def _shallow(bag):
    out = bag.__class__()
    out.__dict__.update(bag.__dict__)
    return out


class GraphedTrainStep:
    def __init__(self, model, optimizer, params, graph_node, graph_edge, graph_cell, rho, mu, dt, edge_one_hot=9, cell_one_hot=0,
                 all_reduce=None, warmup=3, preprocess_kwargs=None):
        """graph_* = one collated batch in the layout of Data_Pool.get (graph_node.x [N,6]); the index tables are static, the node
        fields are refreshed per step with `load(node_x)`.  The optimizer must be capturable (e.g. AdamW(..., capturable=True, fused=True):
        the fused implementation is a handful of launches, the foreach one ~590).
        `all_reduce`: optional callable run between backward and optimizer.step (parallel.GradBucket.all_reduce_mean_)."""
        self.model, self.opt, self.params = model, optimizer, params
        self.gn, self.ge, self.gc = graph_node, graph_edge, graph_cell
        self.args = dict(rho=rho, mu=mu, dt=dt, edge_one_hot=edge_one_hot, cell_one_hot=cell_one_hot)
        self.all_reduce = all_reduce
        self.pre_kw = dict(preprocess_kwargs or {})  # e.g. fixed cell_noise / flip_keep draws (tests); default: fresh device draws
        self.node_x = graph_node.x  # static input buffer
        self.loss = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(2, warmup)):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        l0 = ops.LAUNCHES
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        self.launches_per_step = ops.LAUNCHES - l0  # libfvgn_b200 kernels inside one replay

    def _eager_step(self):
        gn, ge, gc = _shallow(self.gn), _shallow(self.ge), _shallow(self.gc)
        gn.x = self.node_x
        gn, ge, gc, mi, mb = dataset.train_datapreprocessing([gn, ge, gc], **self.pre_kw)
        self.opt.zero_grad(set_to_none=True)
        out = self.model(graph=gc, graph_edge=ge, graph_node=gn, **self.args)
        loss = compute_FVM_loss(params=self.params, graph_cell=gc, graph_edge=ge, mask_face_interior=mi, predicted_edge_attr=out[3],
                                target_face_uvp_normalized=out[4], predicted_delta_u=out[1], target_delta_u=out[2])
        loss[0].backward()
        if self.all_reduce is not None:
            self.all_reduce()
        self.opt.step()
        return torch.stack([loss[0].detach().reshape(()), loss[2].reshape(()), loss[3].reshape(()), loss[4].reshape(())])

    def load(self, node_x, non_blocking=True):
        """next batch's node fields [N,6] (host pinned or device) into the static input buffer"""
        self.node_x.copy_(node_x, non_blocking=non_blocking)

    def step(self):
        """replays the captured step; returns the persistent [loss, loss_mom, loss_face_uv, loss_face_p] tensor of this step"""
        self.graph.replay()
        # the replay has stepped the optimizer without running any of the Python-side signals the packed-weight caches key on (the
        # optimizer-step hook, Tensor._version): advance the epoch so that the next EAGER forward (validation between replays, a
        # GraphedRollout captured earlier, an eval after set_math_mode) re-packs its tensor-core weight tiles instead of reading stale ones
        ops._WEIGHTS_EPOCH[0] += 1
        return self.loss

    def check(self):
        """synchronises; raises FloatingPointError if a fp32-equivalent forward tile produced a non-finite row since the last check"""
        ops.raise_if_nonfinite("training step")
