"""Autoregressive rollout with zero host work per step: the loop body of the reference's rollout.py:368-425
— FVGN.forward (eval) followed by Data_Pool.create_next_graph (dataset/Load_mesh.py:1076-1115) — is captured ONCE into a CUDA graph
(every kernel of libfvgn_b200 only enqueues work on the current stream, nothing synchronises, all buffers are persistent) and
replayed per step.  Results are bit-identical to the eager loop; what disappears is the Python / launch overhead (~60 launches per
step), which dominates on cylinder_flow-sized meshes."""
import torch

from . import ops
from ._lib import MATH_MODES
from .autograd import network_mlps
from .dataset.Load_mesh import RolloutState


class GraphedRollout:
    def __init__(self, model, graph_node, graph_edge, graph_cell, rho, mu, dt, edge_one_hot=9, cell_one_hot=0, warmup=2):
        """graph_* are the bags returned by dataset.valid_datapreprocessing (graph_cell.x = [type, Re, u, v] at the start step).
        The initial state is snapshotted; `reset()` returns to it."""
        if model.training:
            raise RuntimeError("GraphedRollout captures the eval forward: call model.eval() first")
        self.model, self.gn, self.ge, self.gc = model, graph_node, graph_edge, graph_cell
        self.args = dict(rho=rho, mu=mu, dt=dt, edge_one_hot=edge_one_hot, cell_one_hot=cell_one_hot)
        self.state = RolloutState(graph_cell, graph_edge)
        self._x_init, self._ea_init = self.state.x0.clone(), self.state.ea0.clone()
        self.gc.x, self.gc.edge_attr = self.state.x0, self.state.ea0
        self.graph = None
        self.uvp = self.next_uv = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):  # builds the topology / packs weights / sizes the allocator pools outside the capture
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        self.reset()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.uvp, self.next_uv = self._eager_step()
        self._packed_epoch = self._weights_state()
        self.reset()

    def _eager_step(self):
        uvp, nuv = self.model(graph=self.gc, graph_edge=self.ge, graph_node=self.gn, **self.args)
        self.state.advance_(self.gc, nuv)
        return uvp, nuv

    def reset(self):
        self.state.x0.copy_(self._x_init)
        self.state.ea0.copy_(self._ea_init)
        self.gc.x, self.gc.edge_attr = self.state.x0, self.state.ea0

    def _refresh_weights(self):
        """The captured step reads the packed tensor-core weight tiles (persistent buffers, packed outside the capture).  If the fp32
        weights have changed since (an optimizer step, a GraphedTrainStep replay, load_state_dict), re-pack them in place before replaying."""
        mode = MATH_MODES[self.model.math_mode]
        if mode == 0:
            return
        key = self._weights_state()
        if key == self._packed_epoch:
            return
        for m in network_mlps(self.model):
            m.mlp_ref().ensure_packed(mode)
        self._packed_epoch = key

    def _weights_state(self):
        """(optimizer-step epoch, sum of the weight tensors' version counters): ~3 us per call"""
        ts = self.__dict__.get("_wts")
        if ts is None:
            ts = self._wts = [t for m in network_mlps(self.model) for t in m.mlp_ref().tensors[:6:2]]
        v = 0
        for t in ts:
            v += t._version
        return (ops._WEIGHTS_EPOCH[0], v)

    def step(self):
        """one rollout step; returns (predicted_uvp_on_edge[F,3], next_UV_on_cell[C,2]) — persistent buffers, overwritten by the next step"""
        self._refresh_weights()
        self.graph.replay()
        return self.uvp, self.next_uv

    def check(self):
        """synchronises; raises FloatingPointError if a fp32-equivalent ('bf16x3') tile produced a non-finite row since the last check
        (fp16 operand overflow, |activation| >= ~8190): the rollout from that step on is NaN — re-run it with set_math_mode('fp32')"""
        ops.raise_if_nonfinite("rollout step")

    def run(self, n_steps, out_uv=None, out_uvp=None, check=False):
        """n_steps steps; optionally records every step into out_uv[n,C,2] / out_uvp[n,F,3] (device tensors); check=True: one device
        synchronisation at the end + the non-finite status check"""
        self._refresh_weights()
        for k in range(n_steps):
            self.graph.replay()
            if out_uv is not None:
                out_uv[k].copy_(self.next_uv)
            if out_uvp is not None:
                out_uvp[k].copy_(self.uvp)
        if check:
            self.check()
        return self.uvp, self.next_uv
