"""CellBlock / EdgeBlock with the reference's constructor and forward signatures (FVGN_model/GN_blocks.py:10-145),
executing as ONE fused sm_100a kernel each (plus the deterministic node segment-sum for the CellBlock):

  CellBlock:  node_agg = segment_sum(edge halves -> nodes)            fvgn_node_aggregate_fwd   (GN_blocks.py:91,111-122)
              x' = LN(MLP([x | mean3(node_agg[cells_node])]))          fvgn_cell_block_fwd       (:125-143)
  EdgeBlock:  e' = LN(MLP([x'[s] | x'[r] | e]))                        fvgn_edge_block_fwd       (:43-52)

The concatenated [C,192] / [F,384] operands the reference materialises exist only as shared-memory tiles."""
import torch.nn as nn

from .. import ops
from ..data import Data
from ..topology import topology_of


class EdgeBlock(nn.Module):
    def __init__(self, custom_func=None, dual_edge=False):
        super().__init__()
        if dual_edge:
            raise NotImplementedError("dual_edge=True is outside the B200 hot path (SURVEY.md §8f rank 4)")
        self.net = custom_func
        self.dual_edge = dual_edge

    def forward(self, graph, graph_node=None, topo=None):
        if topo is None:
            topo_nc = ops.as_i32(graph.edge_index)
        else:
            topo_nc = topo.nc
        e_prime, _ = ops.edge_block_fwd(self.net.mlp_ref(), graph.x.contiguous(), graph.edge_attr.contiguous(), topo_nc,
                                        self.net.math_mode_id(), want_prime=True, want_out=False)
        return Data(x=graph.x, edge_attr=e_prime, edge_index=graph.edge_index)


class CellBlock(nn.Module):
    def __init__(self, input_size, attention_size, mp_times=2, MultiHead=1, custom_func=None, dual_edge=False):
        super().__init__()
        if attention_size % MultiHead > 0:
            raise ValueError("MultiHead must be the factor of attention_size")
        if dual_edge:
            # (the reference itself cannot run this: GnBlock sizes the cell MLP for 128+64 inputs, dual_edge feeds it 128+128)
            raise NotImplementedError("dual_edge=True is outside the B200 hot path (SURVEY.md §8f rank 4)")
        self.net = custom_func
        self.mp_times = mp_times
        self.dual_edge = dual_edge

    def forward(self, graph, graph_node, topo=None):
        if topo is None:
            topo = topology_of(graph, graph_node)
        a = topo.agg(self.mp_times)  # mp_times >= 2: faces -> nodes -> cells; 1: faces -> cells (GN_blocks.py:110-138)
        node_agg = ops.node_aggregate_fwd(graph.edge_attr.contiguous(), a.ptr, a.items, a.rows)
        x_prime, _ = ops.cell_block_fwd(self.net.mlp_ref(), graph.x.contiguous(), node_agg, a.idx, self.net.math_mode_id(), want_out=False)
        return Data(x=x_prime, edge_attr=graph.edge_attr, edge_index=graph.edge_index)
