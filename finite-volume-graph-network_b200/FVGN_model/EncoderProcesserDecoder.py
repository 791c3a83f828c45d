"""Encoder / GnBlock / Decoder / EncoderProcesserDecoder / Intergrator with the reference's module tree, signatures and
state_dict keys (FVGN_model/EncoderProcesserDecoder.py), dispatching to libfvgn_b200's fused kernels.

Parameter containers are ordinary nn.Linear / nn.LayerNorm modules created in the reference's construction order, so
(a) `state_dict()` has the reference's 283 keys/shapes and reference checkpoints load unchanged, (b) the same
`torch.manual_seed` yields the same initial weights."""
import torch
import torch.nn as nn

from .. import ops
from .._lib import MATH_MODES
from ..data import Data
from ..topology import topology_of
from .GN_blocks import EdgeBlock, CellBlock


class FusedMLP(nn.Sequential):
    """What build_mlp returns: same children as the reference's nn.Sequential, forward = one fused kernel."""

    def __init__(self, *children, lay_norm=True):
        super().__init__(*children)
        self._lay_norm = lay_norm
        self._ref = None
        self.math_mode = "fp32"

    def _apply(self, fn, *a, **k):
        self._ref = None  # parameters may be re-allocated (.to / .cuda / .float)
        return super()._apply(fn, *a, **k)

    def _linears(self):
        lin = self.__dict__.get("_lin")
        if lin is None:  # (indexing an nn.Sequential costs ~1.5 us per item; this runs for every launch)
            seq = self[0] if self._lay_norm else self
            lin = self.__dict__["_lin"] = (seq[0], seq[2], seq[4])
        return lin

    def mlp_ref(self) -> ops.MlpRef:
        l1, l2, l3 = self._linears()
        if self._ref is None or self._ref.tensors[0].data_ptr() != l1.weight.data_ptr():
            ln = self[1] if self._lay_norm else None
            self._ref = ops.MlpRef(l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias,
                                   ln.weight if ln is not None else None, ln.bias if ln is not None else None)
        # split-product count of the bf16x3 kernels: all six (fp32-equivalent) unless the module is in plain "bf16" mode, where the
        # training path runs the same kernels with the leading product only (bf16 tiles, fp32 accumulation)
        self._ref.struct.n_products = 1 if self.math_mode == "bf16" else 0
        return self._ref

    def math_mode_id(self):
        return MATH_MODES[self.math_mode]

    def forward(self, x):
        return ops.mlp_rows_fwd(self.mlp_ref(), x, self.math_mode_id())


def build_mlp(in_size, hidden_size, out_size, drop_out=True, lay_norm=True, dropout_prob=0.2, nn_act="SiLU"):
    """Reference signature (EncoderProcesserDecoder.py:12-48).  The fused kernels implement SiLU, hidden 128, no dropout —
    which is every MLP the reference's model actually builds (it passes drop_out=False everywhere, App. D.6)."""
    if nn_act == "SiLU":
        act = nn.SiLU
    elif nn_act == "ReLU":
        raise NotImplementedError("nn_act='ReLU' is not built into the fused kernels (the reference default is SiLU)")
    else:
        raise NotImplementedError
    if drop_out:
        raise NotImplementedError("build_mlp(drop_out=True) is never used by the reference model; fused kernels have no dropout")
    if hidden_size != 128:
        raise ValueError("hidden_size must be 128")
    module = [nn.Linear(in_size, hidden_size), act(), nn.Linear(hidden_size, hidden_size), act(), nn.Linear(hidden_size, out_size)]
    if lay_norm:
        return FusedMLP(nn.Sequential(*module), nn.LayerNorm(normalized_shape=out_size), lay_norm=True)
    return FusedMLP(*module, lay_norm=False)


class Encoder(nn.Module):
    def __init__(self, edge_input_size=128, cell_input_size=128, hidden_size=128, nn_act="SiLU"):
        super().__init__()
        self.eb_encoder = build_mlp(edge_input_size, hidden_size, hidden_size, drop_out=False, nn_act=nn_act)
        self.cb_encoder = build_mlp(cell_input_size, hidden_size, hidden_size, drop_out=False, nn_act=nn_act)

    def forward(self, graph):
        cell_ = self.cb_encoder(graph.x)
        edge_ = self.eb_encoder(graph.edge_attr)
        return Data(x=cell_, edge_attr=edge_, edge_index=graph.edge_index)


class GnBlock(nn.Module):
    def __init__(self, hidden_size=128, drop_out=False, mp_times=2, MultiHead=1, dual_edge=False, nn_act="SiLU"):
        super().__init__()
        self.dual_edge = dual_edge
        eb_input_dim = 3 * hidden_size
        cb_input_dim = int(hidden_size + 0.5 * hidden_size)
        # construction order (cb first, then eb) fixes the RNG stream of the initial weights
        cb_custom_func = build_mlp(cb_input_dim, hidden_size, hidden_size, drop_out=False, nn_act=nn_act)
        eb_custom_func = build_mlp(eb_input_dim, hidden_size, hidden_size, drop_out=False, nn_act=nn_act)
        self.eb_module = EdgeBlock(custom_func=eb_custom_func, dual_edge=dual_edge)
        self.cb_module = CellBlock(hidden_size, hidden_size, mp_times=mp_times, MultiHead=MultiHead, custom_func=cb_custom_func,
                                   dual_edge=dual_edge)

    def forward(self, graph, graph_node, topo=None):
        """Cell then Edge, residuals added inside the kernels' epilogues (EncoderProcesserDecoder.py:112-124)."""
        if topo is None:
            topo = topology_of(graph, graph_node)
        x, e = fused_gn_block(self, topo, graph.x.contiguous(), graph.edge_attr.contiguous(), inplace=False)
        return Data(x=x, edge_attr=e, edge_index=graph.edge_index)


def fused_gn_block(block: GnBlock, topo, x, e, inplace, node_agg=None, x_prime=None, x_prime_bf16=None, x_prime_split=None,
                   cell_mean_split=None):
    cb, eb = block.cb_module.net, block.eb_module.net
    a = topo.agg(block.cb_module.mp_times)
    node_agg = ops.node_aggregate_fwd(e, a.ptr, a.items, a.rows, out=node_agg)
    if inplace and x_prime_split is not None and cb.math_mode == "bf16x3" and eb.math_mode == "bf16x3":
        # fp32-equivalent fast path: x' lives only as the split fp16 shadow the EdgeBlock gathers (same bits as the two calls below)
        return ops.gn_block_fwd_inplace_split(cb.mlp_ref(), eb.mlp_ref(), x, e, node_agg, a.idx, topo.nc, x_prime_split, cell_mean_split)
    if inplace and x_prime_bf16 is not None and cb.math_mode == "bf16" and eb.math_mode == "bf16" and not (a.idx is not None and a.idx.shape[0] == 4):
        # tensor-core fast path: x' lives only as the bf16 shadow the EdgeBlock gathers (same arithmetic as the two calls below)
        return ops.gn_block_fwd_inplace_bf16(cb.mlp_ref(), eb.mlp_ref(), x, e, node_agg, a.idx, topo.nc, x_prime_bf16)
    idx = a.idx
    if idx is not None and idx.shape[0] == 4:
        # tri / quad hybrid mesh: the k-wide mean once (fvgn_cell_mean_poly), then the per-cell-row convention of the fused kernels
        node_agg, idx = ops.cell_mean_poly(node_agg, idx), None
    if inplace and x_prime_bf16 is not None and cb.math_mode == "bf16" and eb.math_mode == "bf16":
        return ops.gn_block_fwd_inplace_bf16(cb.mlp_ref(), eb.mlp_ref(), x, e, node_agg, idx, topo.nc, x_prime_bf16)
    x_prime, x_new = ops.cell_block_fwd(cb.mlp_ref(), x, node_agg, idx, cb.math_mode_id(), x_prime=x_prime, x_out=x if inplace else None)
    _, e_new = ops.edge_block_fwd(eb.mlp_ref(), x_prime, e, topo.nc, eb.math_mode_id(), e_out=e if inplace else None)
    return x_new, e_new


class Decoder(nn.Module):
    def __init__(self, edge_hidden_size=128, cell_hidden_size=128, edge_output_size=7, cell_output_size=2, cell_input_size=2,
                 dual_edge=False, nn_act="SiLU"):
        super().__init__()
        if dual_edge:
            raise NotImplementedError("dual_edge=True is outside the B200 hot path")
        self.dual_edge = dual_edge
        self.edge_decode_module = build_mlp(edge_hidden_size, edge_hidden_size, edge_output_size, drop_out=False, lay_norm=False, nn_act=nn_act)

    def forward(self, graph=None, features=None, cell_features=None):
        return (False, self.edge_decode_module(graph.edge_attr))


class EncoderProcesserDecoder(nn.Module):
    def __init__(self, message_passing_num, cell_input_size, edge_input_size, cell_output_size, edge_output_size, drop_out,
                 hidden_size=128, mp_times=2, MultiHead=1, dual_edge=False, nn_act="SiLU"):
        super().__init__()
        self.message_passing_num = message_passing_num
        self.dual_edge = dual_edge
        self.encoder = Encoder(edge_input_size=edge_input_size, cell_input_size=cell_input_size, hidden_size=hidden_size, nn_act=nn_act)
        self.processer_list = nn.ModuleList([
            GnBlock(hidden_size=hidden_size, drop_out=drop_out, mp_times=mp_times, MultiHead=MultiHead, dual_edge=dual_edge, nn_act=nn_act)
            for _ in range(message_passing_num)])
        self.decoder = Decoder(edge_hidden_size=hidden_size, cell_hidden_size=hidden_size, cell_output_size=cell_output_size,
                               edge_output_size=edge_output_size, dual_edge=dual_edge, nn_act=nn_act)

    def set_math_mode(self, mode):
        if mode not in MATH_MODES:
            raise ValueError(f"math mode must be one of {sorted(MATH_MODES)}")
        for m in self.modules():
            if isinstance(m, FusedMLP):
                m.math_mode = mode

    def forward(self, graph, graph_node, topo=None, capture=None):
        """(cell_decoded_attr=False, edge_decoded_attr[F,5]) — EncoderProcesserDecoder.py:224-238.  The latents are updated in
        place layer by layer: x / e never leave their two HBM buffers."""
        if topo is None:
            topo = topology_of(graph, graph_node)
        x = self.encoder.cb_encoder(graph.x)
        e = self.encoder.eb_encoder(graph.edge_attr)
        if capture is not None:
            capture["enc_x"], capture["enc_e"] = x.clone(), e.clone()
        mp_times = self.processer_list[0].cb_module.mp_times if len(self.processer_list) else 2
        node_agg = torch.empty((topo.agg(mp_times).rows, 64), dtype=torch.float32, device=x.device)
        all_bf16 = all(b.cb_module.net.math_mode == "bf16" and b.eb_module.net.math_mode == "bf16" for b in self.processer_list)
        all_x3 = all(b.cb_module.net.math_mode == "bf16x3" and b.eb_module.net.math_mode == "bf16x3" for b in self.processer_list)
        x_prime = None if (all_bf16 or all_x3) else torch.empty_like(x)
        x_prime_bf16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if all_bf16 else None
        x_prime_split = torch.empty((x.shape[0], 2, x.shape[1]), dtype=torch.float16, device=x.device) if all_x3 else None
        cell_mean_split = torch.empty((x.shape[0], 2, 64), dtype=torch.float16, device=x.device) if all_x3 else None
        for l, block in enumerate(self.processer_list):
            x, e = fused_gn_block(block, topo, x, e, inplace=True, node_agg=node_agg, x_prime=x_prime, x_prime_bf16=x_prime_bf16,
                                  x_prime_split=x_prime_split, cell_mean_split=cell_mean_split)
            if capture is not None and l in (0, 1, len(self.processer_list) - 1):
                capture[f"x_l{l}"], capture[f"e_l{l}"] = x.clone(), e.clone()
        return (False, self.decoder.edge_decode_module(e))


class Intergrator(nn.Module):
    """Finite-volume face-flux -> cell-divergence integration (EncoderProcesserDecoder.py:241-331) as one kernel."""

    def __init__(self, edge_input_size=7, cell_input_size=2, cell_output_size=2, loss_cont=0):
        super().__init__()
        self.edge_input_size = edge_input_size
        self.cell_input_size = cell_input_size
        self.cell_output_size = cell_output_size
        self.loss_cont = loss_cont

    def forward(self, uv_face=0, p_face=0, flux_D=0, unv=None, rho=None, rhs_coef=None, face_area=None, cell_face=None, device=None,
                decoded=None, cells_face_T=None):
        if decoded is None:  # reference calling convention: three column slices of the decoder output
            decoded = torch.cat((uv_face, p_face, flux_D), dim=-1).contiguous()
        if cells_face_T is None:
            cells_face_T = ops.as_i32(cell_face)
        du, cont = ops.integrator_fwd(decoded, unv.contiguous(), face_area.contiguous(), cells_face_T, rho, rhs_coef,
                                      want_continuity=self.loss_cont > 0)
        return (cont if self.loss_cont > 0 else 0), du
