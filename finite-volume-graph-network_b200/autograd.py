"""Training path: torch.autograd.Function wrappers whose forward AND backward run entirely in libfvgn_b200 kernels, so the
reference's `loss.backward(); optimizer.step()` flow (train.py:194-225) works unchanged on this module.

  NetFunction   normalised features -> Encoder -> 15 GnBlocks -> Decoder -> face-area BatchNorm -> Intergrator
                saves only the per-layer INPUTS (x, e, x', node_agg); the backward kernels recompute inside their tiles
  LossFunction  compute_FVM_loss forward / backward
Strict fp32 (FFMA kernels); every reduction is a fixed-order segment sum => gradients are run-to-run deterministic."""
import torch

from . import ops, parallel


def network_mlps(model):
    """FusedMLPs in `model.parameters()` registration order"""
    epd = model.model
    mlps = [epd.encoder.eb_encoder, epd.encoder.cb_encoder]
    for blk in epd.processer_list:
        mlps += [blk.eb_module.net, blk.cb_module.net]
    mlps.append(epd.decoder.edge_decode_module)
    return mlps


def network_params(model):
    """the network's parameters in registration order (+ the face-area BatchNorm pair).  Cached on the model: walking the module tree
    costs ~2 ms per call; the cache is checked against the first / last parameter objects, so a rebuilt module tree rebuilds it."""
    bn = model._grid_feature_normalizer
    cached = model.__dict__.get("_fvgn_param_cache")
    if cached is not None and cached[0] is model.model.encoder.eb_encoder._linears()[0].weight and cached[-2] is bn.weight and cached[-1] is bn.bias \
            and cached[-3] is model.model.decoder.edge_decode_module._linears()[2].bias:
        return cached
    mlps = network_mlps(model)
    params = [p for m in mlps for p in m.parameters()]
    assert [id(p) for p in params] == [id(p) for p in model.model.parameters()], "parameter order drifted from the module tree"
    out = params + [bn.weight, bn.bias]
    model.__dict__["_fvgn_param_cache"] = out
    return out


# The weight-gradient kernels are off the backward's critical path (nothing downstream reads dW before the optimizer): with
# FVGN_WGRAD_STREAM=1 they are enqueued on a side stream, forked from and joined back into the current stream with events (CUDA-graph
# capturable), so that the small memory-bound kernels of the main chain (segment sums, LayerNorm backward, gradient combination) run
# next to them instead of alone.  Inputs of a side-stream launch are kept alive until the main stream has waited for that launch.
import os as _os

WGRAD_SIDE_STREAM = _os.environ.get("FVGN_WGRAD_STREAM", "1") != "0"
# FVGN_EXECUTOR=1 (default): the processor's layer loops of the tensor-core training path are issued by two C calls
# (csrc/executor.cu: fvgn_processor_train_fwd / _bwd) instead of ~13 ctypes calls per GnBlock — same kernels, order, streams and bits
EXECUTOR = _os.environ.get("FVGN_EXECUTOR", "1") != "0"
_SIDE = {}


def _side_stream(dev):
    s = _SIDE.get(dev)
    if s is None:
        s = _SIDE[dev] = torch.cuda.Stream(device=dev)
    return s


class _SideQueue:
    """runs callables on the side stream; `retire()` makes the main stream wait for everything enqueued BEFORE the previous retire() and
    drops those launches' inputs (one-stage delay: the side stream may lag the main chain by up to one layer)"""

    def __init__(self, dev, enabled):
        self.enabled = enabled
        self.side = _side_stream(dev) if enabled else None
        self.main = torch.cuda.current_stream() if enabled else None  # (the stream object costs ~15 us to look up: once per backward)
        self.keep, self.pending = [], None

    def wgrad(self, mlp, *args, alive=(), slot=None, **kw):
        """ops.mlp_wgrad_x3 on the side stream; its output / workspace buffers come from the MAIN stream's allocator pool (a second
        per-stream pool would have to grow through its own cudaMalloc phase whenever the batch sizes change)"""
        if not self.enabled:
            if slot is not None:
                kw["buffers"] = (slot, ops.mlp_wgrad_x3_buffers(mlp, mlp.tensors[0].device)[1])
            return ops.mlp_wgrad_x3(mlp, *args, **kw)
        bufs = ops.mlp_wgrad_x3_buffers(mlp, mlp.tensors[0].device)
        if slot is not None:  # data parallel: the gradients land in their slot of the flat all-reduce buffer (parallel.OverlappedGradReducer)
            bufs = (slot, bufs[1])
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            out = ops.mlp_wgrad_x3(mlp, *args, buffers=bufs, **kw)
        self.keep.extend(alive)
        self.keep.append(bufs[1])
        return out

    def retire(self, final=False):
        if not self.enabled:
            return
        main = self.main
        if self.pending is not None:
            main.wait_event(self.pending[0])
            self.pending[1].clear()
        ev = torch.cuda.Event()
        ev.record(self.side)
        self.pending, self.keep = (ev, self.keep), []
        if final:
            main.wait_event(ev)
            self.pending[1].clear()
            self.pending = None


class NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, topo, cell_in, edge_in, graph_edge_x, cell_area, unv, rho, dt, *params):
        epd = model.model
        enc_c, enc_e = epd.encoder.cb_encoder, epd.encoder.eb_encoder
        # forward arithmetic: strict FFMA ("fp32": FFMA backward that recomputes each tile from the saved layer inputs), the
        # fp32-equivalent tensor-core mode ("bf16x3") or plain bf16 tiles ("bf16"): tensor-core backward from the saved pre-activations
        mm = {"fp32": 0, "bf16x3": 2, "bf16": 2}[model.math_mode]  # "bf16": the bf16x3 kernels with the leading split product only
        # bf16x3: the forward also stores every MLP's pre-activations (z1 | z2 | z3, 1.5 KB per row): the backward reads them
        # instead of recomputing the forward in its tiles (a third of its FLOPs)
        zbuf = (lambda rows: ops.bucketed_empty((rows, 384), cell_in.device)) if mm == 2 else (lambda rows: None)
        z_enc_c, z_enc_e = zbuf(topo.C), zbuf(topo.F)
        # (inside a CUDA-graph capture the host cost of the per-call path is irrelevant, and its backward interleaves the weight-gradient side
        # stream and the data-parallel reduce groups layer by layer instead of joining the streams at the end of every chunk)
        # (tri / quad tables: per-call path — the k-wide mean and its transpose are separate small kernels there)
        use_exec = EXECUTOR and mm == 2 and len(epd.processer_list) > 0 and not torch.cuda.is_current_stream_capturing() and not topo.poly
        packed = None
        if use_exec:
            L, dev = len(epd.processer_list), cell_in.device
            f32 = lambda *shape: ops.bucketed_empty(shape, dev)
            mp = epd.processer_list[0].cb_module.mp_times
            X, E = f32(L + 1, topo.C, 128), f32(L + 1, topo.F, 128)
            XP, NA, ZC, ZE = f32(L, topo.C, 128), f32(L, topo.C, 64), f32(L, topo.C, 384), f32(L, topo.F, 384)
            ops.mlp_rows_fwd(enc_c.mlp_ref(), cell_in, mm, out=X[0], z_save=z_enc_c)
            ops.mlp_rows_fwd(enc_e.mlp_ref(), edge_in, mm, out=E[0], z_save=z_enc_e)
            cms = [blk.cb_module.net.mlp_ref() for blk in epd.processer_list]
            ems = [blk.eb_module.net.mlp_ref() for blk in epd.processer_list]
            ops.processor_train_fwd(cms, ems, mp, topo.train_tables(mp), X, E, XP, NA, ZC, ZE, f32(topo.N, 64) if mp >= 2 else None)
            x, e = X[L], E[L]
            packed = (X, E, XP, NA, ZC, ZE, mp)
        else:
            x = ops.mlp_rows_fwd(enc_c.mlp_ref(), cell_in, mm, z_save=z_enc_c)
            e = ops.mlp_rows_fwd(enc_e.mlp_ref(), edge_in, mm, z_save=z_enc_e)
        saved = []
        for blk in ([] if use_exec else epd.processer_list):
            cb, eb = blk.cb_module.net, blk.eb_module.net
            ag = topo.agg(blk.cb_module.mp_times)
            na = ops.node_aggregate_fwd(e, ag.ptr, ag.items, ag.rows)
            idx_k = ag.idx  # what the fused CellBlock kernel gathers through (None: `na` is the per-cell row already)
            if ag.idx is not None and (mm == 2 or topo.poly):
                # tensor-core training: the per-cell mean once per layer (same arithmetic), then the forward and the weight-gradient
                # kernel read it as a plain per-cell row (cells_node_T = NULL) instead of three gathered node rows each.
                # tri / quad tables: the k-wide mean (fvgn_cell_mean_poly) in every math mode
                na = ops.cell_mean_poly(na, ag.idx) if topo.poly else ops.cell_mean(na, ag.idx)
                idx_k = None
            zc, ze = zbuf(topo.C), zbuf(topo.F)
            xp, x_new = ops.cell_block_fwd(cb.mlp_ref(), x, na, idx_k, mm, z_save=zc)
            _, e_new = ops.edge_block_fwd(eb.mlp_ref(), xp, e, topo.nc, mm, z_save=ze)
            saved.append((x, e, xp, na, zc, ze))
            x, e = x_new, e_new
        z_dec = zbuf(topo.F)
        decoded = ops.mlp_rows_fwd(epd.decoder.edge_decode_module.mlp_ref(), e, mm, z_save=z_dec)
        bn = model._grid_feature_normalizer
        use_batch = model.training or not bn.track_running_stats
        if use_batch and parallel.is_distributed():
            # cross-rank batch statistics: 3 floats all-reduced, then the kernel applies the given stats
            red = model.__dict__.get("_dp_bn_reduced")  # [sum, sumsq, count] already reduced with the normaliser statistics (FVGN.forward)
            sums = None if red is not None else ops.face_area_sums(graph_edge_x, cell_area, topo.nc, dt)[1]
            gstats, mean, var, n = parallel.global_bn_stats(sums, topo.F, reduced=red)
            model.__dict__["_dp_bn_reduced"] = None
            parallel.update_running_stats_(bn, mean, var, n)
            face_area, raw, stats = ops.face_area_bn_fwd(graph_edge_x, cell_area, topo.nc, dt, True, bn.weight.detach(), bn.bias.detach(),
                                                         bn.running_mean, bn.running_var, want_raw=True, stats=gstats)
        else:
            face_area, raw, stats = ops.face_area_bn_fwd(graph_edge_x, cell_area, topo.nc, dt, use_batch, bn.weight.detach(), bn.bias.detach(),
                                                         bn.running_mean, bn.running_var, want_raw=True)
        if model.training:
            bn.num_batches_tracked += 1
        delta_u, _ = ops.integrator_fwd(decoded, unv, face_area, topo.cfT, rho, 1.0)
        ctx.model, ctx.topo, ctx.rho = model, topo, rho
        ctx.saved = saved
        ctx.packed = packed
        # detached aliases: storing the OUTPUT tensors themselves on ctx would close a reference cycle (output -> grad_fn -> ctx ->
        # output) that only the cyclic GC can free, i.e. ~2 GB per layer of saved latents would pile up between collections
        ctx.aux = (cell_in.detach(), edge_in.detach(), e, decoded.detach(), face_area.detach(), raw, stats, unv.detach())
        ctx.zs = (z_enc_c, z_enc_e, z_dec)
        ctx.mark_non_differentiable(face_area)
        return decoded, delta_u, face_area

    @staticmethod
    def backward(ctx, g_decoded, g_delta_u, _g_fa):
        model, topo = ctx.model, ctx.topo
        epd = model.model
        cell_in, edge_in, e_final, decoded, face_area, raw, stats, unv = ctx.aux
        saved_layers, ctx.saved, ctx.aux = ctx.saved, None, None  # release the saved latents as the backward sweep consumes them
        packed, ctx.packed = ctx.packed, None
        (z_enc_c, z_enc_e, z_dec), ctx.zs = ctx.zs, None
        bt = topo.backward_tables()
        dev = decoded.device
        g_dec = (g_decoded.contiguous().clone() if g_decoded is not None else torch.zeros_like(decoded))
        if g_delta_u is None:
            g_delta_u = torch.zeros((topo.C, 2), dtype=torch.float32, device=dev)
        # Intergrator^T (p_face is detached in the forward, FVGN.py:231) and the BatchNorm affine pair
        fptr, fitems = bt["face_from_cell"]
        g_fa = ops.integrator_bwd_(g_dec, decoded, unv, face_area, g_delta_u.contiguous(), fptr, fitems, topo.C, ctx.rho, 1.0)
        g_bn = ops.face_area_bn_bwd(g_fa, raw, stats)
        grads = {}
        dec = epd.decoder.edge_decode_module
        tc = z_dec is not None  # bf16x3 forward: the data-gradient chain runs on the tensor cores, the FFMA kernel keeps the TN products
        sq = _SideQueue(dev, WGRAD_SIDE_STREAM and tc)
        red = model.__dict__.get("_dp_reducer") if (tc and parallel.is_distributed()) else None
        slot = (lambda m: red.slot(m)) if red is not None else (lambda m: None)
        red_stream = (sq.side if sq.enabled else torch.cuda.current_stream()) if red is not None else None
        if tc:
            ops.ensure_packed_bwd_many([m.mlp_ref() for m in network_mlps(model)])  # one launch instead of one per MLP before its dgrad
            da, g_e = ops.mlp_dgrad_x3(dec.mlp_ref(), g_dec, z_dec, want_dA=True)
            grads[id(dec)] = sq.wgrad(dec.mlp_ref(), 0, g_dec, None, None, z_dec, da, x=e_final, alive=(g_dec, z_dec, da, e_final), slot=slot(dec))
        else:
            grads[id(dec)], g_e = ops.mlp_rows_bwd(dec.mlp_ref(), e_final, g_dec, want_d_in=True, z_saved=z_dec)
        g_x = None  # the last layer's cell latent feeds nothing but its own EdgeBlock
        n_done = 0
        cptr, citems = bt["cell_from_face"]
        nptr, nitems = bt["node_from_cell"]
        if packed is not None:
            # the whole layer loop from C (csrc/executor.cu), in chunks that end where a data-parallel reduce group ends
            import ctypes as _C

            X, E, XP, NA, ZC, ZE, mp = packed
            L = len(epd.processer_list)
            cms = [blk.cb_module.net.mlp_ref() for blk in epd.processer_list]
            ems = [blk.eb_module.net.mlp_ref() for blk in epd.processer_list]
            lib = ops._lib.lib()
            if red is not None:
                slots = [s for blk in epd.processer_list for s in (red.slot(blk.eb_module.net), red.slot(blk.cb_module.net))]
            else:
                sizes = [int(lib.fvgn_mlp_grad_floats(m.k_in)) for pair in zip(ems, cms) for m in pair]
                G = ops.bucketed_empty((sum(sizes),), dev)
                slots, o = [], 0
                for n in sizes:
                    slots.append(G[o:o + n])
                    o += n
            ptrs = (_C.c_void_p * (2 * L))(*[s.data_ptr() for s in slots])
            scratch = ops.bucketed_empty((int(lib.fvgn_processor_train_bwd_scratch_floats(topo.C, topo.F, topo.N)),), dev)
            bounds = sorted({0, L} | ({L - nd for nd in red.cuts} if red is not None else set()))
            g_e = g_e.contiguous()
            ge_bufs = [ops.bucketed_empty((topo.F, 128), dev) for _ in range(2)]
            gx_bufs = [ops.bucketed_empty((topo.C, 128), dev) for _ in range(2)]
            side = sq.side if sq.enabled else None
            k = 0
            for hi, lo in zip(reversed(bounds[1:]), reversed(bounds[:-1])):
                ops.processor_train_bwd(cms, ems, lo, hi, mp, topo.train_tables(mp), X, E, XP, NA, ZC, ZE, g_e, ge_bufs[k & 1], g_x, gx_bufs[k & 1],
                                        ptrs, scratch, side)
                g_e, g_x = ge_bufs[k & 1], gx_bufs[k & 1]
                k += 1
                if red is not None:
                    red.layer_done(L - lo, red_stream)
            for l, blk in enumerate(epd.processer_list):
                grads[id(blk.eb_module.net)] = ops._mlp_grad_views(ems[l], slots[2 * l])
                grads[id(blk.cb_module.net)] = ops._mlp_grad_views(cms[l], slots[2 * l + 1])
            del X, E, XP, NA, ZC, ZE, packed
        for blk in (reversed(epd.processer_list) if saved_layers else []):
            x_in, e_in, xp, na, zc, ze = saved_layers.pop()
            cb, eb = blk.cb_module.net, blk.eb_module.net
            ag = topo.agg(blk.cb_module.mp_times)
            if tc:
                dz, gxh = ops.ln_bwd(g_e, ze, eb[1].weight)
                da, dA_e = ops.mlp_dgrad_x3(eb.mlp_ref(), dz, ze, want_dA=True)
                grads[id(eb)] = sq.wgrad(eb.mlp_ref(), 2, g_e, dz, gxh, ze, da, src0=xp, src1=e_in, idx=topo.nc,
                                         alive=(g_e, dz, gxh, ze, da, xp, e_in), slot=slot(eb))
            else:
                grads[id(eb)], dA_e = ops.edge_block_bwd(eb.mlp_ref(), xp, e_in, topo.nc, g_e, z_saved=ze)
            # d x' = residual path + transpose of the x'[s], x'[r] gathers (cell <- face segment sum)
            g_xp = ops.segment_sum_rows(dA_e, topo.F, 0, 128, 128, cptr, citems, topo.C, base=g_x, scale=1.0)
            if tc:
                dz, gxh = ops.ln_bwd(g_xp, zc, cb[1].weight)
                da, dA_c = ops.mlp_dgrad_x3(cb.mlp_ref(), dz, zc, want_dA=True, dA_base=g_x)
                grads[id(cb)] = sq.wgrad(cb.mlp_ref(), 1, g_xp, dz, gxh, zc, da, src0=x_in, src1=na, idx=None,  # na = the per-cell aggregate
                                         alive=(g_xp, dz, gxh, zc, da, x_in, na), slot=slot(cb))
            else:
                grads[id(cb)], dA_c = ops.cell_block_bwd(cb.mlp_ref(), x_in, na, None if topo.poly else ag.idx, g_xp, d_x_base=g_x, z_saved=zc)
            # transposes of the twice-message aggregation: cell -> node (x 1/3), node -> face halves
            if ag.idx is not None and topo.poly:
                # tri / quad tables: the mean's 1/k differs from cell to cell — scale the rows, then the plain segment sum over the
                # CSR whose -1 entries sit in an extra row that is not walked
                dA_c[:, 128:192].mul_(topo.inv_k())
                g_na = ops.segment_sum_rows(dA_c, topo.C, 128, 0, 64, nptr, nitems, topo.N, base=None, scale=1.0)
            elif ag.idx is not None:
                g_na = ops.segment_sum_rows(dA_c, topo.C, 128, 0, 64, nptr, nitems, topo.N, base=None, scale=1.0 / 3.0)
            else:  # mp_times = 1: the aggregate is per cell and enters the MLP as it is; its transpose is the identity
                g_na = dA_c[:, 128:192].contiguous()
            g_e = ops.edge_grad_combine(g_e, dA_e, g_na, ag.face)
            g_x = dA_c[:, 0:128]
            sq.retire()
            if red is not None:
                n_done += 1
                red.layer_done(n_done, red_stream)
        for enc, inp, g, z in ((epd.encoder.cb_encoder, cell_in, g_x, z_enc_c), (epd.encoder.eb_encoder, edge_in, g_e, z_enc_e)):
            if tc:
                dz, gxh = ops.ln_bwd(g, z, enc[1].weight)
                da, _ = ops.mlp_dgrad_x3(enc.mlp_ref(), dz, z, want_dA=False)
                grads[id(enc)] = sq.wgrad(enc.mlp_ref(), 0, g, dz, gxh, z, da, x=inp, alive=(g, dz, gxh, z, da, inp), slot=slot(enc))
            else:
                grads[id(enc)], _ = ops.mlp_rows_bwd(enc.mlp_ref(), inp, g, z_saved=z)
        if red is not None:
            # the BatchNorm pair joins the tail group; the copy is ordered before the tail's all-reduce through the side stream's fork
            red.bn_slot.copy_(g_bn)
            if sq.enabled:
                sq.side.wait_stream(sq.main)
            red.finish(red_stream)
            g_bn = red.bn_slot
        sq.retire(final=True)
        flat = [g for m in network_mlps(model) for g in grads[id(m)]]
        flat += [g_bn[0:1], g_bn[1:2]]
        return (None,) * 9 + tuple(flat)


class LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_du, target_du, pred_edge, target_face, mask, cell_batch, face_batch, n_graphs, mode, w_mom, w_uv, w_p):
        m8 = mask.to(torch.uint8).contiguous()
        pred_du, target_du = pred_du.contiguous(), target_du.contiguous()
        pred_edge, target_face = pred_edge.contiguous(), target_face.contiguous()
        # data parallel + direct_mean_loss: the numerators / denominators of the three batch means are summed over the ranks before the
        # log, so every rank holds the single-process loss; the per-element gradient then is the single-process one,
        # and since the gradient bucket is AVERAGED over the ranks the incoming gradient is scaled by the world size in backward
        dp = mode == 0 and parallel.is_distributed()
        out, ws = ops.loss_fwd(pred_du, target_du, pred_edge, target_face, m8, cell_batch, face_batch, n_graphs, mode, w_mom, w_uv, w_p,
                               reduce_partials=parallel.all_reduce_sum_ if dp else None)
        ctx.args = (pred_du, target_du, pred_edge, target_face, m8, cell_batch, face_batch, n_graphs, mode, w_mom, w_uv, w_p, ws)
        ctx.gscale = float(parallel.world_size()) if dp else 1.0
        return out

    @staticmethod
    def backward(ctx, g_out):
        pred_du, target_du, pred_edge, target_face, m8, cb, fb, n_graphs, mode, w_mom, w_uv, w_p, ws = ctx.args
        g_loss = g_out[0:1].contiguous()  # only out[0] (the loss) is a differentiable quantity; stays on the device
        if ctx.gscale != 1.0:
            g_loss = g_loss * ctx.gscale
        g_du, g_edge = ops.loss_bwd(pred_du, target_du, pred_edge, target_face, m8, cb, fb, n_graphs, mode, w_mom, w_uv, w_p, g_loss, ws)
        return (g_du, None, g_edge) + (None,) * 9
