"""Data-parallel training over mesh graphs (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The reference has no distributed code at all; what DP needs in order to reproduce the
single-process batch is:

  1. one all-reduce(SUM)/world of the flat fp32 gradient bucket (2,210,695 floats = 8.84 MB) per step
  2. all-reduce(SUM) of the four online normalisers' per-step INCREMENTS (46 floats + 4 counts) so every rank holds the
     statistics of the whole batch (utils/normalization.py:56-67)
  3. all-reduce(SUM) of the face-area BatchNorm (sum, sum of squares, count) so the batch statistics are global (FVGN.py:209-227)
  4. `direct_mean_loss` (run_train.sh:17, utils/loss_compute.py:92-113: ONE log over the batch means): all-reduce(SUM) of the three
     numerators / denominators BEFORE the log (autograd.LossFunction -> ops.loss_fwd(reduce_partials=...)), gradient scaled by the
     world size to undo the bucket's averaging
With `global_mean_*` losses (per-graph log, mean over graphs) and the same number of graphs per rank, DP-N equals the
single-process batch without item 4.  There is no collective
in inference/rollout: ranks are independent replicas."""
import torch
import torch.distributed as dist

NORMALIZERS = ("_acc_output_normalizer", "_face_uvp_flux_output_normalizer", "_edge_normalizer", "_cell_normalizer")
_NORM_BUFS = ("acc_sum", "acc_sum_squared", "acc_count", "num_accumulations")


_DISABLED = [0]


def is_distributed():
    return not _DISABLED[0] and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class single_process:
    """context: treat this rank as a single process (no statistic / loss / gradient collectives) although a process group exists —
    used to compute the single-process batch-(N x b) result that DP-N must reproduce (bench.py --check, tests)"""

    def __enter__(self):
        _DISABLED[0] += 1

    def __exit__(self, *a):
        _DISABLED[0] -= 1
        return False


def world_size():
    return dist.get_world_size() if is_distributed() else 1


def all_reduce_sum_(t, group=None):
    """in-place SUM over the ranks (identity in a single process)"""
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class GradBucket:
    """Flat fp32 bucket over all parameters: one collective per step."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def all_reduce_mean_(self, group=None):
        if not is_distributed():
            return
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.empty(self.numel, dtype=torch.float32, device=dev)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        # gather / scatter with one multi-tensor copy each (two launches per step instead of one per parameter)
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.div_(dist.get_world_size(group))
        if have:
            torch._foreach_copy_([g for _, g in have], [v for v, _ in have])
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()


class OverlappedGradReducer:
    """Gradient all-reduce OVERLAPPED with the backward (tensor-core training path).  The weight-gradient kernels of every MLP write
    straight into their slot of one flat fp32 buffer laid out in BACKWARD order (decoder, GnBlock 14 ... 0, encoders, BatchNorm pair);
    as soon as the weight gradients of a group of layers have been enqueued on the weight-gradient side stream, that group's slice is
    all-reduced (AVG) from that stream — NCCL runs it on its own stream, next to the main chain's dgrad / LayerNorm-backward / segment-sum
    kernels of the following layers — and the parameters' .grad tensors are views of the reduced buffer: no gather / scatter copies, no
    division kernel, one exposed collective (the last group) instead of the whole 8.84 MB bucket at the tail of the step.
    Attach with `attach(model)`; autograd.NetFunction.backward drives it; CUDA-graph capturable."""

    def __init__(self, model, n_groups=4, group=None, cuts=None):
        from . import _lib
        from .autograd import network_mlps

        mlps = network_mlps(model)  # registration order: enc_e, enc_c, (eb_l, cb_l) x L, decoder
        enc_e, enc_c, dec = mlps[0], mlps[1], mlps[-1]
        layers = [(mlps[2 + 2 * l], mlps[3 + 2 * l]) for l in range((len(mlps) - 3) // 2)]
        order = [dec] + [m for eb, cb in reversed(layers) for m in (eb, cb)] + [enc_c, enc_e]
        L = _lib.lib()
        sizes = [int(L.fvgn_mlp_grad_floats(m.mlp_ref().k_in)) for m in order]
        dev = next(model.parameters()).device
        self.flat = torch.zeros(sum(sizes) + 4, dtype=torch.float32, device=dev)
        self.slots, off = {}, 0
        offs = []
        for m, n in zip(order, sizes):
            self.slots[id(m)] = self.flat[off:off + n]
            offs.append(off)
            off += n
        self.bn_slot = self.flat[off:off + 2]
        self.total = off + 4
        # group boundaries (number of GnBlocks finished, counted from the last layer): every ceil(L / n_groups) blocks, plus one just before
        # the first layer so that the EXPOSED tail group (nothing is left to overlap it with) is one GnBlock + the encoders + the BN pair
        per = max(1, -(-len(layers) // n_groups))
        self.cuts = {}  # number of GnBlocks finished -> (start, end) float range to reduce
        start = 0
        nl = len(layers)
        marks = sorted(set(cuts) if cuts is not None else set(range(per, nl, per)) | ({nl - 1} if nl > 1 else set()))
        for done in marks:
            if not 0 < done < nl:
                continue
            end = offs[1 + 2 * done]  # first slot of the next (not yet finished) layer
            self.cuts[done] = (start, end)
            start = end
        self.tail = (start, self.total)
        self.group = group
        self.works = []
        self.model_id = id(model)

    def attach(self, model):
        model.__dict__["_dp_reducer"] = self
        return self

    def slot(self, mlp_module):
        return self.slots[id(mlp_module)]

    def _reduce(self, rng, stream):
        if not is_distributed() or rng[1] <= rng[0]:
            return
        with torch.cuda.stream(stream):
            self.works.append(dist.all_reduce(self.flat[rng[0]:rng[1]], op=dist.ReduceOp.AVG, group=self.group, async_op=True))

    def layer_done(self, n_done, stream):
        """called after the weight gradients of the n_done-th GnBlock (counting from the last layer) were enqueued on `stream`"""
        rng = self.cuts.get(n_done)
        if rng is not None:
            self._reduce(rng, stream)

    def finish(self, stream):
        """the encoders' weight gradients and the BN pair are in: reduce the tail, then make the current stream wait for every group"""
        self._reduce(self.tail, stream)
        for w in self.works:
            w.wait()
        self.works = []


def _norm_tensors(model):
    """(the 12 accumulated tensors [acc_sum, acc_sum_squared, acc_count] x 4 normalisers, the 4 num_accumulations) in a fixed order"""
    acc = [getattr(getattr(model, n), b) for n in NORMALIZERS for b in ("acc_sum", "acc_sum_squared", "acc_count")]
    num = [getattr(model, n).num_accumulations for n in NORMALIZERS]
    return acc, num


def snapshot_normalizers(model):
    """copies of the normaliser buffers before a training forward: two multi-tensor launches (this runs inside the captured training step:
    one clone per buffer was 16 launches)"""
    acc, num = _norm_tensors(model)
    return {"acc": torch._foreach_add(acc, 0.0), "num": torch._foreach_add(num, 0.0)}


def sync_normalizer_increments_(model, before, group=None, extra=None):
    """after a training forward: buffers <- before + sum over ranks of (after - before); num_accumulations advances by one.
    `extra`: optional float64 vector that rides along in the same collective (the face-area BatchNorm's [sum, sum of squares, count]:
    one statistic all-reduce per step instead of two); returns its reduced values.  A handful of multi-tensor launches + one collective."""
    if not is_distributed():
        return extra
    acc, num = _norm_tensors(model)
    inc = torch._foreach_sub(acc, before["acc"])
    flat = torch.cat([t.reshape(-1) for t in inc])
    n_norm = flat.numel()
    if extra is not None:
        flat = torch.cat([flat.to(torch.float64), extra.to(torch.float64).reshape(-1)])  # (fp32 increments are exact in fp64)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    extra_out = flat[n_norm:] if extra is not None else None
    red = flat[:n_norm].to(torch.float32) if extra is not None else flat
    pieces, off = [], 0
    for t in acc:
        k = t.numel()
        pieces.append(red[off:off + k].view_as(t))
        off += k
    torch._foreach_copy_(acc, torch._foreach_add(before["acc"], pieces))
    # every rank accumulated once (or none did): keep the single-process count
    torch._foreach_copy_(num, torch._foreach_maximum(num, before["num"]))
    return extra_out


def global_bn_stats(sums, count, eps=1e-5, group=None, reduced=None):
    """local float64 [sum, sumsq] + count -> (stats=[mean, invstd] fp32, mean, biased var, total count) over all ranks
    (`reduced`: the already all-reduced [sum, sumsq, count], when it rode along with the normaliser statistics)"""
    # (torch.full, not torch.tensor: no host -> device copy, so the step stays CUDA-graph capturable)
    if reduced is not None:
        t = reduced
    else:
        t = torch.cat([sums.to(torch.float64), torch.full((1,), float(count), dtype=torch.float64, device=sums.device)])
        if is_distributed():
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    n = t[2]
    mean = t[0] / n
    var = torch.clamp(t[1] / n - mean * mean, min=0.0)
    stats = torch.stack([mean, 1.0 / torch.sqrt(var + eps)]).to(torch.float32)
    return stats, mean, var, n


def update_running_stats_(bn, mean, var, n, momentum=0.1):
    unbiased = var * (n / torch.clamp(n - 1.0, min=1.0))
    bn.running_mean.mul_(1.0 - momentum).add_(momentum * mean.to(torch.float32))
    bn.running_var.mul_(1.0 - momentum).add_(momentum * unbiased.to(torch.float32))
