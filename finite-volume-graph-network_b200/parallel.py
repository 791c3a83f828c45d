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


def snapshot_normalizers(model):
    return {n: {b: getattr(getattr(model, n), b).clone() for b in _NORM_BUFS} for n in NORMALIZERS}


def sync_normalizer_increments_(model, before, group=None):
    """after a training forward: buffers <- before + sum over ranks of (after - before); num_accumulations advances by one."""
    if not is_distributed():
        return
    parts = []
    for n in NORMALIZERS:
        m = getattr(model, n)
        for b in ("acc_sum", "acc_sum_squared", "acc_count"):
            parts.append((getattr(m, b) - before[n][b]).reshape(-1))
    flat = torch.cat(parts)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for n in NORMALIZERS:
        m = getattr(model, n)
        for b in ("acc_sum", "acc_sum_squared", "acc_count"):
            t = getattr(m, b)
            k = t.numel()
            t.copy_(before[n][b] + flat[off:off + k].view_as(t))
            off += k
        # every rank accumulated once (or none did): keep the single-process count
        m.num_accumulations.copy_(torch.maximum(m.num_accumulations, before[n]["num_accumulations"]))


def global_bn_stats(sums, count, eps=1e-5, group=None):
    """local float64 [sum, sumsq] + count -> (stats=[mean, invstd] fp32, mean, biased var, total count) over all ranks"""
    # (torch.full, not torch.tensor: no host -> device copy, so the step stays CUDA-graph capturable)
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
