import torch


def global_mean_pool(x, batch, size=None):
    """Segment mean by `batch` (loss_compute.py:39-75)."""
    if size is None:
        size = int(batch.max().item()) + 1
    s = torch.zeros((size,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    c = torch.zeros(size, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return s / c.clamp(min=1).view(-1, *([1] * (x.dim() - 1)))
