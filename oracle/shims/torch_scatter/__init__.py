"""scatter_add == zeros.index_add_ (sequential, deterministic on CPU).  GN_blocks.py:117,133."""
import torch


def scatter_add(src, index, dim=0, dim_size=None):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    return torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)


def scatter_mean(src, index, dim=0, dim_size=None):
    s = scatter_add(src, index, dim, dim_size)
    c = scatter_add(torch.ones_like(index, dtype=src.dtype), index, 0, s.shape[0]).clamp(min=1)
    return s / c.view(-1, *([1] * (src.dim() - 1)))
