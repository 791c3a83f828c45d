"""Tensor-level wrappers over the C-ABI (include/fvgn_b200.h).  torch is used for device memory and streams only;
all arithmetic happens in libfvgn_b200.so on the current CUDA stream.  CPU tensors are an error (no fallback)."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import MlpT, check

HID = 128

# bench instrumentation: number of kernels launched through this module and optional per-kernel CUDA-event timing
LAUNCHES = 0
TIMING = None  # None, or {"edge_block": [(ev0, ev1), ...], ...}


def _count(n):
    global LAUNCHES
    LAUNCHES += n


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMING is not None and self.name in TIMING:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if TIMING is not None and self.name in TIMING:
            self.e1.record()
            TIMING[self.name].append((self.e0, self.e1))
        return False


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _stream():
    """the current CUDA stream of the current device as a void* (called once per kernel launch: the raw getter costs ~0.3 us, the
    torch.cuda.current_stream() object ~16 us)"""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(_raw_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name, dim=None, rowmajor=True):
    try:  # fast accept (one validator call per tensor argument of every launch): a contiguous CUDA tensor of the right dtype / rank
        if t.is_cuda and t.dtype is dtype and (dim is None or t.dim() == dim) and t.is_contiguous():
            if t.device.index != _raw_device():
                raise RuntimeError(f"{name}: tensor lives on {t.device} but the current CUDA device is {_raw_device()} (kernels launch on "
                                   "the current device's stream: wrap the call in torch.cuda.device(tensor.device))")
            return t
    except AttributeError:
        pass
    if not torch.is_tensor(t):
        raise TypeError(f"{name}: expected a tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (fvgn_b200 has no CPU path), got {t.device}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise ValueError(f"{name}: expected {dim} dims, got shape {tuple(t.shape)}")
    if rowmajor and t.dim() >= 1 and t.numel() > 0 and t.stride(-1) != 1 and t.shape[-1] != 1:
        raise ValueError(f"{name}: innermost dimension must be contiguous (strides {t.stride()})")
    return t


def _ci(t, name, shape0=None):
    """int32, C-contiguous index table (kernels index it as plain row-major memory)"""
    _req(t, torch.int32, name)
    if not t.is_contiguous():
        raise ValueError(f"{name}: index table must be C-contiguous (strides {t.stride()})")
    if shape0 is not None and t.shape[0] != shape0:
        raise ValueError(f"{name}: expected leading dimension {shape0}, got shape {tuple(t.shape)}")
    return t


def _cf(t, name):
    _req(t, torch.float32, name)
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be C-contiguous (strides {t.stride()})")
    return t


def _ld(t):
    """row stride of a 2-D (possibly sliced) row-major tensor"""
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def bucketed_empty(shape, device, dtype=torch.float32):
    """torch.empty(shape) whose ALLOCATION is rounded up to a size bucket (1/8 of the next power of two): with a new batch topology every
    step (train.py:119-160) the big per-step arrays would otherwise each have a new size, miss the caching allocator's free blocks and pay
    cudaMalloc / cudaFree (a device synchronisation) every step"""
    n = 1
    for d in shape:
        n *= int(d)
    g = max(1 << 18, (1 << max(n - 1, 1).bit_length()) >> 3)
    padded = (n + g - 1) // g * g
    return torch.empty(padded, dtype=dtype, device=device)[:n].view(shape)


def as_i32(t):
    """The reference keeps int64 tables (Load_mesh.py:506-507,531-533,542); the kernels take int32."""
    return t.to(torch.int32).contiguous()


# ------------------------------------------------------------------------------------------------ MLP weights
# Weight-change detection for the packed tensor-core tiles: the tensors' version counters catch in-place updates made through
# ordinary ops (copy_, load_state_dict, the foreach / single-tensor optimizers), but torch's FUSED optimizers (AdamW(fused=True) ...)
# update parameters without bumping `_version`.  A process-wide optimizer-step hook therefore advances an epoch that is part of the
# pack key: any optimizer step re-packs at the next forward (a few microseconds per MLP, on the device).
_WEIGHTS_EPOCH = [0]


def _on_optimizer_step(optimizer, args, kwargs):
    _WEIGHTS_EPOCH[0] += 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook  # noqa: E402

_register_step_hook(_on_optimizer_step)


def _weights_key(tensors):
    return (_WEIGHTS_EPOCH[0],) + tuple(t._version for t in tensors[:6:2])


class MlpRef:
    """Borrowed pointers to one build_mlp's parameters (reference state_dict layout)."""

    __slots__ = ("tensors", "struct", "k_in", "n_out", "packed", "packed_versions", "packed3", "packed3_versions", "packed3b",
                 "packed3b_versions")

    def __init__(self, w1, b1, w2, b2, w3, b3, ln_w=None, ln_b=None):
        ts = [w1, b1, w2, b2, w3, b3] + ([ln_w, ln_b] if ln_w is not None else [])
        for i, t in enumerate(ts):
            _req(t, torch.float32, f"mlp tensor {i}")
            if not t.is_contiguous():
                raise ValueError("mlp parameters must be contiguous")
        if w1.shape[0] != HID or tuple(w2.shape) != (HID, HID) or w3.shape[1] != HID:
            raise ValueError("build_mlp hidden width must be 128 (the reference ignores --hidden_size)")
        self.tensors = ts  # keep alive
        self.k_in, self.n_out = int(w1.shape[1]), int(w3.shape[0])
        self.struct = MlpT(w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(),
                           ln_w.data_ptr() if ln_w is not None else None, ln_b.data_ptr() if ln_b is not None else None,
                           self.k_in, self.n_out, None, None, None, None, None, 0)
        self.packed3b = None
        self.packed3b_versions = None
        self.packed = None
        self.packed_versions = None
        self.packed3 = None
        self.packed3_versions = None

    def ensure_packed_bwd(self):
        """transposed 3-way-split weight tiles of the tensor-core data-gradient chain (fvgn_mlp_dgrad_x3)"""
        ver = _weights_key(self.tensors)
        if self.packed3b is None or self.packed3b_versions != ver:
            L = _lib.lib()
            nbytes = L.fvgn_mlp_packed_x3_bwd_bytes(self.k_in)
            if self.packed3b is None:
                self.packed3b = torch.empty(nbytes, dtype=torch.uint8, device=self.tensors[0].device)
                self.struct.packed_x3_bwd = self.packed3b.data_ptr()
            _count(2 + (self.k_in + 127) // 128)
            check(L.fvgn_mlp_pack_bf16x3_bwd(C.byref(self.struct), _p(self.packed3b), nbytes, _stream()), "mlp_pack_bf16x3_bwd")
            self.packed3b_versions = ver
        return self

    def ensure_packed(self, math_mode=1):
        """(Re)build the swizzled weight tiles the tcgen05 paths read (math_mode 1: bf16; 2: exact 3-way bf16 split), if the fp32
        weights changed since the last pack."""
        ver = _weights_key(self.tensors)
        if math_mode == 2:
            if self.packed3 is None or self.packed3_versions != ver:
                L = _lib.lib()
                nbytes = L.fvgn_mlp_packed_x3_bytes(self.k_in)
                if self.packed3 is None:
                    self.packed3 = torch.empty(nbytes, dtype=torch.uint8, device=self.tensors[0].device)
                    self.struct.packed_x3 = self.packed3.data_ptr()
                _count(3)
                check(L.fvgn_mlp_pack_bf16x3(C.byref(self.struct), _p(self.packed3), nbytes, _stream()), "mlp_pack_bf16x3")
                self.packed3_versions = ver
            return self
        if self.packed is None or self.packed_versions != ver:
            L = _lib.lib()
            nbytes = L.fvgn_mlp_packed_bytes(self.k_in)
            if self.packed is None:
                self.packed = torch.empty(nbytes, dtype=torch.uint8, device=self.tensors[0].device)
                self.struct.packed = self.packed.data_ptr()
            _count(3)
            check(L.fvgn_mlp_pack_bf16(C.byref(self.struct), _p(self.packed), nbytes, _stream()), "mlp_pack_bf16")
            self.packed_versions = ver
        return self


def ensure_packed_bwd_many(mlps):
    """fvgn_mlp_pack_bf16x3_bwd_many: the transposed tiles of the data-gradient chain of every stale MlpRef in one launch"""
    stale = []
    for m in mlps:
        ver = _weights_key(m.tensors)
        if m.packed3b is None or m.packed3b_versions != ver:
            stale.append((m, ver))
    if not stale:
        return
    L = _lib.lib()
    for m, _ in stale:
        if m.packed3b is None:
            m.packed3b = torch.empty(L.fvgn_mlp_packed_x3_bwd_bytes(m.k_in), dtype=torch.uint8, device=m.tensors[0].device)
            m.struct.packed_x3_bwd = m.packed3b.data_ptr()
    arr = (MlpT * len(stale))(*[m.struct for m, _ in stale])
    _count((len(stale) + 47) // 48)
    check(L.fvgn_mlp_pack_bf16x3_bwd_many(arr, len(stale), _stream()), "mlp_pack_bf16x3_bwd_many")
    for m, ver in stale:
        m.packed3b_versions = ver


def ensure_packed_many(mlps, math_mode=2):
    """fvgn_mlp_pack_bf16x3_many: re-packs, in one launch, every MlpRef of `mlps` whose fp32 weights changed since its last pack (after an
    optimizer step: all of them).  The per-call MlpRef.ensure_packed(2) that follows finds them current."""
    if math_mode != 2:
        return
    key = None
    stale = []
    for m in mlps:
        ver = _weights_key(m.tensors)
        if m.packed3 is None or m.packed3_versions != ver:
            stale.append((m, ver))
    if not stale:
        return
    L = _lib.lib()
    for m, _ in stale:
        if m.packed3 is None:
            m.packed3 = torch.empty(L.fvgn_mlp_packed_x3_bytes(m.k_in), dtype=torch.uint8, device=m.tensors[0].device)
            m.struct.packed_x3 = m.packed3.data_ptr()
    arr = (MlpT * len(stale))(*[m.struct for m, _ in stale])
    _count((len(stale) + 47) // 48)
    check(L.fvgn_mlp_pack_bf16x3_many(arr, len(stale), _stream()), "mlp_pack_bf16x3_many")
    for m, ver in stale:
        m.packed3_versions = ver


# split products of the tensor-core BACKWARD kernels (dgrad / wgrad) in the fp32-equivalent mode: 6 = exact 3-way bf16 splits of both
# operands, 3 = the three leading products (relative error of a product <= 3 * 2^-17, i.e. gradients to ~1e-5; half the MMAs)
BWD_PRODUCTS = int(os.environ.get("FVGN_BWD_PRODUCTS", "3"))


class _bwd_products:
    def __init__(self, mlp):
        self.mlp = mlp

    def __enter__(self):
        self.saved = self.mlp.struct.n_products
        if self.saved != 1:
            self.mlp.struct.n_products = BWD_PRODUCTS

    def __exit__(self, *a):
        self.mlp.struct.n_products = self.saved
        return False


class _zsave:
    """sets fvgn_mlp_t.z_save = z[rows, 384] for the duration of one call (forward in bf16x3 mode writes the three layers'
    pre-activations there; the backward reads them instead of recomputing the forward)"""

    def __init__(self, mlp, z, rows, da_in=None):
        self.mlp, self.z, self.da = mlp, z, da_in
        if da_in is not None:
            _req(da_in, torch.float32, "da_in", 2)
            if tuple(da_in.shape) != (rows, 256) or not da_in.is_contiguous():
                raise ValueError(f"da_in must be a contiguous [rows, 256] fp32 tensor, got {tuple(da_in.shape)}")
        if z is not None:
            _req(z, torch.float32, "z_save", 2)
            if tuple(z.shape) != (rows, 384) or not z.is_contiguous():
                raise ValueError(f"z_save must be a contiguous [rows, 384] fp32 tensor, got {tuple(z.shape)}")

    def __enter__(self):
        self.mlp.struct.z_save = self.z.data_ptr() if self.z is not None else None
        self.mlp.struct.da_in = self.da.data_ptr() if self.da is not None else None

    def __exit__(self, *a):
        self.mlp.struct.z_save = None
        self.mlp.struct.da_in = None
        return False


def mlp_rows_fwd(mlp: MlpRef, x, math_mode=0, out=None, z_save=None):
    _req(x, torch.float32, "x", 2)
    rows = x.shape[0]
    if x.shape[1] != mlp.k_in:
        raise ValueError(f"mlp input width {x.shape[1]} != {mlp.k_in}")
    if out is None:
        out = torch.empty((rows, mlp.n_out), dtype=torch.float32, device=x.device)
    if math_mode != 0:
        mlp.ensure_packed(math_mode)
    _count(1)
    if z_save is not None and math_mode != 2:
        raise ValueError("z_save is written by the bf16x3 forward only")
    with _timed("mlp_rows"), _zsave(mlp, z_save, rows):
        check(_lib.lib().fvgn_mlp_rows_fwd(C.byref(mlp.struct), _p(x), _ld(x), rows, _p(out), _ld(out), math_mode, _stream()), "mlp_rows_fwd")
    return out


def node_aggregate_fwd(e, node_ptr, node_items, n_nodes, out=None):
    _req(e, torch.float32, "edge_attr", 2)
    F = e.shape[0]
    if e.shape[1] != HID or not e.is_contiguous():
        raise ValueError("edge latent must be contiguous [F,128]")
    _ci(node_ptr, "node_ptr"), _ci(node_items, "node_items")
    if out is None:
        out = torch.empty((n_nodes, 64), dtype=torch.float32, device=e.device)
    _count(1)
    with _timed("node_aggregate"):
        check(_lib.lib().fvgn_node_aggregate_fwd(_p(e), _p(node_ptr), _p(node_items), n_nodes, F, _p(out), _stream()), "node_aggregate_fwd")
    return out


def cell_mean(node_agg, cells_node_T):
    """per-cell mean of the three node aggregates, fp32 [C,64] (fvgn_cell_mean)"""
    _cf(node_agg, "node_agg"), _ci(cells_node_T, "cells_node_T", 3)
    Cn = cells_node_T.shape[1]
    out = torch.empty((Cn, 64), dtype=torch.float32, device=node_agg.device)
    _count(1)
    with _timed("cell_mean"):
        check(_lib.lib().fvgn_cell_mean(_p(node_agg), _p(cells_node_T), Cn, _p(out), _stream()), "cell_mean")
    return out


def cell_mean_poly(node_agg, cells_node4_T, want_fp32=True, split_out=None):
    """tri / quad hybrid meshes: per-cell mean over the k = 3 or 4 node aggregates ([4,C] table, 4th row -1 on triangles) as fp32 [C,64]
    and / or into the split fp16 shadow [C,2,64] (fvgn_cell_mean_poly)"""
    _cf(node_agg, "node_agg"), _ci(cells_node4_T, "cells_node4_T", 4)
    Cn = cells_node4_T.shape[1]
    out = torch.empty((Cn, 64), dtype=torch.float32, device=node_agg.device) if want_fp32 else None
    _count(1)
    with _timed("cell_mean"):
        check(_lib.lib().fvgn_cell_mean_poly(_p(node_agg), _p(cells_node4_T), Cn, _p(out), _p(split_out), _stream()), "cell_mean_poly")
    return out


def _cell_table_ok(cells_node_T, node_agg, n_cells, who):
    """cells_node_T [3,C] int32 (mp_times >= 2: node_agg is per node), or None (mp_times = 1: node_agg is the per-cell aggregate [C,64])"""
    if cells_node_T is None:
        if tuple(node_agg.shape) != (n_cells, 64):
            raise ValueError(f"{who}: cells_node_T=None (mp_times=1) needs a per-cell aggregate of shape [{n_cells}, 64], got {tuple(node_agg.shape)}")
        return
    _ci(cells_node_T, "cells_node_T", 3)
    if cells_node_T.dim() != 2 or cells_node_T.shape[1] != n_cells:
        raise ValueError(f"{who}: cells_node_T must have shape [3, {n_cells}], got {tuple(cells_node_T.shape)}")


def cell_block_fwd(mlp: MlpRef, x, node_agg, cells_node_T, math_mode=0, x_prime=None, x_out=None, want_out=True, z_save=None):
    _req(x, torch.float32, "x", 2)
    _req(node_agg, torch.float32, "node_agg", 2)
    Cn = x.shape[0]
    _cell_table_ok(cells_node_T, node_agg, Cn, "cell_block_fwd")
    if not (x.is_contiguous() and node_agg.is_contiguous()) or x.shape[1] != HID:
        raise ValueError("cell_block_fwd: bad shapes/strides")
    if x_prime is None:
        x_prime = torch.empty_like(x)
    if x_out is None and want_out:
        x_out = torch.empty_like(x)
    if math_mode != 0:
        mlp.ensure_packed(math_mode)
    _count(1)
    if z_save is not None and math_mode != 2:
        raise ValueError("z_save is written by the bf16x3 forward only")
    with _timed("cell_block"), _zsave(mlp, z_save, Cn):
        check(_lib.lib().fvgn_cell_block_fwd(C.byref(mlp.struct), _p(x), _p(node_agg), _p(cells_node_T), Cn, _p(x_prime), _p(x_out), math_mode,
                                             _stream()), "cell_block_fwd")
    return x_prime, x_out


def edge_block_fwd(mlp: MlpRef, x_prime, e, neighbour_cell, math_mode=0, e_prime=None, e_out=None, want_prime=False, want_out=True,
                   z_save=None):
    _req(x_prime, torch.float32, "x_prime", 2)
    _req(e, torch.float32, "edge_attr", 2)
    _req(neighbour_cell, torch.int32, "neighbour_cell", 2)
    F = e.shape[0]
    if not (x_prime.is_contiguous() and e.is_contiguous() and neighbour_cell.is_contiguous()) or e.shape[1] != HID or neighbour_cell.shape != (2, F):
        raise ValueError("edge_block_fwd: bad shapes/strides")
    if e_prime is None and want_prime:
        e_prime = torch.empty_like(e)
    if e_out is None and want_out:
        e_out = torch.empty_like(e)
    if math_mode != 0:
        mlp.ensure_packed(math_mode)
    _count(1)
    if z_save is not None and math_mode != 2:
        raise ValueError("z_save is written by the bf16x3 forward only")
    with _timed("edge_block"), _zsave(mlp, z_save, F):
        check(_lib.lib().fvgn_edge_block_fwd(C.byref(mlp.struct), _p(x_prime), _p(e), _p(neighbour_cell), F, _p(e_prime), _p(e_out), math_mode,
                                             _stream()), "edge_block_fwd")
    return e_prime, e_out


def gn_block_fwd_inplace_bf16(cell_mlp: MlpRef, edge_mlp: MlpRef, x, e, node_agg, cells_node_T, neighbour_cell, x_prime_bf16):
    """Tensor-core inference fast path of one GnBlock after the node aggregation: x += x', e += e' in place; x' only exists as the
    bf16 shadow the EdgeBlock gathers (fvgn_cell_block_fwd_xbf16 / fvgn_edge_block_fwd_xbf16)."""
    _req(x, torch.float32, "x", 2), _req(e, torch.float32, "edge_attr", 2), _req(node_agg, torch.float32, "node_agg", 2)
    _req(x_prime_bf16, torch.bfloat16, "x_prime_bf16", 2)
    _ci(neighbour_cell, "neighbour_cell", 2)
    Cn, F = x.shape[0], e.shape[0]
    _cell_table_ok(cells_node_T, node_agg, Cn, "gn_block_fwd_inplace_bf16")
    if not (x.is_contiguous() and e.is_contiguous() and node_agg.is_contiguous() and x_prime_bf16.is_contiguous()) or x.shape[1] != HID \
            or e.shape[1] != HID or tuple(x_prime_bf16.shape) != (Cn, HID) or neighbour_cell.shape[1] != F:
        raise ValueError("gn_block_fwd_inplace_bf16: bad shapes/strides")
    cell_mlp.ensure_packed(), edge_mlp.ensure_packed()
    L = _lib.lib()
    _count(2)
    with _timed("cell_block"):
        check(L.fvgn_cell_block_fwd_xbf16(C.byref(cell_mlp.struct), _p(x), _p(node_agg), _p(cells_node_T), Cn, _p(x_prime_bf16), 1, _stream()),
              "cell_block_fwd_xbf16")
    with _timed("edge_block"):
        check(L.fvgn_edge_block_fwd_xbf16(C.byref(edge_mlp.struct), _p(x_prime_bf16), _p(e), _p(neighbour_cell), F, 1, _stream()),
              "edge_block_fwd_xbf16")
    return x, e


def _split_ok(t, rows, name):
    if t.dtype != torch.float16 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (rows, 2, HID):
        raise ValueError(f"{name} must be a contiguous CUDA float16 tensor of shape [{rows}, 2, {HID}]")
    return t


def gn_block_fwd_inplace_split(cell_mlp: MlpRef, edge_mlp: MlpRef, x, e, node_agg, cells_node_T, neighbour_cell, x_prime_split,
                               cell_mean_split=None):
    """Inference fast path of one GnBlock in the fp32-equivalent mode after the node aggregation: x += x', e += e' in place; x' only
    exists as the split fp16 shadow the EdgeBlock gathers (fvgn_cell_block_fwd_split / fvgn_edge_block_fwd_split)."""
    _req(x, torch.float32, "x", 2), _req(e, torch.float32, "edge_attr", 2), _req(node_agg, torch.float32, "node_agg", 2)
    _ci(neighbour_cell, "neighbour_cell", 2)
    Cn, F = x.shape[0], e.shape[0]
    if cells_node_T is None or cells_node_T.shape[0] != 4:
        _cell_table_ok(cells_node_T, node_agg, Cn, "gn_block_fwd_inplace_split")
    if not (x.is_contiguous() and e.is_contiguous() and node_agg.is_contiguous()) or x.shape[1] != HID or e.shape[1] != HID \
            or neighbour_cell.shape[1] != F:
        raise ValueError("gn_block_fwd_inplace_split: bad shapes/strides")
    _split_ok(x_prime_split, Cn, "x_prime_split")
    if cell_mean_split is None:
        cell_mean_split = torch.empty((Cn, 2, 64), dtype=torch.float16, device=x.device)
    if cell_mean_split.dtype != torch.float16 or not cell_mean_split.is_contiguous() or tuple(cell_mean_split.shape) != (Cn, 2, 64):
        raise ValueError("cell_mean_split must be a contiguous float16 tensor of shape [C, 2, 64]")
    cell_mlp.ensure_packed(2), edge_mlp.ensure_packed(2)
    L = _lib.lib()
    _count(3)
    with _timed("cell_mean"):
        if cells_node_T is not None and cells_node_T.shape[0] == 4:  # tri / quad hybrid mesh
            check(L.fvgn_cell_mean_poly(_p(node_agg), _p(cells_node_T), Cn, None, _p(cell_mean_split), _stream()), "cell_mean_poly")
        else:
            check(L.fvgn_cell_mean_split(_p(node_agg), _p(cells_node_T), Cn, _p(cell_mean_split), _stream()), "cell_mean_split")
    with _timed("cell_block"):
        check(L.fvgn_cell_block_fwd_split(C.byref(cell_mlp.struct), _p(x), _p(cell_mean_split), Cn, _p(x_prime_split), _stream()),
              "cell_block_fwd_split")
    with _timed("edge_block"):
        check(L.fvgn_edge_block_fwd_split(C.byref(edge_mlp.struct), _p(x_prime_split), _p(e), _p(neighbour_cell), F, _stream()),
              "edge_block_fwd_split")
    return x, e


# ------------------------------------------------------------------------------------------------ connectivity
def build_connectivity(cells, mesh_pos, node_type, face_rule="B"):
    """a1 on device.  cells int32[C,3], mesh_pos f32[N,2], node_type int32[N] (CUDA) -> dict of tables in the H5 layout of
    extract_mesh_state (without the leading [1,...] axis)."""
    cells = _req(cells, torch.int32, "cells", 2).contiguous()
    mesh_pos = _req(mesh_pos, torch.float32, "mesh_pos", 2).contiguous()
    node_type = _req(node_type.reshape(-1), torch.int32, "node_type", 1).contiguous()
    Cn, N = cells.shape[0], mesh_pos.shape[0]
    dev = cells.device
    L = _lib.lib()
    if cells.shape[1] == 4:  # tri / quad hybrid mesh: [C,4] tables, 4th column -1 on triangles (fvgn_connectivity_poly_*)
        nbytes = L.fvgn_connectivity_poly_workspace_bytes(Cn)
        ws = _ws(nbytes, dev)
        cells_node = torch.empty((Cn, 4), dtype=torch.int32, device=dev)
        nf = C.c_int(0)
        check(L.fvgn_connectivity_poly_faces(_p(cells), Cn, N, _p(cells_node), C.byref(nf), _p(ws), nbytes, _stream()), "connectivity_poly_faces")
        F = nf.value
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        t = dict(cells_node=cells_node, face=i32(2, F), cells_face=i32(Cn, 4), neighbour_cell=i32(2, F), face_type=i32(F, 1), cells_type=i32(Cn, 1),
                 face_length=f32(F, 1), unit_norm_v=f32(Cn, 4, 2), centroid=f32(Cn, 2), cells_area=f32(Cn, 1), cell_factor=f32(Cn, 4))
        check(L.fvgn_connectivity_poly_tables(_p(mesh_pos), _p(node_type), _p(cells_node), Cn, N, F, {"A": 0, "B": 1}[face_rule], _p(t["face"]),
                                              _p(t["cells_face"]), _p(t["neighbour_cell"]), _p(t["face_type"]), _p(t["cells_type"]),
                                              _p(t["face_length"]), _p(t["unit_norm_v"]), _p(t["centroid"]), _p(t["cells_area"]),
                                              _p(t["cell_factor"]), _p(ws), nbytes, _stream()), "connectivity_poly_tables")
        t["mesh_pos"], t["node_type"] = mesh_pos, node_type.view(-1, 1)
        t["cell_k"] = (cells_node >= 0).sum(1).to(torch.int32)
        return t
    if cells.shape[1] != 3:
        raise ValueError(f"cells must be [C,3] (triangles) or [C,4] (tri / quad, -1 padded), got {tuple(cells.shape)}")
    nbytes = L.fvgn_connectivity_workspace_bytes(Cn)
    ws = _ws(nbytes, dev)
    cells_node = torch.empty((Cn, 3), dtype=torch.int32, device=dev)
    nf = C.c_int(0)
    check(L.fvgn_connectivity_faces(_p(cells), Cn, N, _p(cells_node), C.byref(nf), _p(ws), nbytes, _stream()), "connectivity_faces")
    F = nf.value
    i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    t = dict(cells_node=cells_node, face=i32(2, F), cells_face=i32(Cn, 3), neighbour_cell=i32(2, F), face_type=i32(F, 1), cells_type=i32(Cn, 1),
             face_length=f32(F, 1), unit_norm_v=f32(Cn, 3, 2), centroid=f32(Cn, 2), cells_area=f32(Cn, 1), cell_factor=f32(Cn, 3))
    rule = {"A": 0, "B": 1}[face_rule]
    check(L.fvgn_connectivity_tables(_p(mesh_pos), _p(node_type), _p(cells_node), Cn, N, F, rule, _p(t["face"]), _p(t["cells_face"]),
                                     _p(t["neighbour_cell"]), _p(t["face_type"]), _p(t["cells_type"]), _p(t["face_length"]),
                                     _p(t["unit_norm_v"]), _p(t["centroid"]), _p(t["cells_area"]), _p(t["cell_factor"]), _p(ws), nbytes,
                                     _stream()), "connectivity_tables")
    t["mesh_pos"], t["node_type"] = mesh_pos, node_type.view(-1, 1)
    return t


def build_csr(dst, n_dst):
    dst = _req(dst, torch.int32, "dst", 1).contiguous()
    n = dst.numel()
    L = _lib.lib()
    nbytes = L.fvgn_csr_workspace_bytes(n)
    ws = _ws(nbytes, dst.device)
    ptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dst.device)
    items = torch.empty(n, dtype=torch.int32, device=dst.device)
    _count(4)
    check(L.fvgn_build_csr(_p(dst), n, n_dst, _p(ptr), _p(items), _p(ws), nbytes, _stream()), "build_csr")
    return ptr, items


# ------------------------------------------------------------------------------------------------ pre / post
def interp_nodes(node_field, col0, width, cells_node_T=None, cell_factor=None, face_node=None, want_cell=True, want_face=True):
    _req(node_field, torch.float32, "node_field", 2)
    if want_cell and cells_node_T is not None and cells_node_T.dim() == 2 and cells_node_T.shape[0] == 4:
        # tri / quad hybrid mesh: cells through fvgn_interp_cells_poly, faces through the common kernel
        _ci(cells_node_T, "cells_node4_T", 4), _cf(cell_factor, "cell_factor4")
        Cn = cells_node_T.shape[1]
        if tuple(cell_factor.shape) != (Cn, 4):
            raise ValueError(f"cell_factor must be [{Cn}, 4] for a [4,C] node table, got {tuple(cell_factor.shape)}")
        on_cell = torch.empty((Cn, width), dtype=torch.float32, device=node_field.device)
        _count(1)
        check(_lib.lib().fvgn_interp_cells_poly(_p(node_field), _ld(node_field), col0, width, _p(cells_node_T), _p(cell_factor), Cn, _p(on_cell),
                                                _stream()), "interp_cells_poly")
        on_face = interp_nodes(node_field, col0, width, face_node=face_node, want_cell=False, want_face=True)[1] if want_face else None
        return on_cell, on_face
    if want_cell:
        _ci(cells_node_T, "cells_node_T", 3), _cf(cell_factor, "cell_factor")
    if want_face:
        _ci(face_node, "face_node", 2)
    dev = node_field.device
    Cn = cells_node_T.shape[1] if want_cell else 0
    F = face_node.shape[1] if want_face else 0
    on_cell = torch.empty((Cn, width), dtype=torch.float32, device=dev) if want_cell else None
    on_face = torch.empty((F, width), dtype=torch.float32, device=dev) if want_face else None
    _count(1)
    check(_lib.lib().fvgn_interp_nodes(_p(node_field), _ld(node_field), col0, width, _p(cells_node_T), _p(cell_factor), _p(face_node), Cn, F,
                                       _p(on_cell), _p(on_face), _stream()), "interp_nodes")
    return on_cell, on_face


def assemble_edge_attr(cell_uv, centroid, neighbour_cell, face_type_len, bc_uv, flip_keep=None, want_mask=True):
    _ci(neighbour_cell, "neighbour_cell", 2), _cf(centroid, "centroid"), _cf(face_type_len, "face_type_len")
    _req(cell_uv, torch.float32, "cell_uv", 2), _req(bc_uv, torch.float32, "bc_uv", 2)
    F = neighbour_cell.shape[1]
    dev = cell_uv.device
    ea = torch.empty((F, 5), dtype=torch.float32, device=dev)
    mask = torch.empty(F, dtype=torch.uint8, device=dev) if want_mask else None
    if flip_keep is not None:
        flip_keep = flip_keep.to(torch.uint8).contiguous()
    _count(1)
    check(_lib.lib().fvgn_assemble_edge_attr(_p(cell_uv), _ld(cell_uv), _p(centroid), _p(neighbour_cell), _p(face_type_len), _p(bc_uv), _ld(bc_uv),
                                             _p(flip_keep), F, _p(ea), _p(mask), _stream()), "assemble_edge_attr")
    return ea, (mask.bool() if want_mask else None)


def create_next_graph_(x, edge_attr, next_uv, neighbour_cell, face_type_len, bc_uv):
    """in-place on x[C,4] and edge_attr[F,5]"""
    _ci(neighbour_cell, "neighbour_cell", 2), _cf(face_type_len, "face_type_len"), _req(bc_uv, torch.float32, "bc_uv", 2)
    Cn, F = x.shape[0], edge_attr.shape[0]
    if not (x.is_contiguous() and edge_attr.is_contiguous() and next_uv.is_contiguous()):
        raise ValueError("create_next_graph_: contiguous tensors required")
    _count(1)
    check(_lib.lib().fvgn_create_next_graph(_p(next_uv), _p(neighbour_cell), _p(face_type_len), _p(bc_uv), _ld(bc_uv), Cn, F, _p(x), _p(edge_attr),
                                            _stream()), "create_next_graph")


# ------------------------------------------------------------------------------------------------ normalizer
def normalizer_accumulate(x, acc_count, num_acc, acc_sum, acc_sq, max_accumulations, type_col=None, one_hot=0):
    _req(x, torch.float32, "x", 2)
    width = x.shape[1] + one_hot
    L = _lib.lib()
    ws = torch.empty(L.fvgn_norm_workspace_floats(width) // 2 + 8, dtype=torch.float64, device=x.device)
    _count(2)
    check(L.fvgn_normalizer_accumulate(_p(x), _ld(x), x.shape[1], _p(type_col), _ld(type_col) if type_col is not None else 0, one_hot, x.shape[0],
                                       float(max_accumulations), _p(acc_count), _p(num_acc), _p(acc_sum), _p(acc_sq), _p(ws), _stream()),
          "normalizer_accumulate")


def normalizer_apply(x, acc_count, acc_sum, acc_sq, inverse=False, type_col=None, one_hot=0, addend=None, width=None, out=None):
    _req(x, torch.float32, "x", 2)
    wr = x.shape[1] if width is None else width
    w = wr + one_hot
    if out is None:
        out = torch.empty((x.shape[0], w), dtype=torch.float32, device=x.device)
    _count(1)
    check(_lib.lib().fvgn_normalizer_apply(_p(x), _ld(x), wr, _p(type_col), _ld(type_col) if type_col is not None else 0, one_hot, x.shape[0],
                                           _p(acc_count), _p(acc_sum), _p(acc_sq), 1 if inverse else 0, _p(addend),
                                           _ld(addend) if addend is not None else 0, _p(out), _ld(out), _stream()), "normalizer_apply")
    return out


# ------------------------------------------------------------------------------------------------ FV head
def face_area_sums(face_type_len, cell_area, neighbour_cell, dt):
    """local raw feature [F,1] and float64 [sum, sum of squares] (for the cross-rank BatchNorm statistics)"""
    _ci(neighbour_cell, "neighbour_cell", 2), _cf(face_type_len, "face_type_len"), _cf(cell_area, "cell_area")
    F = neighbour_cell.shape[1]
    L = _lib.lib()
    raw = torch.empty((F, 1), dtype=torch.float32, device=cell_area.device)
    sums = torch.empty(2, dtype=torch.float64, device=cell_area.device)
    ws = torch.empty(L.fvgn_norm_workspace_floats(1) // 2 + 8, dtype=torch.float64, device=cell_area.device)
    _count(2)
    check(L.fvgn_face_area_sums(_p(face_type_len), _p(cell_area), _p(neighbour_cell), F, float(dt), _p(raw), _p(sums), _p(ws), _stream()), "face_area_sums")
    return raw, sums


def face_area_bn_fwd(face_type_len, cell_area, neighbour_cell, dt, training, bn_weight, bn_bias, running_mean, running_var, want_raw=False,
                     stats=None):
    _ci(neighbour_cell, "neighbour_cell", 2), _cf(face_type_len, "face_type_len"), _cf(cell_area, "cell_area")
    F = neighbour_cell.shape[1]
    dev = cell_area.device
    L = _lib.lib()
    out = torch.empty((F, 1), dtype=torch.float32, device=dev)
    raw = torch.empty((F, 1), dtype=torch.float32, device=dev) if want_raw else None
    mode = 1 if training else 0
    if stats is not None:  # caller-provided [mean, invstd] (cross-rank batch statistics, parallel.global_bn_stats)
        stats = _cf(stats, "stats")
        mode = 2
    else:
        stats = torch.empty(2, dtype=torch.float32, device=dev)
    ws = torch.empty(L.fvgn_norm_workspace_floats(1) // 2 + 8, dtype=torch.float64, device=dev)
    _count(3 if mode != 2 else 2)
    check(L.fvgn_face_area_bn_fwd(_p(face_type_len), _p(cell_area), _p(neighbour_cell), F, float(dt), mode, _p(bn_weight),
                                  _p(bn_bias), _p(running_mean), _p(running_var), _p(out), _p(raw), _p(stats), _p(ws), _stream()), "face_area_bn_fwd")
    return out, raw, stats


def integrator_fwd(decoded, unv, face_area, cells_face_T, rho, rhs_coef=1.0, want_continuity=False):
    poly = cells_face_T.dim() == 2 and cells_face_T.shape[0] == 4  # tri / quad hybrid mesh: [4,C] face table, unv [C,4,2]
    _ci(cells_face_T, "cells_face_T", 4 if poly else 3)
    Cn = cells_face_T.shape[1]
    if poly:
        if tuple(unv.shape) != (Cn, 4, 2) or not (decoded.is_contiguous() and unv.is_contiguous() and face_area.is_contiguous()):
            raise ValueError(f"integrator_fwd: a [4,C] face table needs contiguous unv [{Cn},4,2], got {tuple(unv.shape)}")
        du = torch.empty((Cn, 2), dtype=torch.float32, device=decoded.device)
        cont = torch.empty((Cn, 1), dtype=torch.float32, device=decoded.device) if want_continuity else None
        _count(1)
        check(_lib.lib().fvgn_integrator_fwd_poly(_p(decoded), _p(unv), _p(face_area), _p(cells_face_T), Cn, float(rho), float(rhs_coef), _p(du),
                                                  _p(cont), _stream()), "integrator_fwd_poly")
        return du, cont
    dev = decoded.device
    du = torch.empty((Cn, 2), dtype=torch.float32, device=dev)
    cont = torch.empty((Cn, 1), dtype=torch.float32, device=dev) if want_continuity else None
    if not (decoded.is_contiguous() and unv.is_contiguous() and face_area.is_contiguous()):
        raise ValueError("integrator_fwd: contiguous tensors required")
    _count(1)
    check(_lib.lib().fvgn_integrator_fwd(_p(decoded), _p(unv), _p(face_area), _p(cells_face_T), Cn, float(rho), float(rhs_coef), _p(du), _p(cont),
                                         _stream()), "integrator_fwd")
    return du, cont


LOSS_MODES = {"direct_mean_loss": 0, "global_mean_sum": 1, "global_mean": 2}


LOSS_CHUNKS = 32  # partial sums per segment in the loss workspace (csrc/prepost.cu)


def loss_fwd(pred_du, target_du, pred_edge, target_face, mask_interior, cell_batch, face_batch, n_graphs, mode, w_mom, w_uv, w_p,
             reduce_partials=None):
    """compute_FVM_loss forward.  `reduce_partials` (data-parallel direct_mean_loss): callable that all-reduces (SUM, in place) the
    float64 [n_seg, 6] column sums {sq momentum error, n_cells, sq interior uv error, n_interior, sq p error, n_faces} over the ranks
    BEFORE the log is taken (utils/loss_compute.py:92-113 takes one log over the batch means)."""
    dev = pred_du.device
    L = _lib.lib()
    out = torch.empty(5, dtype=torch.float32, device=dev)
    ws = torch.empty(L.fvgn_loss_workspace_floats(n_graphs) // 2 + 8, dtype=torch.float64, device=dev)
    m8 = mask_interior.to(torch.uint8).contiguous()
    _count(2)
    if reduce_partials is None:
        check(L.fvgn_loss_fwd(_p(pred_du), _p(target_du), _p(pred_edge), _p(target_face), _p(m8), _p(cell_batch), _p(face_batch), pred_du.shape[0],
                              pred_edge.shape[0], n_graphs, mode, float(w_mom), float(w_uv), float(w_p), _p(out), _p(ws), _stream()), "loss_fwd")
        return out, ws
    check(L.fvgn_loss_partials(_p(pred_du), _p(target_du), _p(pred_edge), _p(target_face), _p(m8), _p(cell_batch), _p(face_batch), pred_du.shape[0],
                               pred_edge.shape[0], n_graphs, mode, _p(ws), _stream()), "loss_partials")
    n_seg = 1 if mode == 0 else n_graphs
    part = ws[:n_seg * LOSS_CHUNKS * 6].view(n_seg, LOSS_CHUNKS, 6)
    tot = part.sum(dim=1)
    tot[:, 1], tot[:, 5] = part[:, 0, 1], part[:, 0, 5]  # the counts are per segment, repeated in every chunk
    reduce_partials(tot)
    part.zero_()
    part[:, 0, :] = tot
    check(L.fvgn_loss_finish(n_graphs, mode, float(w_mom), float(w_uv), float(w_p), _p(out), _p(ws), _stream()), "loss_finish")
    return out, ws


STATUS_NONFINITE_X3 = 1


def status(clear=True):
    """device status word (fvgn_status_read; synchronises the device).  Bit STATUS_NONFINITE_X3: a row of a fp32-equivalent ('bf16x3')
    forward tile came out non-finite since the last clear (fp16 operand terms overflow for |activation| >= ~8190)."""
    v = C.c_uint(0)
    check(_lib.lib().fvgn_status_read(C.byref(v), 1 if clear else 0), "status_read")
    return int(v.value)


def raise_if_nonfinite(what="forward"):
    if status(clear=True) & STATUS_NONFINITE_X3:
        raise FloatingPointError(
            f"fvgn_b200: non-finite rows in the fp32-equivalent ('bf16x3') tensor-core {what}: an activation reached the fp16 operand range "
            "(|a| >= ~8190) or the inputs held inf/NaN.  The reference's fp32 path stays finite on large activations: re-run with "
            "model.set_math_mode('fp32').")


# ------------------------------------------------------------------------------------------------ backward (training)
def _mlp_grad_views(mlp: MlpRef, g):
    """split the flat gradient buffer of fvgn_mlp_*_bwd into tensors shaped like the parameters"""
    k = mlp.k_in
    o = 0
    dW1 = g[o:o + 128 * k].view(128, k); o += 128 * k
    dW2 = g[o:o + 16384].view(128, 128); o += 16384
    dW3 = g[o:o + 16384].view(128, 128)[:mlp.n_out]; o += 16384
    db1, db2, db3, dg, dbeta = (g[o + 128 * i:o + 128 * (i + 1)] for i in range(5))
    out = [dW1, db1, dW2, db2, dW3, db3[:mlp.n_out]]
    if len(mlp.tensors) == 8:
        out += [dg, dbeta]
    return out


def _bwd_buffers(mlp: MlpRef, device):
    L = _lib.lib()
    grads = torch.empty(L.fvgn_mlp_grad_floats(mlp.k_in), dtype=torch.float32, device=device)
    ws = torch.empty(L.fvgn_mlp_bwd_workspace_floats(mlp.k_in), dtype=torch.float32, device=device)
    return grads, ws


def ln_bwd(d_out, z_saved, ln_w):
    """LayerNorm backward of d_out with the statistics of the saved z3: (dz[rows,128] for fvgn_mlp_dgrad_x3, the per-row-range partial
    column sums of dz | d_out * xhat | d_out that fvgn_mlp_wgrad_x3 finishes into db3, dLN_w, dLN_b)"""
    _req(d_out, torch.float32, "d_out", 2), _cf(z_saved, "z_saved"), _cf(ln_w, "ln_w")
    rows = d_out.shape[0]
    dz = torch.empty((rows, 128), dtype=torch.float32, device=d_out.device)
    gxhat = torch.empty(_lib.lib().fvgn_ln_bwd_sums_floats(rows), dtype=torch.float32, device=d_out.device)
    _count(1)
    check(_lib.lib().fvgn_ln_bwd(_p(d_out), _ld(d_out), _p(z_saved), _p(ln_w), rows, _p(dz), _p(gxhat), _stream()), "ln_bwd")
    return dz, gxhat


def mlp_wgrad_x3_buffers(mlp: MlpRef, device):
    """(grads, workspace) of one fvgn_mlp_wgrad_x3 call — allocated by the caller when the call itself is enqueued on another stream
    (the caching allocator's pools are per stream; the workspace must then stay alive until that stream has been waited for)"""
    L = _lib.lib()
    return (torch.empty(L.fvgn_mlp_grad_floats(mlp.k_in), dtype=torch.float32, device=device),
            torch.empty(L.fvgn_mlp_wgrad_x3_workspace_floats(mlp.k_in), dtype=torch.float32, device=device))


def mlp_wgrad_x3(mlp: MlpRef, mode, d_out, dz, gxhat, z_saved, da, x=None, src0=None, src1=None, idx=None, buffers=None):
    """all parameter gradients of one MLP on the tensor cores (MN-major P^T . Q products, deterministic).  mode 0: x = layer input
    [rows,k_in]; 1 (CellBlock): src0 = x, src1 = node_agg, idx = cells_node_T; 2 (EdgeBlock): src0 = x', src1 = e, idx = neighbour_cell"""
    _req(d_out, torch.float32, "d_out", 2)
    rows = d_out.shape[0]
    L = _lib.lib()
    grads, ws = buffers if buffers is not None else mlp_wgrad_x3_buffers(mlp, d_out.device)
    if mode == 0:
        _req(x, torch.float32, "x", 2)
    elif mode == 1:
        _cf(src0, "src0"), _cf(src1, "src1"), _cell_table_ok(idx, src1, rows, "mlp_wgrad_x3")
    else:
        _cf(src0, "src0"), _cf(src1, "src1"), _ci(idx, "idx")
    _count(3)  # wgrad, range reduction, column-sum finish
    with _zsave(mlp, z_saved, rows, da), _bwd_products(mlp):
        check(L.fvgn_mlp_wgrad_x3(C.byref(mlp.struct), mode, _p(x), _ld(x) if x is not None else 0, _p(src0), _p(src1), _p(idx), rows, _p(d_out),
                                  _ld(d_out), _p(dz), _p(gxhat), _p(grads), _p(ws), _stream()), "mlp_wgrad_x3")
    return _mlp_grad_views(mlp, grads)


def mlp_dgrad_x3(mlp: MlpRef, dz, z_saved, want_dA=True, dA_base=None):
    """tensor-core data-gradient chain: returns (da[rows,256] = da2 | da1, dA[rows,k_in] or None)"""
    _req(dz, torch.float32, "dz", 2)
    rows = dz.shape[0]
    mlp.ensure_packed_bwd()
    da = torch.empty((rows, 256), dtype=torch.float32, device=dz.device)
    dA = torch.empty((rows, mlp.k_in), dtype=torch.float32, device=dz.device) if want_dA else None
    _count(1)
    with _zsave(mlp, z_saved, rows), _bwd_products(mlp):
        check(_lib.lib().fvgn_mlp_dgrad_x3(C.byref(mlp.struct), _p(dz), _ld(dz), rows, _p(da), _p(dA), mlp.k_in if want_dA else 0, _p(dA_base),
                                           _ld(dA_base) if dA_base is not None else 0, _stream()), "mlp_dgrad_x3")
    return da, dA


def mlp_rows_bwd(mlp: MlpRef, x, d_out, want_d_in=False, z_saved=None, da_in=None):
    _req(x, torch.float32, "x", 2), _req(d_out, torch.float32, "d_out", 2)
    grads, ws = _bwd_buffers(mlp, x.device)
    d_in = torch.empty((x.shape[0], mlp.k_in), dtype=torch.float32, device=x.device) if want_d_in else None
    _count(2)
    with _zsave(mlp, z_saved, x.shape[0], da_in):
        check(_lib.lib().fvgn_mlp_rows_bwd(C.byref(mlp.struct), _p(x), _ld(x), x.shape[0], _p(d_out), _ld(d_out), _p(d_in),
                                           _ld(d_in) if d_in is not None else 0, _p(grads), _p(ws), _stream()), "mlp_rows_bwd")
    return _mlp_grad_views(mlp, grads), d_in


def cell_block_bwd(mlp: MlpRef, x, node_agg, cells_node_T, d_x_prime, d_x_base=None, z_saved=None, da_in=None):
    _cf(x, "x"), _cf(node_agg, "node_agg"), _cf(d_x_prime, "d_x_prime")
    Cn = x.shape[0]
    _cell_table_ok(cells_node_T, node_agg, Cn, "cell_block_bwd")
    grads, ws = _bwd_buffers(mlp, x.device)
    d_in = torch.empty((Cn, 192), dtype=torch.float32, device=x.device) if da_in is None else None
    _count(2)
    with _zsave(mlp, z_saved, Cn, da_in):
        check(_lib.lib().fvgn_cell_block_bwd(C.byref(mlp.struct), _p(x), _p(node_agg), _p(cells_node_T), Cn, _p(d_x_prime), _p(d_x_base),
                                             _ld(d_x_base) if d_x_base is not None else 0, _p(d_in), _p(grads), _p(ws), _stream()),
              "cell_block_bwd")
    return _mlp_grad_views(mlp, grads), d_in


def edge_block_bwd(mlp: MlpRef, x_prime, e, neighbour_cell, d_e_prime, z_saved=None, da_in=None):
    _cf(x_prime, "x_prime"), _cf(e, "e"), _ci(neighbour_cell, "neighbour_cell", 2), _cf(d_e_prime, "d_e_prime")
    F = e.shape[0]
    grads, ws = _bwd_buffers(mlp, e.device)
    d_in = torch.empty((F, 384), dtype=torch.float32, device=e.device) if da_in is None else None
    _count(2)
    with _zsave(mlp, z_saved, F, da_in):
        check(_lib.lib().fvgn_edge_block_bwd(C.byref(mlp.struct), _p(x_prime), _p(e), _p(neighbour_cell), F, _p(d_e_prime), _p(d_in), _p(grads),
                                             _p(ws), _stream()), "edge_block_bwd")
    return _mlp_grad_views(mlp, grads), d_in


def segment_sum_rows(src, n_src, col0, part_stride, width, ptr, items, n_dst, base=None, scale=1.0):
    _req(src, torch.float32, "src", 2), _ci(ptr, "ptr"), _ci(items, "items")
    out = torch.empty((n_dst, width), dtype=torch.float32, device=src.device)
    _count(1)
    check(_lib.lib().fvgn_segment_sum_rows(_p(src), _ld(src), n_src, col0, part_stride, width, _p(ptr), _p(items), n_dst, _p(base),
                                           _ld(base) if base is not None else 0, float(scale), _p(out), width, _stream()), "segment_sum_rows")
    return out


def edge_grad_combine(g_e_out, dA_edge, g_node, face_node):
    _cf(g_e_out, "g_e_out"), _cf(dA_edge, "dA_edge"), _cf(g_node, "g_node"), _ci(face_node, "face_node", 2)
    out = torch.empty_like(g_e_out)
    _count(1)
    check(_lib.lib().fvgn_edge_grad_combine(_p(g_e_out), _p(dA_edge), _p(g_node), _p(face_node), g_e_out.shape[0], _p(out), _stream()),
          "edge_grad_combine")
    return out


def loss_bwd(pred_du, target_du, pred_edge, target_face, mask8, cell_batch, face_batch, n_graphs, mode, w_mom, w_uv, w_p, g_loss, fwd_ws):
    dev = pred_du.device
    g_du = torch.empty_like(pred_du)
    g_edge = torch.empty((pred_edge.shape[0], 5), dtype=torch.float32, device=dev)
    coef = torch.empty(3 * max(n_graphs, 1), dtype=torch.float32, device=dev)
    _count(2)
    check(_lib.lib().fvgn_loss_bwd(_p(pred_du), _p(target_du), _p(pred_edge), _p(target_face), _p(mask8), _p(cell_batch), _p(face_batch),
                                   pred_du.shape[0], pred_edge.shape[0], n_graphs, mode, float(w_mom), float(w_uv), float(w_p), _p(g_loss),
                                   _p(fwd_ws), _p(g_du), _p(g_edge), _p(coef), _stream()), "loss_bwd")
    return g_du, g_edge


def integrator_bwd_(g_decoded, decoded, unv, face_area, g_delta_u, face_ptr, face_items, n_cells, rho, rhs_coef=1.0):
    """g_decoded[F,5] is updated in place; returns g_face_area[F,1]"""
    _cf(g_decoded, "g_decoded"), _cf(decoded, "decoded"), _cf(unv, "unv"), _cf(face_area, "face_area"), _cf(g_delta_u, "g_delta_u")
    F = decoded.shape[0]
    g_fa = torch.empty((F, 1), dtype=torch.float32, device=decoded.device)
    _count(1)
    L = _lib.lib()
    poly = unv.dim() == 3 and unv.shape[1] == 4  # tri / quad mesh: unv [C,4,2], CSR over the [4,C] face table
    fn = L.fvgn_integrator_bwd_poly if poly else L.fvgn_integrator_bwd
    check(fn(_p(decoded), _p(unv), _p(face_area), _p(g_delta_u), _p(face_ptr), _p(face_items), n_cells, F, float(rho), float(rhs_coef),
             _p(g_decoded), _p(g_fa), _stream()), "integrator_bwd")
    return g_fa


def face_area_bn_bwd(g_face_area, raw, stats):
    L = _lib.lib()
    out = torch.empty(2, dtype=torch.float32, device=raw.device)
    ws = torch.empty(L.fvgn_norm_workspace_floats(1) // 2 + 8, dtype=torch.float64, device=raw.device)
    _count(2)
    check(L.fvgn_face_area_bn_bwd(_p(g_face_area), _p(raw), _p(stats), raw.shape[0], _p(out), _p(ws), _stream()), "face_area_bn_bwd")
    return out


# ------------------------------------------------------------------------------------------------ C-side executor (training layer loops)
def processor_train_fwd(cell_mlps, edge_mlps, mp_times, tables, X, E, XP, NA, ZC, ZE, node_scratch):
    """fvgn_processor_train_fwd: all GnBlocks of a training forward in one call (layer-major arrays, see include/fvgn_b200.h)"""
    L = len(cell_mlps)
    ca = (MlpT * L)(*[m.struct for m in cell_mlps])
    ea = (MlpT * L)(*[m.struct for m in edge_mlps])
    _count((4 if mp_times >= 2 else 3) * L)
    with _timed("processor_fwd"):
        check(_lib.lib().fvgn_processor_train_fwd(ca, ea, L, int(mp_times), C.byref(tables), _p(X), _p(E), _p(XP), _p(NA), _p(ZC), _p(ZE),
                                                  _p(node_scratch), _stream()), "processor_train_fwd")


def processor_train_bwd(cell_mlps, edge_mlps, lo, hi, mp_times, tables, X, E, XP, NA, ZC, ZE, g_e_in, g_e_out, g_x_in, g_x_out, grad_ptrs,
                        scratch, side_stream):
    """fvgn_processor_train_bwd: the backward of the GnBlocks hi-1 ... lo in one call; grad_ptrs = ctypes array [2L] of device pointers"""
    L = len(cell_mlps)
    ca = (MlpT * L)(*[m.struct for m in cell_mlps])
    ea = (MlpT * L)(*[m.struct for m in edge_mlps])
    _count(13 * (hi - lo) + 1)
    check(_lib.lib().fvgn_processor_train_bwd(ca, ea, int(lo), int(hi), int(mp_times), BWD_PRODUCTS, C.byref(tables), _p(X), _p(E), _p(XP), _p(NA),
                                              _p(ZC), _p(ZE), _p(g_e_in), _p(g_e_out), _p(g_x_in), _p(g_x_out), grad_ptrs, _p(scratch),
                                              scratch.numel(), C.c_void_p(side_stream.cuda_stream) if side_stream is not None else None,
                                              _stream()), "processor_train_bwd")
