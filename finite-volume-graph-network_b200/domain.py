"""Single-mesh domain decomposition: ONE mesh partitioned across the GPUs of a box, the rollout step of
`FVGN.forward` (eval) + `create_next_graph` (FVGN_model/FVGN.py:249-326, dataset/Load_mesh.py:1076-1115) run on every rank's part with two
exchanges per GnBlock — strong scaling for meshes beyond one GPU's memory (the reference has no multi-GPU code at all: SURVEY §2.2).

Partition (host, numpy, once per mesh): cells in `n_parts` stripes of equal count along x (centroid order).  Rank r holds
  * its OWNED cells (rows [0, C_own)) and the GHOST cells across its cut faces (rows [C_own, C_own + G)) — only their x' / uv are ever used;
  * all faces of its owned cells (a cut face lives on both ranks: both compute its latent from the same inputs, bit for bit);
  * all nodes of its owned cells.
What a GnBlock needs from other ranks (GN_blocks.py:78-145, 16-54):
  1. the node aggregate of an INTERFACE node sums the halves of faces that live on several ranks: every face is aggregated by exactly one
     rank (the owner of its first neighbour cell), and the partial sums of the interface nodes are exchanged and added in ascending rank
     order (the same order on every rank: deterministic; vs the single-GPU CSR order the sum is re-associated, i.e. equal to ~1e-7 relative);
  2. the EdgeBlock gathers x' of both neighbour cells: the ghost rows are received from their owners after the CellBlock.
Per rollout step additionally the next velocities of the ghost cells (2 floats per cell) for `create_next_graph`.
Exchanges are `torch.distributed` point-to-point batches (NCCL over NVLink on a B200 box); `LocalComm` runs all parts in one process with
direct copies (the CPU test of this orchestration: tests/test_domain_cpu.py, with the oracle's torch arithmetic as the backend).
"""
from __future__ import annotations

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ partition (host)
class Part:
    """tables of one rank's part, local numbering (numpy)"""

    def __init__(self, rank, **kw):
        self.rank = rank
        self.__dict__.update(kw)


def decompose(tab: dict, n_parts: int):
    """tab: the connectivity / geometry tables of ONE mesh (ops.build_connectivity / the H5 group, numpy or CPU tensors; [C,3] triangles).
    Returns (parts, part_of_cell).  All exchange lists are in ascending GLOBAL id order, so the two sides of every exchange agree."""
    g = {k: (v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in tab.items()}
    cn, cf, nc, face = g["cells_node"].astype(np.int64), g["cells_face"].astype(np.int64), g["neighbour_cell"].astype(np.int64), g["face"].astype(np.int64)
    C, F, N = cn.shape[0], face.shape[1], g["mesh_pos"].shape[0]
    if cn.shape[1] != 3:
        raise NotImplementedError("domain decomposition takes triangle meshes ([C,3] tables)")
    order = np.lexsort((g["centroid"][:, 1], g["centroid"][:, 0]))
    part_of_cell = np.empty(C, dtype=np.int64)
    for r, chunk in enumerate(np.array_split(order, n_parts)):
        part_of_cell[chunk] = r
    face_owner = part_of_cell[nc[0]]  # the rank that aggregates this face's halves onto the nodes
    parts = []
    for r in range(n_parts):
        own = np.where(part_of_cell == r)[0]
        lf = np.unique(cf[own].reshape(-1))
        nb = nc[:, lf]
        ghost = np.setdiff1d(np.unique(nb.reshape(-1)), own)
        ln = np.unique(cn[own].reshape(-1))
        cell_l = np.full(C, -1, dtype=np.int64)
        cell_l[own] = np.arange(own.size)
        cell_l[ghost] = own.size + np.arange(ghost.size)
        face_l = np.full(F, -1, dtype=np.int64)
        face_l[lf] = np.arange(lf.size)
        node_l = np.full(N, -1, dtype=np.int64)
        node_l[ln] = np.arange(ln.size)
        fn_l = node_l[face[:, lf]]
        assert (fn_l >= 0).all()
        # node <- (face, half) destinations: halves of faces this rank does not own go to a dummy row (index N_local)
        agg_dst = np.where(face_owner[lf][None, :] == r, fn_l, ln.size).reshape(-1)
        parts.append(Part(
            r, n_parts=n_parts, own=own, ghost=ghost, faces=lf, nodes=ln, C_own=own.size, G=ghost.size, F=lf.size, N=ln.size,
            cells_node_T=np.ascontiguousarray(node_l[cn[own]].T), cells_face_T=np.ascontiguousarray(face_l[cf[own]].T),
            face_node=fn_l, neighbour_cell=cell_l[nb], agg_dst=agg_dst, ghost_owner=part_of_cell[ghost],
            face_type=g["face_type"].reshape(-1)[lf], face_length=g["face_length"].reshape(-1)[lf], cells_type=g["cells_type"].reshape(-1)[np.concatenate([own, ghost])],
            unv=g["unit_norm_v"][own], cells_area=g["cells_area"].reshape(-1)[np.concatenate([own, ghost])],
            centroid=g["centroid"][np.concatenate([own, ghost])], cell_factor=g["cell_factor"][own]))
    # exchange plans
    node_sets = [set(p.nodes.tolist()) for p in parts]
    for p in parts:
        p.node_share, p.halo_send, p.halo_recv = {}, {}, {}
        node_l = {int(n): i for i, n in enumerate(p.nodes)}
        own_l = {int(c): i for i, c in enumerate(p.own)}
        for q in parts:
            if q.rank == p.rank:
                continue
            shared = sorted(node_sets[p.rank] & node_sets[q.rank])
            if shared:
                p.node_share[q.rank] = np.array([node_l[n] for n in shared], dtype=np.int64)
            need = q.ghost[q.ghost_owner == p.rank]  # q's ghost cells that I own (ascending global id): I send their rows
            if need.size:
                p.halo_send[q.rank] = np.array([own_l[int(c)] for c in need], dtype=np.int64)
            mine = np.where(p.ghost_owner == q.rank)[0]
            if mine.size:
                p.halo_recv[q.rank] = p.C_own + mine
    return parts, part_of_cell


# ------------------------------------------------------------------------------------------------ communicators
class LocalComm:
    """all parts in ONE process: `exchange` is called once per part per phase; the copies happen when the last part has posted"""

    def __init__(self, n):
        self.n, self.posted = n, {}

    def post(self, rank, sends):
        self.posted[rank] = sends

    def collect(self, rank):
        return {q: self.posted[q][rank] for q in self.posted if rank in self.posted[q]}


class DistComm:
    """torch.distributed point-to-point batches (NCCL on the GPU box)"""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group)

    def exchange(self, sends: dict, recv_shapes: dict, like):
        dist = self.dist
        recvs = {q: torch.empty(shape, dtype=like.dtype, device=like.device) for q, shape in recv_shapes.items()}
        ops = [dist.P2POp(dist.isend, t.contiguous(), q, self.group) for q, t in sorted(sends.items())]
        ops += [dist.P2POp(dist.irecv, recvs[q], q, self.group) for q in sorted(recvs)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return recvs


# ------------------------------------------------------------------------------------------------ exchanges over peer memory
EXCHANGE_KINDS = (("node", 256), ("halo", 512), ("uv", 8))  # bytes per row: 64 floats | one x' row (fp32, or the fp16 hi / lo shadow) | next_uv


def _align(n, a=256):
    return (int(n) + a - 1) // a * a


def exchange_layout(parts, r):
    """byte layout of rank r's receive buffer — flag block (one 8-byte sequence flag per kind and source rank), then per kind the regions of
    its peers in ascending rank order, twice (sequence parity).  Every rank computes every rank's layout from the same `parts`."""
    p = parts[r]
    n_parts = len(parts)
    off = _align(len(EXCHANGE_KINDS) * n_parts * 8, 4096)
    lay = {}
    for kind, rb in EXCHANGE_KINDS:
        if kind == "node":
            cnt = {q: len(ix) for q, ix in p.node_share.items()}
        else:
            cnt = {q: len(p.halo_recv.get(q, ())) for q in set(p.halo_recv) | set(p.halo_send)}
        offs, o = {}, 0
        for q in sorted(cnt):
            offs[q] = o
            o += _align(cnt[q] * rb)
        lay[kind] = dict(base=off, stride=o, off=offs, n=cnt)
        off += 2 * o
    lay["total"] = off
    return lay


class PeerExchange:
    """the three exchange plans (fvgn_exchange_plan_t) of one part: csrc/halo.cu stores rows straight into the neighbours' receive buffers and
    raises a sequence flag there; the neighbour's wait kernel consumes them.  `bufs[q]` = base address of rank q's receive buffer as THIS
    device sees it (peer mapping of a symmetric allocation, or another tensor on the same device)."""

    def __init__(self, parts, run: "PartRollout", bufs, device):
        from ._lib import EXCHANGE_MAX_PEERS, ExchangePlanT

        p = run.p
        r, n_parts = p.rank, len(parts)
        t32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int32).to(device)
        self.keep, self.plans = [], {}
        self.state = torch.zeros(16, dtype=torch.int64, device=device)      # per kind: seq[2]
        self.counters = torch.zeros(16, dtype=torch.int32, device=device)   # per kind: counters[2]; [15] = status word
        lays = {q: exchange_layout(parts, q) for q in range(n_parts)}
        for k, (kind, rb) in enumerate(EXCHANGE_KINDS):
            mine = lays[r][kind]
            peers = sorted(mine["n"])
            if len(peers) > EXCHANGE_MAX_PEERS:
                raise NotImplementedError(f"a part with {len(peers)} neighbours (the exchange plans hold {EXCHANGE_MAX_PEERS})")
            pl = ExchangePlanT()
            pl.n_peers, pl.row_bytes = len(peers), rb
            pl.local_parity_stride = mine["stride"]
            pl.seq = self.state.data_ptr() + 16 * k
            pl.counters = self.counters.data_ptr() + 8 * k
            pl.status = self.counters.data_ptr() + 4 * 15
            pl.local_buf = bufs[r] + mine["base"]
            pl.self_pos = sum(1 for q in peers if q < r)
            if kind == "node" and peers:
                if_rows = run.if_rows
                pl.n_if = int(if_rows.numel())
                rows32 = if_rows.to(torch.int32).contiguous()
                self.keep.append(rows32)
                pl.if_rows = rows32.data_ptr()
            for i, q in enumerate(peers):
                theirs = lays[q][kind]
                pl.remote_buf[i] = bufs[q] + theirs["base"]
                pl.remote_off[i] = theirs["off"][r]
                pl.remote_parity_stride[i] = theirs["stride"]
                pl.remote_flag[i] = bufs[q] + (k * n_parts + r) * 8
                pl.local_flag[i] = bufs[r] + (k * n_parts + q) * 8
                pl.local_off[i] = mine["off"][q]
                if kind == "node":
                    send = t32(p.node_share[q])
                    inv = torch.full((pl.n_if,), -1, dtype=torch.int32, device=device)
                    inv[run.if_pos[q]] = torch.arange(send.numel(), dtype=torch.int32, device=device)
                    recv, n_recv = inv, int(send.numel())
                else:
                    send = t32(p.halo_send.get(q, np.zeros(0, dtype=np.int64)))
                    recv = t32(p.halo_recv.get(q, np.zeros(0, dtype=np.int64)))
                    n_recv = int(recv.numel())
                assert theirs["n"][r] == int(send.numel()) and mine["n"][q] == n_recv, (kind, r, q)
                self.keep += [send, recv]
                pl.send_rows[i], pl.n_send[i] = send.data_ptr(), int(send.numel())
                pl.recv_rows[i], pl.n_recv[i] = recv.data_ptr(), n_recv
            self.plans[kind] = pl

    def push(self, kind, src):
        from . import ops
        import ctypes as C_

        assert src.is_contiguous()
        ops._count(1)
        ops.check(ops._lib.lib().fvgn_exchange_push(C_.byref(self.plans[kind]), ops._p(src), src.stride(0) * src.element_size(), ops._stream()), "exchange_push")

    def wait(self, kind, dst):
        from . import ops
        import ctypes as C_

        assert dst.is_contiguous()
        L = ops._lib.lib()
        ops._count(1)
        if kind == "node":
            ops.check(L.fvgn_exchange_wait_node_sum(C_.byref(self.plans[kind]), ops._p(dst), ops._stream()), "exchange_wait_node_sum")
        else:
            ops.check(L.fvgn_exchange_wait_scatter(C_.byref(self.plans[kind]), ops._p(dst), dst.stride(0) * dst.element_size(), ops._stream()),
                      "exchange_wait_scatter")

    def timed_out(self):
        return int(self.counters[15]) != 0


class _PeerCommBase:
    fused = True  # rollout_steps: exchanges go through `exchange_rows` (push on every part, then wait on every part)

    def exchange_rows(self, kind, runs, tensors):
        for pr, t in zip(runs, tensors):
            self.ex[pr.p.rank].push(kind, t)
        for pr, t in zip(runs, tensors):
            self.ex[pr.p.rank].wait(kind, t)

    def check(self):
        bad = [r for r, e in self.ex.items() if e.timed_out()]
        if bad:
            raise RuntimeError(f"peer-memory exchange: a wait timed out on rank(s) {bad} (a neighbour never raised its flag)")


class LocalPeerComm(_PeerCommBase):
    """all parts in ONE process on one device: the "peer" buffers are the other parts' tensors (tests; the kernels and plans are the ones a
    rank per GPU runs)"""

    def __init__(self, parts, runs, device):
        self.bufs = {p.rank: torch.zeros(exchange_layout(parts, p.rank)["total"], dtype=torch.uint8, device=device) for p in parts}
        ptr = {r: b.data_ptr() for r, b in self.bufs.items()}
        self.ex = {pr.p.rank: PeerExchange(parts, pr, ptr, device) for pr in runs}


class SymmPeerComm(_PeerCommBase):
    """one rank per GPU: every rank's receive buffer is a symmetric allocation (torch.distributed._symmetric_memory: CUDA VMM + fabric / fd
    handles) that all peers map, so a rank's kernels store into its neighbours' HBM over NVLink / NVSwitch directly"""

    def __init__(self, parts, run, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        total = max(exchange_layout(parts, q)["total"] for q in range(len(parts)))
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        torch.cuda.synchronize(device)
        self.hdl.barrier()
        ptr = {q: int(self.hdl.buffer_ptrs[q]) for q in range(len(parts))}
        self.ex = {run.p.rank: PeerExchange(parts, run, ptr, device)}
        torch.cuda.synchronize(device)
        self.hdl.barrier()


# ------------------------------------------------------------------------------------------------ the part's rollout step
class PartRollout:
    """One rank's share of the rollout of ONE mesh.  `backend` supplies the arithmetic on torch tensors of this part:
    KernelBackend (libfvgn_b200, CUDA) or tests' OracleBackend (torch CPU, the oracle's functions)."""

    def __init__(self, part: Part, backend, x0, edge_attr0, bc_uv, device):
        """x0 [C_own+G, 4] = [type, Re, u, v] of the owned + ghost cells, edge_attr0 [F,5], bc_uv [F,2] (boundary-face velocities of the rollout)"""
        self.p, self.be, self.dev = part, backend, device
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(device)
        self.cnT, self.cfT = t(part.cells_node_T, torch.int32), t(part.cells_face_T, torch.int32)
        self.nc, self.fn = t(part.neighbour_cell, torch.int32), t(part.face_node, torch.int32)
        self.agg_dst = t(part.agg_dst, torch.int32)
        self.ftl = torch.stack([t(part.face_type, torch.float32), t(part.face_length, torch.float32)], dim=1).contiguous()
        self.unv, self.area = t(part.unv, torch.float32), t(part.cells_area, torch.float32).view(-1, 1)
        self.x, self.ea, self.bc = x0.to(device).clone(), edge_attr0.to(device).clone(), bc_uv.to(device).contiguous()
        self.share = {q: t(ix, torch.long) for q, ix in part.node_share.items()}
        self.hsend = {q: t(ix, torch.long) for q, ix in part.halo_send.items()}
        self.hrecv = {q: t(ix, torch.long) for q, ix in part.halo_recv.items()}
        self.agg = backend.build_agg(self.agg_dst, part.N + 1)
        # interface rows and their positions in the compact accumulation buffer: computed once (torch.unique synchronises)
        if self.share:
            self.if_rows = torch.unique(torch.cat([self.share[q] for q in sorted(self.share)]))
            pos = torch.full((part.N + 1,), -1, dtype=torch.long, device=device)
            pos[self.if_rows] = torch.arange(self.if_rows.numel(), device=device)
            self.if_pos = {q: pos[ix] for q, ix in self.share.items()}

    # -- exchanges (split in post / finish so that LocalComm can run all parts in one process)
    def node_sends(self, na):
        return {q: na.index_select(0, ix) for q, ix in self.share.items()}

    def node_finish(self, na, recvs):
        """interface nodes <- sum of the partial sums of all sharing ranks, added in ascending rank order (own partial at its rank's place):
        the same order on every rank, so a node shared by three parts gets the same bits everywhere"""
        if not recvs:
            return na
        rows = self.if_rows
        tot = torch.zeros((rows.numel(), na.shape[1]), dtype=na.dtype, device=na.device)
        for q in sorted(list(recvs) + [self.p.rank]):
            if q == self.p.rank:
                tot += na.index_select(0, rows)
            else:
                tot.index_add_(0, self.if_pos[q], recvs[q])  # (a row not shared with q is not in q's list: nothing is added)
        na[rows] = tot  # (na is this layer's own buffer)
        return na

    def halo_sends(self, buf):
        return {q: buf.index_select(0, ix) for q, ix in self.hsend.items()}

    def halo_finish(self, buf, recvs):
        for q, t in recvs.items():
            buf[self.hrecv[q]] = t
        return buf


def rollout_steps(parts_run, comm, model_state, n_steps):
    """drives `n_steps` rollout steps over a list of PartRollout (LocalComm: all parts; DistComm: this rank's single part).
    Returns per part the list of (uvp_on_local_faces [F,3], next_uv_of_owned_cells [C_own,2]) per step."""
    outs = [[] for _ in parts_run]

    def exchange(kind, payloads):
        """payloads: per part (sends dict, recv shapes dict, like tensor) -> per part recvs dict"""
        if isinstance(comm, LocalComm):
            for pr, (s, _, _) in zip(parts_run, payloads):
                comm.post(pr.p.rank, s)
            return [comm.collect(pr.p.rank) for pr in parts_run]
        return [comm.exchange(s, shapes, like) for (s, shapes, like) in payloads]

    fused = getattr(comm, "fused", False)  # peer-memory exchanges (csrc/halo.cu): two launches per exchange, rows consumed in place
    for _ in range(n_steps):
        st = [pr.be.encode(pr, model_state) for pr in parts_run]  # (x latent [C_own,128], e latent [F,128], uv_own)
        for layer in range(model_state["n_layers"]):
            nas = [pr.be.node_partial(pr, s, layer) for pr, s in zip(parts_run, st)]
            if fused:
                comm.exchange_rows("node", parts_run, nas)
            else:
                rec = exchange("node", [(pr.node_sends(na), {q: (ix.numel(), na.shape[1]) for q, ix in pr.share.items()}, na) for pr, na in zip(parts_run, nas)])
                nas = [pr.node_finish(na, r) for pr, na, r in zip(parts_run, nas, rec)]
            xps = [pr.be.cell_block(pr, s, na, layer, model_state) for pr, s, na in zip(parts_run, st, nas)]  # buffers [C_own+G, .]: owned rows filled
            if fused:
                comm.exchange_rows("halo", parts_run, xps)
            else:
                rec = exchange("halo", [(pr.halo_sends(xp), {q: (ix.numel(),) + tuple(xp.shape[1:]) for q, ix in pr.hrecv.items()}, xp) for pr, xp in zip(parts_run, xps)])
                xps = [pr.halo_finish(xp, r) for pr, xp, r in zip(parts_run, xps, rec)]
            for pr, s, xp in zip(parts_run, st, xps):
                pr.be.edge_block(pr, s, xp, layer, model_state)
        res = [pr.be.decode_integrate(pr, s, model_state) for pr, s in zip(parts_run, st)]  # (uvp [F,3], next_uv_own [C_own,2])
        uvb = []
        for pr, (uvp, nuv) in zip(parts_run, res):
            b = torch.zeros((pr.p.C_own + pr.p.G, 2), dtype=nuv.dtype, device=nuv.device)
            b[:pr.p.C_own] = nuv
            uvb.append(b)
        if fused:
            comm.exchange_rows("uv", parts_run, uvb)
        else:
            rec = exchange("uv", [(pr.halo_sends(b), {q: (ix.numel(), 2) for q, ix in pr.hrecv.items()}, b) for pr, b in zip(parts_run, uvb)])
            for pr, b, r in zip(parts_run, uvb, rec):
                pr.halo_finish(b, r)
        for pr, b, (uvp, nuv), o in zip(parts_run, uvb, res, outs):
            pr.be.next_graph(pr, b)
            o.append((uvp, nuv))
    return outs


class GraphedPartRollout:
    """One rollout step of this rank's part — its kernels AND its NCCL point-to-point exchanges — captured into one CUDA graph: a step is
    then one graph launch per rank instead of ~100 kernel launches, ~150 small torch ops and 31 NCCL batches issued from Python.  Every
    rank captures the same sequence of exchanges, so the replays pair up like the eager calls do.  The part's state (`run.x`, `run.ea`)
    lives in fixed buffers that the captured next-graph kernel advances in place; `step()` returns the graph's output buffers
    (uvp [F,3], next_uv [C_own,2]) — clone what has to outlive the next step."""

    def __init__(self, run, comm, model_state, warmup=2):
        """run: this rank's PartRollout (DistComm) — or the list of all parts' PartRollout in one process (LocalComm: tests)"""
        self.single = isinstance(run, PartRollout)
        self.runs = [run] if self.single else list(run)
        dev = self.runs[0].dev
        saved = [(r.x.clone(), r.ea.clone()) for r in self.runs]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(warmup):  # NCCL connections, weight packs, allocator: none of it may happen inside the capture
                rollout_steps(self.runs, comm, model_state, 1)
            for r, (x0, ea0) in zip(self.runs, saved):
                r.x.copy_(x0), r.ea.copy_(ea0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            outs = rollout_steps(self.runs, comm, model_state, 1)
        self.out = outs[0][0] if self.single else [o[0] for o in outs]

    def step(self):
        self.graph.replay()
        return self.out


# ------------------------------------------------------------------------------------------------ CUDA backend (libfvgn_b200 kernels)
class KernelBackend:
    """the part's arithmetic on the kernels of libfvgn_b200 (same entry points as the single-GPU forward: FVGN_model/FVGN.py here)"""

    def __init__(self, model):
        from . import ops
        from .autograd import network_mlps

        self.ops, self.model = ops, model
        self.mode = model.model.encoder.cb_encoder.math_mode_id()
        if self.mode == 1:
            raise NotImplementedError("domain decomposition runs the strict ('fp32') and the fp32-equivalent ('bf16x3') kernels")
        self.mlps = network_mlps(model)
        if self.mode != 0:
            for m in self.mlps:
                m.mlp_ref().ensure_packed(self.mode)

    @staticmethod
    def state(model):
        return {"n_layers": len(model.model.processer_list)}

    def build_agg(self, agg_dst, n_rows):
        return self.ops.build_csr(agg_dst, n_rows) + (n_rows,)

    def encode(self, pr, ms):
        m, ops = self.model, self.ops
        C = pr.p.C_own
        uv = pr.x[:C, 2:4].contiguous()
        cell_in = m._cell_normalizer(uv, False)
        edge_in = m._edge_normalizer(pr.ea, False, type_col=pr.ftl[:, 0:1], one_hot=9)
        x = m.model.encoder.cb_encoder(cell_in)
        e = m.model.encoder.eb_encoder(edge_in)
        return {"x": x, "e": e, "uv": uv}

    def node_partial(self, pr, s, layer):
        # rows 0 .. N-1 only: row N of the CSR is the dump of the face halves another rank aggregates (the cut faces of a part: thousands
        # of items that ONE half-warp would add up serially — 0.3 ms per layer for a sum nobody reads)
        ptr, items, n = pr.agg
        return self.ops.node_aggregate_fwd(s["e"], ptr, items, n - 1)

    def cell_block(self, pr, s, na, layer, ms):
        import ctypes as C_

        blk = self.model.model.processer_list[layer]
        C = pr.p.C_own
        ops = self.ops
        if self.mode == 2:
            # fp32-equivalent mode: the in-place split-shadow kernels of the single-GPU fast path; x' exists only as the fp16 hi / lo shadow
            # [C_own + G, 2, 128] whose ghost rows the owners send (512 B per row, as an fp32 row)
            L = ops._lib.lib()
            xps = torch.empty((C + pr.p.G, 2, 128), dtype=torch.float16, device=s["x"].device)
            cms = torch.empty((C, 2, 64), dtype=torch.float16, device=s["x"].device)
            mlp = blk.cb_module.net.mlp_ref().ensure_packed(2)
            ops._count(2)
            ops.check(L.fvgn_cell_mean_split(ops._p(na), ops._p(pr.cnT), C, ops._p(cms), ops._stream()), "cell_mean_split")
            ops.check(L.fvgn_cell_block_fwd_split(C_.byref(mlp.struct), ops._p(s["x"]), ops._p(cms), C, ops._p(xps), ops._stream()), "cell_block_fwd_split")
            return xps
        buf = torch.empty((C + pr.p.G, 128), dtype=torch.float32, device=s["x"].device)
        ops.cell_block_fwd(blk.cb_module.net.mlp_ref(), s["x"], na, pr.cnT, self.mode, x_prime=buf[:C], x_out=s["x"])
        return buf

    def edge_block(self, pr, s, xp, layer, ms):
        import ctypes as C_

        blk = self.model.model.processer_list[layer]
        ops = self.ops
        if self.mode == 2:
            mlp = blk.eb_module.net.mlp_ref().ensure_packed(2)
            ops._count(1)
            ops.check(ops._lib.lib().fvgn_edge_block_fwd_split(C_.byref(mlp.struct), ops._p(xp), ops._p(s["e"]), ops._p(pr.nc), pr.p.F, ops._stream()),
                      "edge_block_fwd_split")
            return
        ops.edge_block_fwd(blk.eb_module.net.mlp_ref(), xp, s["e"], pr.nc, self.mode, e_out=s["e"])

    def decode_integrate(self, pr, s, ms):
        m, ops = self.model, self.ops
        decoded = m.model.decoder.edge_decode_module(s["e"])
        bn = m._grid_feature_normalizer
        fa, _, _ = ops.face_area_bn_fwd(pr.ftl, pr.area, pr.nc, ms["dt"], False, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        du, _ = ops.integrator_fwd(decoded, pr.unv, fa, pr.cfT, ms["rho"], 1.0)
        nuv = m._acc_output_normalizer.inverse(du, addend=s["uv"])
        uvp = m._face_uvp_flux_output_normalizer.inverse(decoded, width=3)
        return uvp, nuv

    def next_graph(self, pr, uv_all):
        self.ops.create_next_graph_(pr.x, pr.ea, uv_all.contiguous(), pr.nc, pr.ftl, pr.bc)


def initial_part_state(part: Part, x_global, edge_attr_global, bc_uv_global):
    """slices of the single-domain rollout inputs (graph_cell.x [C,4], edge_attr [F,5], graph_edge.y[:, 0, 0:2]) for one part"""
    cells = torch.as_tensor(np.concatenate([part.own, part.ghost]))
    faces = torch.as_tensor(part.faces)
    return x_global.cpu()[cells], edge_attr_global.cpu()[faces], bc_uv_global.cpu()[faces]
