"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy + torch-CPU) of the reference's hot path.

This file is the parity oracle for the CUDA path.  It may be imported only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg; the product package
(`finite-volume-graph-network_b200/`) never imports it and has no CPU fallback.

Every function cites the reference lines (relative to /root/reference/FVGN-src/main) it restates.
PARITY PINNING: the reference ships no tests or golden vectors.  The oracle is pinned
instead against OUTPUTS OF THE UNMODIFIED REFERENCE executed in the build container
(oracle/ref_loader.py + oracle/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_*.py,
and live in the build container by tests/test_oracle_vs_reference.py).

The op structure deliberately mirrors the reference's torch-eager structure (cat / index_select / Linear
/ LayerNorm calls), because this file is also the "port" CPU baseline that bench.py times.
"""
from __future__ import annotations

import types
import numpy as np
import torch
import torch.nn.functional as Fnn

NORMAL, OBSTACLE, AIRFOIL, HANDLE, INFLOW, OUTFLOW, WALL_BOUNDARY = 0, 1, 2, 3, 4, 5, 6  # utils/utilities.py:6-15


class Bag(types.SimpleNamespace):
    """Attribute bag standing in for torch_geometric.data.Data (keys as property, num_nodes=x.shape[0])."""

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    @property
    def num_nodes(self):
        return self.x.shape[0]


def _torch_norm2_f32(first: np.ndarray, second: np.ndarray) -> np.ndarray:
    """Bit-exact restatement of torch.norm(v, dim=-1) for 2-vectors on CPU fp32: the reduction runs
    acc = first*first ; acc = fma(second, second, acc) ; sqrt(acc), every step rounded to fp32.
    (fma emulated in fp64: a 24x24-bit product is exact in fp64 and the sum is rounded once.)"""
    a2 = (first.astype(np.float32) * first.astype(np.float32)).astype(np.float32).astype(np.float64)
    s = second.astype(np.float64)
    return np.sqrt((s * s + a2).astype(np.float32)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# a1. connectivity + geometry tables: Extract_mesh/parse_tfrecord_refactor.py:419-971
# ----------------------------------------------------------------------------------------------
def build_connectivity(mesh_pos: np.ndarray, cells: np.ndarray, node_type: np.ndarray, face_rule: str = "B") -> dict:
    """Vectorised restatement of extract_mesh_state's tables (the reference uses Python dict loops).

    face_rule selects the face-typing branch of parse_tfrecord_refactor.py:491-620:
      "A" = the `if "compressible" in solving_params["name"]` branch (:491-509).  NOTE the reference quirk:
            the shipped __main__ passes name="incompressible" (:3032), which CONTAINS "compressible", so this
            is the branch the shipped converter actually takes: only AIRFOIL/NORMAL/INFLOW same-type pairs are
            typed, every other face (walls, outflow, mixed) is 0.
      "B" = the else branch (:511-620): WALL/INFLOW/OUTFLOW/NORMAL same-type pairs + the corner rules.

    Returns the H5-layout arrays WITHOUT the leading [1,...] axis:
      cells_node i32[C,3], face i32[2,F], cells_face i32[C,3], neighbour_cell i32[2,F],
      face_type i32[F,1], cells_type i32[C,1], face_length f32[F,1], unit_norm_v f32[C,3,2],
      centroid f32[C,2], cells_area f32[C,1], cell_factor f32[C,3]
    """
    pos = np.ascontiguousarray(mesh_pos, dtype=np.float32)
    nt = np.asarray(node_type).reshape(-1).astype(np.int32)
    cn = np.sort(np.asarray(cells), axis=1).astype(np.int32)  # :431
    C = cn.shape[0]
    # centroid: ((p0+p1)+p2)/3.0 in fp32 numpy  (:437-445)
    centroid = ((pos[cn[:, 0]] + pos[cn[:, 1]]) + pos[cn[:, 2]]) / np.float32(3.0)
    # triangles_to_faces (:1412-1437): edges (0,1),(1,2),(2,0) stacked, packed as (max,min), unique rows
    e0 = np.concatenate([cn[:, 0], cn[:, 1], cn[:, 2]])
    e1 = np.concatenate([cn[:, 1], cn[:, 2], cn[:, 0]])
    snd = np.maximum(e0, e1).astype(np.int64)
    rcv = np.minimum(e0, e1).astype(np.int64)
    key = (snd << 32) | rcv
    ukey, inv = np.unique(key, return_inverse=True)  # lexicographic (max,min) == torch.unique(dim=0)
    F = ukey.shape[0]
    face = np.stack([(ukey >> 32), (ukey & 0xFFFFFFFF)], 0).astype(np.int32)  # [2,F]  (:452-456)
    # cells_face[c,j] = index of packed[j*C+c]  (:627-631,692-695)
    cells_face = inv.reshape(3, C).T.astype(np.int32)
    # face_length = ||p[s]-p[r]|| (:459-467)
    d = pos[face[0]] - pos[face[1]]
    face_length = _torch_norm2_f32(d[:, 0], d[:, 1])[:, None]
    # face_type, incompressible branch (:470-620)
    a = nt[face[0]]
    b = nt[face[1]]
    fc = (pos[face[0]] + pos[face[1]]) / np.float32(2.0)
    outlet = np.max(fc[:, 0])
    inlet = np.min(fc[:, 0])
    ft = np.zeros(F, dtype=np.int32)
    if face_rule == "A":
        for T in (AIRFOIL, NORMAL, INFLOW):
            ft[(a == b) & (a == T)] = T
    elif face_rule == "B":
        for T in (WALL_BOUNDARY, INFLOW, OUTFLOW, NORMAL):
            ft[(a == b) & (a == T)] = T
        # limits are python floats (float64) compared against float32 face centres (:526-536, 569-612)
        wi = ((a == WALL_BOUNDARY) & (b == INFLOW)) | ((b == WALL_BOUNDARY) & (a == INFLOW))
        ft[wi & (fc[:, 0] < inlet + 1e-4) & (fc[:, 0] > inlet - 1e-4)] = INFLOW
        wo = ((a == WALL_BOUNDARY) & (b == OUTFLOW)) | ((b == WALL_BOUNDARY) & (a == OUTFLOW))
        ft[wo & (fc[:, 0] < outlet + 1e-4) & (fc[:, 0] > outlet - 1e-4)] = OUTFLOW
    else:
        raise ValueError(face_rule)
    # cells_type from its three face types (:639-691)
    t3 = ft[cells_face]  # [C,3]
    n_in = (t3 == INFLOW).sum(1)
    n_wall = (t3 == WALL_BOUNDARY).sum(1)
    n_out = (t3 == OUTFLOW).sum(1)
    n_foil = (t3 == AIRFOIL).sum(1)
    n_norm = 3 - n_in - n_wall - n_out - n_foil
    ct = np.zeros(C, dtype=np.int32)
    wall = (n_wall > 0) & (
        ((n_norm > 0) & (n_in == 0) & (n_out == 0))
        | ((n_norm == 0) & (n_in > 0) & (n_out == 0))
        | ((n_norm == 0) & (n_in == 0) & (n_out > 0))
        | ((n_norm > 0) & (n_in > 0) & (n_out == 0))
        | ((n_norm > 0) & (n_in == 0) & (n_out > 0))
    )
    foil = ~wall & (n_foil > 0) & (n_norm > 0) & (n_in == 0) & (n_out == 0)
    infl = ~wall & ~foil & (n_in > 0) & (n_norm > 0) & (n_wall == 0) & (n_out == 0)
    outf = ~wall & ~foil & ~infl & (n_out > 0) & (n_norm > 0) & (n_wall == 0) & (n_in == 0)
    ct[wall] = WALL_BOUNDARY
    ct[foil] = AIRFOIL
    ct[infl] = INFLOW
    ct[outf] = OUTFLOW
    # unit normals (:699-738): (-dy,dx)/||.||, flipped per cell so n.(centroid - face_centre) <= 0
    unv = np.stack([-d[:, 1], d[:, 0]], 1)
    unv = unv / _torch_norm2_f32(unv[:, 0], unv[:, 1])[:, None]
    unit_norm_v = np.zeros((C, 3, 2), dtype=np.float32)
    for j in range(3):
        f = cells_face[:, j]
        ccv = centroid - fc[f]
        dot = unv[f, 0] * ccv[:, 0] + unv[f, 1] * ccv[:, 1]
        unit_norm_v[:, j, :] = np.where((dot > 0)[:, None], unv[f] * np.float32(-1.0), unv[f])
    # neighbour_cell (:741-781): first / second cell in scan order c=0..C-1 (j inner); boundary -> (c,c)
    flat_f = cells_face.reshape(-1)  # scan order c-major, j-minor
    flat_c = np.repeat(np.arange(C, dtype=np.int64), 3)
    order = np.argsort(flat_f, kind="stable")
    sf = flat_f[order]
    sc = flat_c[order]
    first_pos = np.searchsorted(sf, np.arange(F), side="left")
    last_pos = np.searchsorted(sf, np.arange(F), side="right") - 1
    nb = np.stack([sc[first_pos], sc[last_pos]], 1)  # [F,2]: (first, second) or (c,c)
    # reorder_face (:1343-1375) on centroids: keep (a,b) iff dx>0 or (dx==0 and dy>0), else swap
    ev = centroid[nb[:, 0]] - centroid[nb[:, 1]]
    keep = (ev[:, 0] > 0) | ((ev[:, 0] == 0) & (ev[:, 1] > 0))
    neighbour_cell = np.where(keep[:, None], nb, nb[:, ::-1]).T.astype(np.int32)  # [2,F]
    # cell_factor (:792-869): ||p[cn_k]-centroid||^2 / sum_k
    dist = np.stack([((pos[cn[:, k]] - centroid) ** 2).sum(1) for k in range(3)], 1).astype(np.float32)
    tot = (dist[:, 0] + dist[:, 1]) + dist[:, 2]
    cell_factor = (dist / tot[:, None]).astype(np.float32)
    # cells_area by Heron from the three face lengths (:876-884)
    l1, l2, l3 = (face_length[cells_face[:, j], 0] for j in range(3))
    p = np.float32(0.5) * ((l1 + l2) + l3)
    cells_area = np.sqrt(p * (p - l1) * (p - l2) * (p - l3)).astype(np.float32)[:, None]
    return dict(
        mesh_pos=pos, node_type=nt[:, None], cells_node=cn, face=face, cells_face=cells_face,
        neighbour_cell=neighbour_cell, face_type=ft[:, None], cells_type=ct[:, None],
        face_length=face_length, unit_norm_v=unit_norm_v, centroid=centroid.astype(np.float32),
        cells_area=cells_area, cell_factor=cell_factor,
    )


# ----------------------------------------------------------------------------------------------
# L1 graph assembly: dataset/Load_mesh.py:589-690 (_build_graph) and :467-587 (get)
# ----------------------------------------------------------------------------------------------
def build_graphs(tab: dict, velocity: np.ndarray, pressure: np.ndarray, training_pair: int | None = None):
    """Returns (graph_node, graph_edge, graph_cell) bags of torch CPU tensors with int64 indices.
    training_pair=t -> the `get()` layout: graph_node.x[N,6] = [u,v,p](t) ++ [u,v,p](t+1)  (:516-534);
    None -> the rollout layout: graph_node.x[N,T,3]  (:624-630)."""
    tl = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(torch.long)
    tf = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
    vp = torch.cat((tf(velocity), tf(pressure)), dim=2)  # [T,N,3]
    if training_pair is None:
        node_x = vp.transpose(0, 1).contiguous()
    else:
        node_x = torch.cat(torch.unbind(vp[training_pair:training_pair + 2]), dim=1)
    cells_type = tl(tab["cells_type"])
    cell_attr = torch.cat((cells_type, torch.zeros_like(cells_type)), dim=1)  # :545 / :652
    graph_node = Bag(x=node_x, edge_index=tl(tab["face"]), face=tl(tab["cells_node"]).T.contiguous(),
                     pos=tf(tab["mesh_pos"]))
    face_types = tl(tab["face_type"])
    graph_edge = Bag(x=torch.cat((face_types, tf(tab["face_length"])), dim=1),
                     face=tl(tab["cells_face"]).T.contiguous(), y=face_types)
    graph_cell = Bag(x=cell_attr, edge_index=tl(tab["neighbour_cell"]), unv=tf(tab["unit_norm_v"]),
                     centroid=tf(tab["centroid"]), cell_area=tf(tab["cells_area"]),
                     cell_factor=tf(tab["cell_factor"]), pos=tf(tab["centroid"]))
    return graph_node, graph_edge, graph_cell


def batch_graphs(bags: list):
    """PyG Batch.from_data_list rule: keys containing 'index' or == 'face' get the running num_nodes offset
    and are concatenated on the last dim; the rest on dim 0; adds `.batch`."""
    out = Bag()
    off = 0
    cols = {k: [] for k in bags[0].keys}
    bvec = []
    for gi, g in enumerate(bags):
        n = g.num_nodes
        for k in cols:
            v = getattr(g, k)
            cols[k].append(torch.where(v >= 0, v + off, v) if ("index" in k or k == "face") else v)  # (-1 = padding of [4,C] polygon tables)
        bvec.append(torch.full((n,), gi, dtype=torch.long))
        off += n
    for k, vs in cols.items():
        setattr(out, k, torch.cat(vs, dim=-1 if ("index" in k or k == "face") else 0))
    out.batch = torch.cat(bvec)
    out.num_graphs = len(bags)
    return out


# ----------------------------------------------------------------------------------------------
# a2/a3. pre/post step: dataset/Load_mesh.py:904-1115
# ----------------------------------------------------------------------------------------------
def face_masks(graph_edge):
    """Load_mesh.py:931-935 / 1014-1018"""
    interior = torch.logical_or(graph_edge.x[:, 0] == NORMAL, graph_edge.x[:, 0] == OUTFLOW)
    return interior, torch.logical_not(interior)


def valid_datapreprocessing(graph_node, graph_edge, graph_cell, start_epoch=0):
    """Load_mesh.py:904-981 (dual_edge=False).  graph_node.x is [N,T,3]."""
    cf = graph_cell.cell_factor
    x = graph_node.x
    target_on_cell = (cf[:, 0:1].unsqueeze(1) * torch.index_select(x, 0, graph_node.face[0])
                      + cf[:, 1:2].unsqueeze(1) * torch.index_select(x, 0, graph_node.face[1])
                      + cf[:, 2:3].unsqueeze(1) * torch.index_select(x, 0, graph_node.face[2]))
    if graph_node.face.shape[0] == 4:  # tri/quad tables (oracle/fvgn_oracle_poly.py): 4th node, weight 0 / index -1 on triangles
        target_on_cell = target_on_cell + cf[:, 3:4].unsqueeze(1) * torch.index_select(x, 0, graph_node.face[3].clamp(min=0))
    target_on_edge = (torch.index_select(x, 0, graph_node.edge_index[0])
                      + torch.index_select(x, 0, graph_node.edge_index[1])) / 2.0
    graph_cell.y = target_on_cell[:, start_epoch + 1:, :]
    graph_edge.y = target_on_edge[:, start_epoch + 1:, :]
    senders, receivers = graph_cell.edge_index[0], graph_cell.edge_index[1]
    interior, boundary = face_masks(graph_edge)
    uv0 = target_on_cell[:, start_epoch, 0:2]
    graph_cell.x = torch.cat((graph_cell.x, uv0), dim=1)
    rel = torch.index_select(graph_cell.pos, 0, senders) - torch.index_select(graph_cell.pos, 0, receivers)
    euv = torch.index_select(graph_cell.x[:, 2:4], 0, senders) - torch.index_select(graph_cell.x[:, 2:4], 0, receivers)
    euv[boundary, 0:2] = target_on_edge[boundary, 0, 0:2]
    graph_cell.edge_attr = torch.cat((euv, rel, graph_edge.x[:, 1:2]), dim=1)
    return graph_node, graph_edge, graph_cell, interior, boundary


def train_datapreprocessing(graph_node, graph_edge, graph_cell, cell_noise, flip_keep):
    """Load_mesh.py:983-1074 (dual_edge=False, is_training=True).  graph_node.x is [N,6].
    `cell_noise` f32[C,2] (utils/noise.py:13-26) and `flip_keep` bool[F] (randbool, :442-446,1046) are
    INPUTS here so both implementations consume the same random draws."""
    cf = graph_cell.cell_factor
    x = graph_node.x
    target_on_cell = (cf[:, 0:1] * torch.index_select(x, 0, graph_node.face[0])
                      + cf[:, 1:2] * torch.index_select(x, 0, graph_node.face[1])
                      + cf[:, 2:3] * torch.index_select(x, 0, graph_node.face[2]))
    if graph_node.face.shape[0] == 4:
        target_on_cell = target_on_cell + cf[:, 3:4] * torch.index_select(x, 0, graph_node.face[3].clamp(min=0))
    target_on_edge = (torch.index_select(x, 0, graph_node.edge_index[0])
                      + torch.index_select(x, 0, graph_node.edge_index[1])) / 2.0
    graph_cell.y = target_on_cell[:, 3:5]
    graph_edge.y = target_on_edge[:, 3:6]
    senders, receivers = graph_cell.edge_index[0], graph_cell.edge_index[1]
    interior, boundary = face_masks(graph_edge)
    noised = target_on_cell[:, 0:2] + cell_noise
    graph_cell.x = torch.cat((graph_cell.x, noised), dim=1)
    mask = flip_keep.view(1, -1).repeat(2, 1)
    rde = torch.where(mask, torch.stack((senders, receivers), 0), torch.stack((receivers, senders), 0))
    rel = torch.index_select(graph_cell.pos, 0, rde[0]) - torch.index_select(graph_cell.pos, 0, rde[1])
    euv = torch.index_select(graph_cell.x[:, 2:4], 0, rde[0]) - torch.index_select(graph_cell.x[:, 2:4], 0, rde[1])
    euv[boundary, 0:2] = target_on_edge[boundary, 0:2]
    graph_cell.edge_attr = torch.cat((euv, rel, graph_edge.x[:, 1:2]), dim=1)
    return graph_node, graph_edge, graph_cell, interior, boundary


def create_next_graph(graph_old, graph_edge, mask_face_boundary, cell_next_UV, cell_type, Re, edge_RMP_EU):
    """Load_mesh.py:1076-1115 (dual_edge=False)."""
    graph_old.x = torch.cat((cell_type, Re, cell_next_UV), dim=1)
    euv = cell_next_UV[graph_old.edge_index[0]] - cell_next_UV[graph_old.edge_index[1]]
    euv[mask_face_boundary, 0:2] = graph_edge.y[mask_face_boundary, 0, 0:2]
    graph_old.edge_attr = torch.cat((euv, edge_RMP_EU), dim=1)
    return graph_old


# ----------------------------------------------------------------------------------------------
# a4/a13. Normalizer: utils/normalization.py:5-86
# ----------------------------------------------------------------------------------------------
class NormState:
    """acc_count / num_accumulations / acc_sum / acc_sum_squared, all initialised to ONE (:26-31)."""

    def __init__(self, size, max_accumulations, dtype=torch.float32):
        self.max_accumulations = max_accumulations
        self.epsilon = 1e-8
        self.acc_count = torch.tensor(1.0, dtype=dtype)
        self.num_accumulations = torch.tensor(1.0, dtype=dtype)
        self.acc_sum = torch.ones(size, dtype=dtype)
        self.acc_sum_squared = torch.ones(size, dtype=dtype)

    def mean(self):  # :69-75
        return self.acc_sum / torch.clamp(self.acc_count, min=1.0)

    def std(self):  # :77-86
        sc = torch.clamp(self.acc_count, min=1.0)
        std = torch.sqrt(self.acc_sum_squared / sc - self.mean() ** 2)
        std = std.clone()
        std[std < self.epsilon] = 1.0
        return std

    def __call__(self, data, accumulate=True):  # :33-45, 56-67
        if accumulate and self.num_accumulations < self.max_accumulations:
            self.acc_sum = self.acc_sum + data.sum(0)
            self.acc_sum_squared = self.acc_sum_squared + (data ** 2).sum(0)
            self.acc_count = self.acc_count + float(data.shape[0])
            self.num_accumulations = self.num_accumulations + 1
        return (data - self.mean()) / self.std()

    def inverse(self, data):  # :47-54
        return data * self.std() + self.mean()

    def load(self, sd, prefix):
        for k in ("acc_count", "num_accumulations", "acc_sum", "acc_sum_squared"):
            setattr(self, k, sd[prefix + k].detach().clone().to(self.acc_sum.dtype))


# ----------------------------------------------------------------------------------------------
# a5-a9. network: FVGN_model/EncoderProcesserDecoder.py, GN_blocks.py
# ----------------------------------------------------------------------------------------------
# Emulation of the tensor-core modes for the parity tests of the bf16 path (NOT part of the reference): when set to
# torch.bfloat16, both operands of every nn.Linear are rounded to bf16 (round-to-nearest-even) before the product,
# accumulation stays in the working dtype — exactly what SURVEY.md Appendix E measured.
OPERAND_ROUND = None


def _rnd(t):
    return t if OPERAND_ROUND is None else t.to(OPERAND_ROUND).to(t.dtype)


def _linear(x, w, b):
    return Fnn.linear(_rnd(x), _rnd(w), b)


def mlp(sd, prefix, x, layer_norm=True):
    """build_mlp (EncoderProcesserDecoder.py:12-48): Linear-SiLU-Linear-SiLU-Linear [+ LayerNorm eps 1e-5]."""
    p = prefix + ("0." if layer_norm else "")
    h = Fnn.silu(_linear(x, sd[p + "0.weight"], sd[p + "0.bias"]))
    h = Fnn.silu(_linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
    h = _linear(h, sd[p + "4.weight"], sd[p + "4.bias"])
    if layer_norm:
        h = Fnn.layer_norm(h, (h.shape[-1],), sd[prefix + "1.weight"], sd[prefix + "1.bias"], 1e-5)
    return h


def cell_block(sd, prefix, x, edge_attr, node_edge_index, node_face, num_nodes, mp_times=2, edge_index=None):
    """CellBlock.forward, dual_edge=False (GN_blocks.py:78-145); mp_times >= 2: faces -> nodes -> cells, mp_times = 1 (needs the
    cell graph's edge_index): faces -> their two cells directly."""
    twoway = torch.cat(torch.chunk(edge_attr, 2, dim=-1), dim=0)  # :91
    if mp_times < 2:
        cidx = torch.cat([edge_index[0], edge_index[1]], dim=0)  # :131-132
        cell_agg = torch.zeros((x.shape[0], twoway.shape[1]), dtype=x.dtype).index_add_(0, cidx, twoway)  # :133-138
        return mlp(sd, prefix + "net.", torch.cat((x, cell_agg), dim=-1))  # :140-143
    idx = torch.cat([node_edge_index[0], node_edge_index[1]], dim=0)  # :111-114
    node_agg = torch.zeros((num_nodes, twoway.shape[1]), dtype=x.dtype).index_add_(0, idx, twoway)  # :117-122
    if node_face.shape[0] == 4:  # tri/quad tables: (a + b + c [+ d]) / k, the 4th term absent (index -1) on triangles
        has4 = (node_face[3] >= 0).unsqueeze(1)
        s = (torch.index_select(node_agg, 0, node_face[0]) + torch.index_select(node_agg, 0, node_face[1])
             + torch.index_select(node_agg, 0, node_face[2]))
        s = torch.where(has4, s + torch.index_select(node_agg, 0, node_face[3].clamp(min=0)), s)
        cell_agg = s / torch.where(has4, 4.0, 3.0).to(s.dtype)
        return mlp(sd, prefix + "net.", torch.cat((x, cell_agg), dim=-1))
    cell_agg = (torch.index_select(node_agg, 0, node_face[0]) + torch.index_select(node_agg, 0, node_face[1])
                + torch.index_select(node_agg, 0, node_face[2])) / 3.0  # :125-129
    return mlp(sd, prefix + "net.", torch.cat((x, cell_agg), dim=-1))  # :140-143


def edge_block(sd, prefix, x, edge_attr, edge_index):
    """EdgeBlock.forward, dual_edge=False (GN_blocks.py:16-54)."""
    collected = torch.cat((x[edge_index[0]], x[edge_index[1]], edge_attr), dim=1)  # :43-50
    return mlp(sd, prefix + "net.", collected)


def gn_block(sd, prefix, x, e, edge_index, node_edge_index, node_face, num_nodes, mp_times=2):
    """GnBlock.forward (EncoderProcesserDecoder.py:112-124): Cell then Edge, residuals at the end."""
    xn = cell_block(sd, prefix + "cb_module.", x, e, node_edge_index, node_face, num_nodes, mp_times, edge_index)
    en = edge_block(sd, prefix + "eb_module.", xn, e, edge_index)
    return x + xn, e + en


def epd_forward(sd, prefix, x, edge_attr, edge_index, node_edge_index, node_face, num_nodes, n_mp=15,
                capture=None, mp_times=2):
    """EncoderProcesserDecoder.forward (EncoderProcesserDecoder.py:224-238)."""
    xc = mlp(sd, prefix + "encoder.cb_encoder.", x)  # Encoder :65-76
    e = mlp(sd, prefix + "encoder.eb_encoder.", edge_attr)
    if capture is not None:
        capture["enc_x"], capture["enc_e"] = xc, e
    for l in range(n_mp):
        xc, e = gn_block(sd, f"{prefix}processer_list.{l}.", xc, e, edge_index, node_edge_index, node_face, num_nodes, mp_times)
        if capture is not None and l in (0, 1, n_mp - 1):
            capture[f"x_l{l}"], capture[f"e_l{l}"] = xc, e
    return mlp(sd, prefix + "decoder.edge_decode_module.", e, layer_norm=False)  # Decoder :161-169


# ----------------------------------------------------------------------------------------------
# a11. Intergrator: EncoderProcesserDecoder.py:251-331
# ----------------------------------------------------------------------------------------------
def integrator(uv_face, p_face, flux_D, unv, rho, rhs_coef, face_area, cell_face, loss_cont=0):
    if cell_face.shape[0] == 4:
        return _integrator_poly(uv_face, p_face, flux_D, unv, rho, rhs_coef, face_area, cell_face, loss_cont)
    e = [torch.index_select(face_area, 0, cell_face[k]) for k in range(3)]
    uu_vu = torch.cat((uv_face[:, 0:1] * uv_face, uv_face[:, 1:2] * uv_face), dim=-1)

    def dot(a, b):
        return torch.sum(a * b, dim=-1, keepdim=True)

    def flux_dot(a, b):
        return torch.cat([dot(a[:, i:i + 2], b) for i in range(0, a.size(1), 2)], dim=-1)

    cont = 0
    if loss_cont > 0:
        cont = sum(dot(torch.index_select(uv_face, 0, cell_face[k]), unv[:, k, :]) * e[k] for k in range(3))
    A = (flux_dot(uu_vu[cell_face[0]], unv[:, 0, :]) * e[0] + flux_dot(uu_vu[cell_face[1]], unv[:, 1, :]) * e[1]
         + flux_dot(uu_vu[cell_face[2]], unv[:, 2, :]) * e[2])
    D = (torch.index_select(flux_D, 0, cell_face[0]) + torch.index_select(flux_D, 0, cell_face[1])
         + torch.index_select(flux_D, 0, cell_face[2]))
    P = (torch.index_select(p_face, 0, cell_face[0]) * unv[:, 0, :] * e[0]
         + torch.index_select(p_face, 0, cell_face[1]) * unv[:, 1, :] * e[1]
         + torch.index_select(p_face, 0, cell_face[2]) * unv[:, 2, :] * e[2])
    return cont, rhs_coef * ((-A - (1.0 / rho) * P)) + D


def _integrator_poly(uv_face, p_face, flux_D, unv, rho, rhs_coef, face_area, cell_face, loss_cont=0):
    """the same sums over the k = 3 or 4 faces of a cell ([4,C] table, -1 = no 4th face; oracle/fvgn_oracle_poly.py): the first three terms
    are added exactly as above, the 4th only where it exists"""
    has4 = (cell_face[3] >= 0).unsqueeze(1)
    cf = cell_face.clamp(min=0)
    e = [torch.index_select(face_area, 0, cf[k]) for k in range(4)]
    uu_vu = torch.cat((uv_face[:, 0:1] * uv_face, uv_face[:, 1:2] * uv_face), dim=-1)

    def dot(a, b):
        return torch.sum(a * b, dim=-1, keepdim=True)

    def flux_dot(a, b):
        return torch.cat([dot(a[:, i:i + 2], b) for i in range(0, a.size(1), 2)], dim=-1)

    def plus4(s, t):
        return torch.where(has4, s + t, s)

    cont = 0
    if loss_cont > 0:
        cont = sum(dot(torch.index_select(uv_face, 0, cf[k]), unv[:, k, :]) * e[k] for k in range(3))
        cont = plus4(cont, dot(torch.index_select(uv_face, 0, cf[3]), unv[:, 3, :]) * e[3])
    A = (flux_dot(uu_vu[cf[0]], unv[:, 0, :]) * e[0] + flux_dot(uu_vu[cf[1]], unv[:, 1, :]) * e[1]
         + flux_dot(uu_vu[cf[2]], unv[:, 2, :]) * e[2])
    A = plus4(A, flux_dot(uu_vu[cf[3]], unv[:, 3, :]) * e[3])
    D = (torch.index_select(flux_D, 0, cf[0]) + torch.index_select(flux_D, 0, cf[1]) + torch.index_select(flux_D, 0, cf[2]))
    D = plus4(D, torch.index_select(flux_D, 0, cf[3]))
    P = (torch.index_select(p_face, 0, cf[0]) * unv[:, 0, :] * e[0] + torch.index_select(p_face, 0, cf[1]) * unv[:, 1, :] * e[1]
         + torch.index_select(p_face, 0, cf[2]) * unv[:, 2, :] * e[2])
    P = plus4(P, torch.index_select(p_face, 0, cf[3]) * unv[:, 3, :] * e[3])
    return cont, rhs_coef * ((-A - (1.0 / rho) * P)) + D


# ----------------------------------------------------------------------------------------------
# a3-a13. FVGN.forward: FVGN_model/FVGN.py:147-326
# ----------------------------------------------------------------------------------------------
class FVGNOracle:
    """Functional FVGN over a reference-layout state_dict."""

    def __init__(self, state_dict, params=None, dtype=torch.float32):
        from types import SimpleNamespace

        self.params = params or SimpleNamespace(
            edge_input_size=5, edge_one_hot=9, cell_input_size=2, cell_one_hot=0, message_passing_num=15,
            loss_cont=0, dataset_size=1000, batch_size=16, cumulative_length=600)
        self.dtype = dtype
        self.training = False
        self.load_state_dict(state_dict)

    def load_state_dict(self, state_dict):
        p = self.params
        self.sd = {k: v.detach().clone().to(self.dtype) if v.is_floating_point() else v.clone()
                   for k, v in state_dict.items()}
        big = (int(p.dataset_size / p.batch_size) + 1) * p.cumulative_length  # FVGN.py:47-74
        small = (int(1000 / p.batch_size) + 1) * p.cumulative_length
        self.norm = {
            "_acc_output_normalizer.": NormState(2, big, self.dtype),
            "_face_uvp_flux_output_normalizer.": NormState(3, small, self.dtype),
            "_edge_normalizer.": NormState(p.edge_input_size + p.edge_one_hot, big, self.dtype),
            "_cell_normalizer.": NormState(p.cell_input_size + p.cell_one_hot, big, self.dtype),
        }
        for k, n in self.norm.items():
            n.load(self.sd, k)
        g = "_grid_feature_normalizer."
        self.bn = dict(weight=self.sd[g + "weight"], bias=self.sd[g + "bias"],
                       running_mean=self.sd[g + "running_mean"].clone(), running_var=self.sd[g + "running_var"].clone(),
                       num_batches_tracked=self.sd[g + "num_batches_tracked"].clone())

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def _edge_feats(self, edge_attr, faces_type, one_hot):
        oh = Fnn.one_hot(faces_type.long().view(-1), one_hot)  # FVGN.py:127-129
        return torch.cat((edge_attr, oh), dim=1)

    def _face_area(self, graph, graph_edge, dt):
        """FVGN.py:209-227 / 286-304: BatchNorm1d(1) of len*dt/((A[s]+A[r])/2); eps 1e-5, momentum 0.1."""
        raw = (graph_edge.x[:, 1:2] * (dt / ((torch.index_select(graph.cell_area, 0, graph.edge_index[0])
                                                + torch.index_select(graph.cell_area, 0, graph.edge_index[1])) / 2))).view(-1, 1)
        bn = self.bn
        if self.training:
            bn["num_batches_tracked"] += 1
        return Fnn.batch_norm(raw, bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"],
                              self.training, 0.1, 1e-5), raw

    def forward(self, graph, graph_edge, graph_node, rho, mu, dt, edge_one_hot=9, cell_one_hot=0, capture=None):
        p = self.params
        cells_type = graph.x[:, 0:1]
        faces_type = graph_edge.x[:, 0:1]
        uv = graph.x[:, 2:4]
        nz = self.norm
        if self.training:
            target_acc = graph.y[:, 0:2] - uv[:, 0:2]  # FVGN.py:142-145,173-176
            target_delta_u = nz["_acc_output_normalizer."](target_acc, True)
            target_face = nz["_face_uvp_flux_output_normalizer."](graph_edge.y[:, 0:3], True)
        graph.x = nz["_cell_normalizer."](uv, self.training)  # cell_one_hot == 0: FVGN.py:92-106
        graph.edge_attr = nz["_edge_normalizer."](self._edge_feats(graph.edge_attr, faces_type, edge_one_hot), self.training)
        if capture is not None:
            capture["cell_in"], capture["edge_in"] = graph.x, graph.edge_attr
        pred = epd_forward(self.sd, "model.", graph.x, graph.edge_attr, graph.edge_index, graph_node.edge_index,
                           graph_node.face, graph_node.num_nodes, p.message_passing_num, capture, getattr(p, "mp_times", 2))
        face_area, raw = self._face_area(graph, graph_edge, dt)
        cont, delta_u = integrator(pred[:, 0:2], pred[:, 2:3].clone().detach(), pred[:, 3:5], graph.unv, rho, 1.0,
                                   face_area, graph_edge.face, p.loss_cont)
        if capture is not None:
            capture["decoded"], capture["face_area"], capture["delta_u"] = pred, face_area, delta_u
        if self.training:
            return cont, delta_u, target_delta_u, pred, target_face  # FVGN.py:241-247
        next_uv = nz["_acc_output_normalizer."].inverse(delta_u) + uv  # :318-321
        uvp_face = nz["_face_uvp_flux_output_normalizer."].inverse(pred[:, 0:3])  # :322-324
        return uvp_face, next_uv

    __call__ = forward


# ----------------------------------------------------------------------------------------------
# a12. loss: utils/loss_compute.py:24-115
# ----------------------------------------------------------------------------------------------
def _segment_mean(x, batch, size):
    s = torch.zeros((size,) + tuple(x.shape[1:]), dtype=x.dtype).index_add_(0, batch, x)
    c = torch.zeros(size, dtype=x.dtype).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return s / c.clamp(min=1).view(-1, 1)


def compute_fvm_loss(params, cell_batch, face_batch, mask_face_interior, predicted_edge_attr,
                     target_face_uvp_normalized, predicted_delta_u, target_delta_u, num_graphs=None):
    """Returns (loss, loss_continuity(=0.0, Appendix D.2), loss_mom, loss_face_uv, loss_face_p)."""
    m = mask_face_interior
    if "global_mean" in params.loss:
        B = num_graphs if num_graphs is not None else int(cell_batch.max().item()) + 1
        red = (lambda t: t.sum(dim=-1, keepdim=True)) if params.loss == "global_mean_sum" else (lambda t: t.mean(dim=-1, keepdim=True))
        l_uv = _segment_mean((target_face_uvp_normalized[m, 0:2] - predicted_edge_attr[m, 0:2]) ** 2, face_batch[m], B).sum(dim=-1, keepdim=True)
        l_uv = red(l_uv)
        l_p = red(_segment_mean((target_face_uvp_normalized[:, 2:3] - predicted_edge_attr[:, 2:3]) ** 2, face_batch, B))
        l_mom = red(_segment_mean((target_delta_u - predicted_delta_u) ** 2, cell_batch, B))
    else:
        l_uv = Fnn.mse_loss(predicted_edge_attr[m, 0:2], target_face_uvp_normalized[m, 0:2])
        l_p = Fnn.mse_loss(predicted_edge_attr[:, 2:3], target_face_uvp_normalized[:, 2:3])
        l_mom = Fnn.mse_loss(predicted_delta_u, target_delta_u)
    loss = params.loss_cont * 0.0 + params.loss_mom * l_mom + params.loss_face_uv * l_uv + params.loss_face_p * l_p
    return torch.mean(torch.log(loss)), 0.0, l_mom, l_uv, l_p


# ----------------------------------------------------------------------------------------------
# rollout-error metric: rollout.py:446-482
# ----------------------------------------------------------------------------------------------
def rollout_relative_error(pred_uv, target_uv, pred_p, target_p):
    """pred_uv/target_uv [T,C,2], pred_p/target_p [T,F,1] -> dict of cumulative mean relative MSE series."""
    steps = torch.arange(1, pred_uv.shape[0] + 1, dtype=pred_uv.dtype)
    se_uv = ((target_uv - pred_uv) ** 2).sum(dim=[1, 2])
    se_p = ((target_p - pred_p) ** 2).sum(dim=[1, 2])
    den_uv = (pred_uv ** 2).sum(dim=[1, 2])
    den_p = (pred_p ** 2).sum(dim=[1, 2])
    return dict(uv_rmse=torch.cumsum(se_uv / den_uv, 0) / steps, p_rmse=torch.cumsum(se_p / den_p, 0) / steps,
                uv_mse=torch.cumsum(((target_uv - pred_uv) ** 2).mean(dim=[1, 2]), 0) / steps,
                p_mse=torch.cumsum(((target_p - pred_p) ** 2).mean(dim=[1, 2]), 0) / steps)
