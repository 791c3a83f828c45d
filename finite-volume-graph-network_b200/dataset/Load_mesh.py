"""Device-side pre/post step of the hot path — the arithmetic of Data_Pool.valid_datapreprocessing /
train_datapreprocessing / create_next_graph (dataset/Load_mesh.py:904-1115) as three small kernels, plus graph assembly
from connectivity tables (the role of Data_Pool._build_graph / get, :467-690, without the H5/TFRecord readers, which are
out of scope: SURVEY.md §2 row 8)."""
import torch

from .. import ops
from ..data import Data, Batch
from ..utils.utilities import NodeType


def build_graphs(tab: dict, velocity, pressure, training_pair=None, device=None):
    """tables (H5 layout of extract_mesh_state, numpy or torch) + fields -> [graph_node, graph_edge, graph_cell] with the
    reference's attribute names, dtypes (int64 indices) and layouts.  training_pair=t gives `get()`'s [N,6] node field
    (steps t, t+1); None gives the rollout layout [N,T,3]."""
    dev = device
    tl = lambda a: torch.as_tensor(a).to(device=dev, dtype=torch.long)
    tf = lambda a: torch.as_tensor(a).to(device=dev, dtype=torch.float32)
    vp = torch.cat((tf(velocity), tf(pressure)), dim=2)
    node_x = vp.transpose(0, 1).contiguous() if training_pair is None else torch.cat(torch.unbind(vp[training_pair:training_pair + 2]), dim=1).contiguous()
    cells_type = tl(tab["cells_type"]).view(-1, 1)
    cell_attr = torch.cat((cells_type, torch.zeros_like(cells_type)), dim=1)
    face_types = tl(tab["face_type"]).view(-1, 1)
    graph_node = Data(x=node_x, edge_index=tl(tab["face"]), face=tl(tab["cells_node"]).T.contiguous(), pos=tf(tab["mesh_pos"]))
    graph_edge = Data(x=torch.cat((face_types.float(), tf(tab["face_length"]).view(-1, 1)), dim=1), face=tl(tab["cells_face"]).T.contiguous(),
                      y=face_types)
    graph_cell = Data(x=cell_attr, edge_index=tl(tab["neighbour_cell"]), unv=tf(tab["unit_norm_v"]), centroid=tf(tab["centroid"]),
                      cell_area=tf(tab["cells_area"]).view(-1, 1), cell_factor=tf(tab["cell_factor"]), pos=tf(tab["centroid"]))
    return [graph_node, graph_edge, graph_cell]


def collate(lists):
    """PyG DataLoader collation of [graph_node, graph_edge, graph_cell] samples into three Batches."""
    return [Batch.from_data_list([l[k] for l in lists]) for k in range(3)]


def _idx(graph_node, graph_cell):
    return ops.as_i32(graph_node.face), ops.as_i32(graph_node.edge_index), ops.as_i32(graph_cell.edge_index)


def valid_datapreprocessing(graph_node, graph_edge, graph_cell, start_epoch=0, dual_edge=False):
    """Load_mesh.py:904-981.  graph_node.x is [N,T,3]; returns [graph_node, graph_edge, graph_cell, mask_interior, mask_boundary]."""
    if dual_edge:
        raise NotImplementedError("dual_edge=True is outside the B200 hot path")
    cnT, fn, nc = _idx(graph_node, graph_cell)
    N, T, W = graph_node.x.shape
    flat = graph_node.x.reshape(N, T * W)
    on_cell, on_face = ops.interp_nodes(flat, 0, T * W, cnT, graph_cell.cell_factor.contiguous(), fn)
    on_cell, on_face = on_cell.view(-1, T, W), on_face.view(-1, T, W)
    graph_cell.y = on_cell[:, start_epoch + 1:, :]
    graph_edge.y = on_face[:, start_epoch + 1:, :]
    uv0 = on_cell[:, start_epoch, 0:2]
    graph_cell.x = torch.cat((graph_cell.x.to(torch.float32), uv0), dim=1).contiguous()
    ftl = graph_edge.x.contiguous()
    graph_cell.edge_attr, interior = ops.assemble_edge_attr(graph_cell.x[:, 2:4], graph_cell.pos.contiguous(), nc, ftl, on_face[:, 0, 0:2])
    return [graph_node, graph_edge, graph_cell, interior, torch.logical_not(interior)]


def randbool(*size, device=None):
    """50% True (Load_mesh.py:442-446)."""
    return torch.randint(2, size, device=device) == torch.randint(2, size, device=device)


def train_datapreprocessing(graph_list, cell_noise=None, flip_keep=None, noise_std=0.02, dual_edge=False):
    """Load_mesh.py:983-1074 (is_training=True).  graph_list = [graph_node, graph_edge, graph_cell] with graph_node.x [N,6].
    `cell_noise` [C,2] and `flip_keep` bool[F] default to fresh device draws; tests pass the reference's own draws."""
    if dual_edge:
        raise NotImplementedError("dual_edge=True is outside the B200 hot path")
    graph_node, graph_edge, graph_cell = graph_list[:3]
    cnT, fn, nc = _idx(graph_node, graph_cell)
    dev = graph_node.x.device
    on_cell, on_face = ops.interp_nodes(graph_node.x, 0, 6, cnT, graph_cell.cell_factor.contiguous(), fn)
    graph_cell.y = on_cell[:, 3:5]
    graph_edge.y = on_face[:, 3:6]
    C, F = on_cell.shape[0], on_face.shape[0]
    if cell_noise is None:
        cell_noise = torch.randn((C, 2), device=dev) * noise_std
    if flip_keep is None:
        flip_keep = randbool(F, device=dev)
    noised = on_cell[:, 0:2] + cell_noise
    graph_cell.x = torch.cat((graph_cell.x.to(torch.float32), noised), dim=1).contiguous()
    graph_cell.edge_attr, interior = ops.assemble_edge_attr(graph_cell.x[:, 2:4], graph_cell.pos.contiguous(), nc, graph_edge.x.contiguous(),
                                                            on_face[:, 0:2], flip_keep=flip_keep)
    return graph_node, graph_edge, graph_cell, interior, torch.logical_not(interior)


def create_next_graph(graph_old, graph_edge, mask_face_boundary, cell_next_UV, cell_type, Re, edge_RMP_EU, dual_edge=False):
    """Load_mesh.py:1076-1115: x <- [type, Re, next_UV]; edge_attr <- [uv_s-uv_r (boundary <- graph_edge.y[:,0,0:2]), RMP_EU]."""
    if dual_edge:
        raise NotImplementedError("dual_edge=True is outside the B200 hot path")
    C, F = cell_next_UV.shape[0], edge_RMP_EU.shape[0]
    x = torch.empty((C, 4), dtype=torch.float32, device=cell_next_UV.device)
    x[:, 0:1] = cell_type
    x[:, 1:2] = Re
    ea = torch.empty((F, 5), dtype=torch.float32, device=cell_next_UV.device)
    ea[:, 2:5] = edge_RMP_EU
    ops.create_next_graph_(x, ea, cell_next_UV.contiguous(), ops.as_i32(graph_old.edge_index), graph_edge.x.contiguous(), graph_edge.y[:, 0, 0:2])
    graph_old.x, graph_old.edge_attr = x, ea
    return graph_old


class RolloutState:
    """Allocation-free autoregressive stepping (the loop body of rollout.py:368-425) for CUDA-graph capture: keeps x[C,4],
    edge_attr[F,5] and the int32 tables resident and rewrites them in place each step."""

    def __init__(self, graph_cell, graph_edge):
        self.x0 = graph_cell.x.clone()
        self.ea0 = graph_cell.edge_attr.clone()
        self.nc = ops.as_i32(graph_cell.edge_index)
        self.ftl = graph_edge.x.contiguous()
        self.bc = graph_edge.y[:, 0, 0:2]

    def advance_(self, graph_cell, next_uv):
        x, ea = self.x0, self.ea0
        ops.create_next_graph_(x, ea, next_uv, self.nc, self.ftl, self.bc)
        graph_cell.x, graph_cell.edge_attr = x, ea
        return graph_cell
