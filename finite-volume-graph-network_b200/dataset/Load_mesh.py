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


def _idx(graph_node, graph_cell, graph_edge=None):
    """int32 (cells_node_T, face_node, neighbour_cell): the tables of the batch's cached MeshTopology when the edge bag is at hand (the
    forward that follows builds the same topology: the int64 -> int32 narrowing then happens once per batch), else converted here"""
    if graph_edge is not None and hasattr(graph_edge, "face"):
        from ..topology import topology_of

        topo = topology_of(graph_cell, graph_node, graph_edge)
        return topo.cnT, topo.fn, topo.nc
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
    cnT, fn, nc = _idx(graph_node, graph_cell, graph_edge)
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


# ------------------------------------------------------------------------------------------------------------------------------------
# Data_Pool: the reference's dataset object (dataset/Load_mesh.py:364-1135) over any h5py-like mapping with the key layout
# extract_mesh_state writes (Extract_mesh/parse_tfrecord_refactor.py:890-899, 962-965: static arrays [1, ...], fields [T, N, .]).
# Same constructor / method names, argument meaning and return values, so train.py:111-160, rollout.py:252-263, 416-425 and
# validate.py can be pointed at this package: `get` / `_build_graph` assemble the three PyG-style bags on the host,
# `require_minibatch_mesh` / `require_one_mesh` collate them into pinned, size-bucketed staging buffers (one H2D copy per tensor of the
# batch) and run the pre-step on the device, `train_datapreprocessing` / `valid_datapreprocessing` / `create_next_graph` are the kernels
# above behind the reference's signatures.
# ------------------------------------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402

from .mesh_store import MeshStore, open_store  # noqa: E402,F401


class PinnedBatcher:
    """Block-diagonal collation (PyG increment rule, data.Batch) of graph bags straight into PINNED host buffers whose capacities are
    rounded up to size buckets (powers of two of 4096 rows / sqrt(2) steps), so that consecutive minibatches of similar size reuse the
    same staging memory and the device copy is one cudaMemcpyAsync per tensor key."""

    def __init__(self, pin=None):
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.buffers = {}
        self.allocations = 0

    @staticmethod
    def bucket(n):
        b = 4096
        while b < n:
            b = int(b * 1.5) if (b & (b - 1)) == 0 else int(b / 1.5 * 2)  # 4096, 6144, 8192, 12288, 16384, ...
        return b

    def _buffer(self, name, rows, tail, dtype):
        key = (name, tuple(tail), dtype)
        buf = self.buffers.get(key)
        if buf is None or buf.shape[0] < rows:
            buf = torch.empty((self.bucket(rows),) + tuple(tail), dtype=dtype)
            if self.pin:
                buf = buf.pin_memory()
            self.buffers[key] = buf
            self.allocations += 1
        return buf[:rows]

    def collate(self, name, bags):
        """list of Data -> Batch whose tensors are views of the staging buffers (valid until the next collate of the same name)"""
        from ..data import _is_index

        out = Batch()
        keys = bags[0].keys
        sizes = [b.num_nodes for b in bags]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        for k in keys:
            vs = [getattr(b, k) for b in bags]
            if not torch.is_tensor(vs[0]):
                setattr(out, k, vs)
                continue
            if vs[0].dim() == 0:
                setattr(out, k, torch.stack(vs))
                continue
            if _is_index(k):  # [r, n_i] tables, concatenated on the last dim with the running node offset added
                r, tot = vs[0].shape[0], sum(v.shape[-1] for v in vs)
                buf = self._buffer(f"{name}.{k}", tot, (r,), vs[0].dtype)  # stored [tot, r] then viewed transposed: rows stay the bucketed dim
                flat = buf.t()  # [r, tot] view (strided)
                o = 0
                for v, off in zip(vs, offs[:-1]):
                    flat[:, o:o + v.shape[-1]] = torch.where(v >= 0, v + int(off), v) if r == 4 else v + int(off)
                    o += v.shape[-1]
                setattr(out, k, flat)
            else:
                tot = sum(v.shape[0] for v in vs)
                buf = self._buffer(f"{name}.{k}", tot, vs[0].shape[1:], vs[0].dtype)
                o = 0
                for v in vs:
                    buf[o:o + v.shape[0]] = v
                    o += v.shape[0]
                setattr(out, k, buf)
        bv = self._buffer(f"{name}.batch", int(offs[-1]), (), torch.long)
        for gi in range(len(bags)):
            bv[int(offs[gi]):int(offs[gi + 1])] = gi
        out.batch = bv
        out.ptr = torch.as_tensor(offs, dtype=torch.long)
        out.num_graphs = len(bags)
        return out


class Data_Pool(torch.utils.data.Dataset):
    """dataset/Load_mesh.py:364-1135.  `pool` maps str(case) -> group (h5py.File, MeshStore, or a dict of dicts of arrays)."""

    def __init__(self, params=None, is_traning=True, device=None):
        super().__init__()
        self.params = params
        self.is_training = is_traning
        self.device = self._device = device
        self.pool = {}
        self.mbatch_graph_node, self.mbatch_graph_cell, self.mbatch_graph_edge = [], [], []
        self.dataset_size = getattr(params, "dataset_size", 0)
        self.train_epoch = 0
        self.epoch = 0
        self.start_epoch = 0
        L = int(getattr(params, "train_traj_length", 1))
        self._time_step_list = np.tile(np.random.permutation(L).astype(np.int64)[None], (max(int(self.dataset_size), 1), 1))  # :387-394
        self._batcher = PinnedBatcher()

    # -- bookkeeping setters of the reference (:399-440)
    def _set_epoch(self, epoch):
        self.epoch = epoch

    def _set_fetch_time_steps(self, epoch):
        self.train_epoch = epoch

    def _set_status(self, is_training=False):
        self.is_training = is_training

    def _set_time_step_list(self, new_time_step_list):
        self._time_step_list = new_time_step_list

    def _set_dataset(self, state="simple"):
        if state != "full":
            raise NotImplementedError("the simple / complex pools exist for the TFRecord reader only (Load_mesh.py:700-735), which is out of scope")
        return True

    def randbool(self, *size):
        """50 % True (:442-446); drawn on the device"""
        return randbool(*size, device=self.device)

    def random_time_steps(self, data_idx):
        return self._time_step_list[data_idx][self.train_epoch % self.params.train_traj_length]

    # -- torch Dataset protocol (the reference subclasses torch_geometric.data.Dataset: len() / get())
    def len(self):
        return len(self.pool) * self.params.train_traj_length

    def __len__(self):
        return self.len()

    def __getitem__(self, idx):
        return self.get(int(idx))

    def _group(self, case_id):
        if getattr(self.params, "dataset_type", "h5") != "h5":
            raise ValueError("wrong dataset type")  # (the reference's 'tf' pools come from the TFRecord iterator: out of scope)
        return self.pool[str(case_id)]

    @staticmethod
    def _reynolds(g, n_cells):
        """:489-503 — `reynolds_num` if stored, else mean_u * charac_scale / 0.001"""
        try:
            return torch.as_tensor(np.asarray(g["reynolds_num"][()])).view(-1, 1).repeat(n_cells, 1)
        except Exception:
            mean_u = torch.as_tensor(np.asarray(g["mean_u"][()])).view(-1, 1).repeat(n_cells, 1)
            R = torch.as_tensor(np.asarray(g["charac_scale"])).reshape(-1)[0:1].repeat(n_cells, 1)
            return (mean_u * 1.0 * R / 0.001).to(torch.float32)

    def _static(self, g):
        t = lambda k, dt: torch.as_tensor(np.asarray(g[k][0])).to(dt)
        return dict(mesh_pos=torch.as_tensor(np.asarray(g["mesh_pos"][0])), face_node=t("face", torch.long), cells_node=t("cells_node", torch.long),
                    node_type=torch.as_tensor(np.asarray(g["node_type"][0])), nc=t("neighbour_cell", torch.long), cell_area=t("cells_area", torch.float32),
                    centroid=t("centroid", torch.float32), cell_factor=t("cell_factor", torch.float32), cell_face=t("cells_face", torch.long),
                    cells_type=t("cells_type", torch.long), face_length=t("face_length", torch.float32), face_types=t("face_type", torch.long),
                    unv=t("unit_norm_v", torch.float32))

    def get(self, idx):
        """one training sample (:467-587): graph_node.x [N,6] = [u,v,p](t) ++ [u,v,p](t+1)"""
        L = self.params.train_traj_length
        case_id, t0 = int(idx / L), idx % L
        g = self._group(case_id)
        s = self._static(g)
        vel = torch.as_tensor(np.asarray(g["target|velocity_on_node"][t0:t0 + 2]))
        prs = torch.as_tensor(np.asarray(g["target|pressure_on_node"][t0:t0 + 2]))
        target_on_node = torch.cat(torch.unbind(torch.cat((vel, prs), dim=2)), dim=1)
        cell_attr = torch.cat((s["cells_type"], torch.zeros_like(s["cells_type"])), dim=1)
        kw = {} if self.is_training else dict(node_type=s["node_type"])
        graph_node = Data(x=target_on_node, edge_index=s["face_node"], face=s["cells_node"].T, pos=s["mesh_pos"], **kw)
        graph_edge = Data(x=torch.cat((s["face_types"], s["face_length"]), dim=1), face=s["cell_face"].T, y=s["face_types"])
        graph_cell = Data(x=cell_attr, edge_index=s["nc"], unv=s["unv"], centroid=s["centroid"], cell_area=s["cell_area"],
                          cell_factor=s["cell_factor"], pos=s["centroid"], y=self._reynolds(g, cell_attr.shape[0]), graph_index=torch.as_tensor(idx))
        return [graph_node, graph_edge, graph_cell]

    def _build_graph(self, minibatch_data, graph_indices):
        """one whole trajectory for rollouts (:589-690): graph_node.x [N,T,3]"""
        g = self._group(graph_indices)
        s = self._static(g)
        target_on_node = torch.cat((torch.as_tensor(np.asarray(g["target|velocity_on_node"][:])), torch.as_tensor(np.asarray(g["target|pressure_on_node"][:]))),
                                   dim=2).transpose(0, 1)
        cell_attr = torch.cat((s["cells_type"], torch.zeros_like(s["cells_type"])), dim=1)
        mean_u = torch.as_tensor(np.asarray(g["mean_u"][()])).view(-1, 1).repeat(cell_attr.shape[0], 1)
        graph_node = Data(x=target_on_node, edge_index=s["face_node"], face=s["cells_node"].T, pos=s["mesh_pos"], node_type=s["node_type"])
        graph_edge = Data(x=torch.cat((s["face_types"], s["face_length"]), dim=1), face=s["cell_face"].T, y=s["face_types"])
        graph_cell = Data(x=cell_attr, edge_index=s["nc"], unv=s["unv"], centroid=s["centroid"], cell_area=s["cell_area"], cell_factor=s["cell_factor"],
                          pos=s["centroid"], y=cell_attr, graph_index=torch.as_tensor(graph_indices), mean_u=mean_u)
        self.mbatch_graph_node.append(graph_node)
        self.mbatch_graph_cell.append(graph_cell)
        self.mbatch_graph_edge.append(graph_edge)

    def reset(self):
        self.mbatch_graph_node.clear()
        self.mbatch_graph_cell.clear()
        self.mbatch_graph_edge.clear()

    def collate(self, batch_list):
        return batch_list

    def load_mesh_to_cpu(self, mode="h5", split="train", dataset_dir=None):
        """:700-760, 'h5' mode: opens <dir>/<split>.h5 (h5py when installed) or <dir>/<split>.fvgn.npz (MeshStore: the same key layout in a
        numpy archive, written by fvgn_b200.convert with this package's own tools)"""
        if mode != "h5":
            raise NotImplementedError("the TFRecord ('tf') iterator is out of scope: convert with fvgn_b200.convert and load mode='h5'")
        d = dataset_dir if dataset_dir is not None else self.params.dataset_dir_h5
        self.valid_pool = []
        self.pool = open_store(d, split)
        return d

    # -- minibatches for rollouts / validation (:762-814)
    def _to_device_batches(self):
        dev = self.device
        out = []
        for name, bags in (("node", self.mbatch_graph_node), ("edge", self.mbatch_graph_edge), ("cell", self.mbatch_graph_cell)):
            b = self._batcher.collate(name, bags)
            out.append(b.to(dev, non_blocking=True) if dev is not None and str(dev) != "cpu" else b.clone())
        if torch.cuda.is_available() and dev is not None and str(dev) != "cpu":
            torch.cuda.current_stream().synchronize()  # the pinned staging buffers are rewritten by the next minibatch
        return out

    def require_minibatch_mesh(self, start_epoch=None, batch_index=None, is_training=False, dual_edge=False, mode="tf"):
        self.is_training = is_training
        self.start_epoch = start_epoch % 599
        for i in batch_index:
            self._build_graph(None, i)
        gn, ge, gc = self._to_device_batches()
        return self.valid_datapreprocessing(gn, ge, gc, cell_noise=None, dual_edge=dual_edge)

    def require_one_mesh(self, start_epoch, rollout_index, is_training=False, dual_edge=False):
        self.is_training = is_training
        self.start_epoch = start_epoch % 599
        self._build_graph(None, rollout_index)
        gn, ge, gc = self._to_device_batches()
        return self.valid_datapreprocessing(gn, ge, gc, cell_noise=None, dual_edge=dual_edge)

    # -- the device pre / post step behind the reference's signatures (:904-1115)
    def valid_datapreprocessing(self, graph_node, graph_edge, graph_cell, cell_noise, dual_edge=False):
        return valid_datapreprocessing(graph_node, graph_edge, graph_cell, start_epoch=self.start_epoch, dual_edge=dual_edge)

    def cudacopy(self, data):
        dev = self.device if (self.device is not None and str(self.device) != "cpu") else "cuda"
        if isinstance(data, (tuple, list)):
            return [xi.to(dev, non_blocking=True) for xi in data]
        return data.to(dev, non_blocking=True)

    def get_noise(self, graph_list):
        """utils/noise.py:13-26 (two N(0, std) columns per cell); drawn on the device"""
        gc = graph_list[2]
        std = float(getattr(self.params, "Noise_injection_factor", 0.02))
        graph_list.append(torch.randn((gc.x.shape[0], 2), device=gc.x.device) * std)
        return graph_list

    def train_datapreprocessing(self, graph_list, dual_edge=False):
        graph_list = self.cudacopy(list(graph_list))
        graph_node, graph_edge, graph_cell, cell_noise = self.get_noise(graph_list)
        flip_keep = self.randbool(graph_cell.edge_index.shape[1]) if not dual_edge else None
        gn, ge, gc, mi, mb = train_datapreprocessing([graph_node, graph_edge, graph_cell], cell_noise=cell_noise, flip_keep=flip_keep,
                                                     dual_edge=dual_edge)
        if not self.is_training:  # (:1005-1008: the evaluation flavour also carries the pressure target on cells)
            raise NotImplementedError("train_datapreprocessing with is_training=False is never reached by the reference's scripts")
        return gn, ge, gc, mi, mb

    create_next_graph = staticmethod(create_next_graph)


def make_data_loader(data_pool=None, batch_size=None, num_workers=0, prefetch_factor=None, pin_memory=False, shuffle=False):
    """what train.py:119-125 builds with torch_geometric.loader.DataLoader: batches of `get()` samples collated into the three Batches"""
    kw = dict(prefetch_factor=prefetch_factor, persistent_workers=True) if num_workers > 0 else {}
    return torch.utils.data.DataLoader(data_pool, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=pin_memory,
                                       collate_fn=collate, **kw)


DataLoader = make_data_loader
