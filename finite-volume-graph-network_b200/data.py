"""Attribute bags with the torch_geometric.data.Data / Batch behaviour the reference relies on (PyG is not a
dependency of this package): `keys` is a property, `num_nodes == x.shape[0]`, block-diagonal batching offsets keys
containing 'index' or equal to 'face' (SURVEY.md §8b "Graph objects")."""
import torch


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            if v is not None:
                setattr(self, k, v)

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    @property
    def num_nodes(self):
        return self.x.shape[0]

    def _apply(self, fn):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = fn(v)
        return self

    def to(self, device, non_blocking=False):
        return self._apply(lambda t: t.to(device, non_blocking=non_blocking))

    def cuda(self, device=None, non_blocking=False):
        return self.to(torch.device("cuda", torch.cuda.current_device()) if device is None else device, non_blocking)

    def pin_memory(self):
        return self._apply(lambda t: t.pin_memory())

    def clone(self):
        out = self.__class__()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else v
        return out


class Batch(Data):
    @classmethod
    def from_data_list(cls, lst):
        """PyG's block-diagonal collation: keys containing 'index' or equal to 'face' are offset by the running num_nodes and concatenated
        on the last dim, everything else on dim 0.  Vectorised: one cat + one offset add per key (not one add per graph: with a new batch every
        training step, train.py:119-160, this runs on the step's critical host path)."""
        out = cls()
        G = len(lst)
        dev = lst[0].x.device
        sizes = [g.num_nodes for g in lst]
        ptr = [0]
        for n in sizes:
            ptr.append(ptr[-1] + n)
        iptr = {}
        for k in lst[0].keys:
            vs = [getattr(g, k) for g in lst]
            v0 = vs[0]
            if not torch.is_tensor(v0):
                setattr(out, k, vs)
            elif v0.dim() == 0:
                setattr(out, k, torch.stack(vs))
            elif _is_index(k):
                ip = [0]
                for v in vs:
                    ip.append(ip[-1] + int(v.shape[-1]))
                iptr[k] = ip
                if v0.shape[0] == 4:  # (entries < 0 are the padding of [4, C] tri / quad tables and stay as they are)
                    shifted = [torch.where(v >= 0, v + o, v) for v, o in zip(vs, ptr[:-1])]
                else:                 # one multi-tensor launch for all graphs (no host -> device traffic, nothing synchronises)
                    shifted = torch._foreach_add(vs, ptr[:-1])
                setattr(out, k, torch.cat(shifted, dim=-1))
            else:
                setattr(out, k, torch.cat(vs, dim=0))
        out.batch = torch.zeros(ptr[-1], dtype=torch.long, device=dev)
        torch._foreach_add_(list(out.batch.split(sizes)), list(range(G)))
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        out.num_graphs = G
        out._index_ptr = iptr
        return out

    def to_data_list(self):
        """inverse of from_data_list for tensors whose leading (or, for index tables, last) dimension is the batched one
        (rollout.py:271-273 splits the minibatch back into per-mesh graphs)"""
        ptr = [int(v) for v in self.ptr]
        out = []
        for gi in range(self.num_graphs):
            a, b = ptr[gi], ptr[gi + 1]
            d = Data()
            for k, v in self.__dict__.items():
                if k in ("batch", "ptr", "num_graphs") or k.startswith("_"):
                    continue
                if not torch.is_tensor(v):
                    d.__dict__[k] = v[gi] if isinstance(v, (list, tuple)) and len(v) == self.num_graphs else v
                elif _is_index(k):
                    # tables indexed by another bag's rows (e.g. graph_edge.face: per CELL) cannot be split by this bag's ptr: select
                    # the columns whose entries fall into this graph's own node range of the table's TARGET — done by the caller
                    cols = self.__dict__.get("_index_ptr", {}).get(k)
                    if cols is None:
                        d.__dict__[k] = v
                    else:
                        d.__dict__[k] = v[:, cols[gi]:cols[gi + 1]] - a
                elif v.dim() > 0 and v.shape[0] == ptr[-1]:
                    d.__dict__[k] = v[a:b]
                elif v.dim() > 0 and v.shape[0] == self.num_graphs:
                    d.__dict__[k] = v[gi]
                else:
                    d.__dict__[k] = v
            out.append(d)
        return out


def _is_index(k):
    return ("index" in k and k != "graph_index") or k == "face"
