"""Block-diagonal batching with PyG's increment rule: keys containing 'index' or equal to 'face'
are offset by the running num_nodes and concatenated on the last dim; everything else on dim 0."""
import torch
from . import Data


class Batch(Data):
    @classmethod
    def from_data_list(cls, lst):
        out = cls()
        keys = lst[0].keys
        offset = 0
        cols = {k: [] for k in keys}
        batch = []
        for gi, d in enumerate(lst):
            n = d.num_nodes
            for k in keys:
                v = getattr(d, k)
                if torch.is_tensor(v) and ("index" in k or k == "face") and v.dim() >= 1 and v.dtype in (torch.int64, torch.int32):
                    if k == "graph_index":
                        cols[k].append(v.reshape(1))
                    else:
                        cols[k].append(v + offset)
                else:
                    cols[k].append(v)
            batch.append(torch.full((n,), gi, dtype=torch.long))
            offset += n
        for k in keys:
            vs = cols[k]
            if torch.is_tensor(vs[0]):
                if ("index" in k or k == "face") and k != "graph_index":
                    setattr(out, k, torch.cat(vs, dim=-1))
                elif vs[0].dim() == 0:
                    setattr(out, k, torch.stack(vs))
                else:
                    setattr(out, k, torch.cat(vs, dim=0))
            else:
                setattr(out, k, vs)
        out.batch = torch.cat(batch)
        out.num_graphs = len(lst)
        out._slices = [d.num_nodes for d in lst]
        return out

    def to_data_list(self):
        raise NotImplementedError
