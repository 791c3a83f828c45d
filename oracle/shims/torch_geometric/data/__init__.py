"""Attribute-bag `Data`/`Dataset` with the two PyG behaviours the reference relies on:
`keys` is a PROPERTY (utils/utilities.py:26,41) and `num_nodes == x.shape[0]` (GN_blocks.py:121,137)."""
import copy


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

    def clone(self):
        out = self.__class__()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if hasattr(v, "clone") else copy.deepcopy(v)
        return out

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self


class Dataset:
    def __init__(self, root=None, transform=None, pre_transform=None):
        pass
