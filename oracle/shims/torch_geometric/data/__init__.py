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

    def _move(self, device):
        import torch

        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self

    def cuda(self, device=None, *a, **k):
        """follows torch.Tensor.cuda: identity where that is patched to the identity (CPU-only box, or the CPU-mode reference harness on
        a GPU box: oracle/ref_loader.py, oracle/ref_harness.py), a real move otherwise (bench.py's torch-eager CUDA leg)"""
        import torch

        probe = torch.empty(0).cuda()
        if not probe.is_cuda:
            return self
        return self._move(probe.device if (device is None or not str(device).startswith("cuda")) else device)

    def to(self, device=None, *a, **k):
        import torch

        if device is None or (str(device).startswith("cuda") and not torch.cuda.is_available()):
            return self
        return self._move(device)


class Dataset:
    def __init__(self, root=None, transform=None, pre_transform=None):
        pass
