"""Online Normalizer with the reference's interface and buffers (utils/normalization.py:5-86): `acc_count`,
`num_accumulations`, `acc_sum`, `acc_sum_squared`, all initialised to ONE.  Accumulation, the
mean/std derivation and the (x-mean)/std map run in libfvgn_b200 kernels; no host synchronisation (the
`num_accumulations < max_accumulations` test happens on the device)."""
import torch
import torch.nn as nn

from .. import ops


class Normalizer(nn.Module):
    def __init__(self, size, max_accumulations=10 ** 7, epsilon=1e-8, device=None):
        super().__init__()
        self.max_accumulations = max_accumulations
        self.epsilon = epsilon  # the kernels use the reference default 1e-8
        self.size = size
        self.register_buffer("acc_count", torch.tensor(1.0, dtype=torch.float32, device=device))
        self.register_buffer("num_accumulations", torch.tensor(1.0, dtype=torch.float32, device=device))
        self.register_buffer("acc_sum", torch.ones(size, dtype=torch.float32, device=device))
        self.register_buffer("acc_sum_squared", torch.ones(size, dtype=torch.float32, device=device))

    def forward(self, batched_data, accumulate=True, type_col=None, one_hot=0):
        """(batched_data - mean) / std.  With one_hot>0 the features are [batched_data, one_hot(type_col)] assembled on the
        fly (FVGN.update_edge_attr, FVGN.py:110-138)."""
        if accumulate:
            self._accumulate(batched_data, type_col, one_hot)
        return ops.normalizer_apply(batched_data, self.acc_count, self.acc_sum, self.acc_sum_squared, inverse=False,
                                    type_col=type_col, one_hot=one_hot)

    def inverse(self, normalized_batch_data, addend=None, width=None):
        return ops.normalizer_apply(normalized_batch_data, self.acc_count, self.acc_sum, self.acc_sum_squared, inverse=True,
                                    addend=addend, width=width)

    def _accumulate(self, batched_data, type_col=None, one_hot=0):
        ops.normalizer_accumulate(batched_data, self.acc_count, self.num_accumulations, self.acc_sum, self.acc_sum_squared,
                                  self.max_accumulations, type_col=type_col, one_hot=one_hot)

    def _mean(self):
        return self.acc_sum / torch.clamp(self.acc_count, min=1.0)

    def _std(self):
        std = torch.sqrt(self.acc_sum_squared / torch.clamp(self.acc_count, min=1.0) - self._mean() ** 2)
        std[std < self.epsilon] = 1.0
        return std
