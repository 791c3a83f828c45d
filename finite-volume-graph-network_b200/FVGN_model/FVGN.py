"""FVGN top-level module with the reference's constructor / forward / checkpoint interface (FVGN_model/FVGN.py:19-402).

forward(graph, graph_edge, graph_node, rho, mu, dt, edge_one_hot, cell_one_hot, device)
  train: (loss_continuity, predicted_delta_u[C,2], target_delta_u[C,2], predicted_edge_attr[F,5], target_face_uvp_normalized[F,3])
  eval : (predicted_uvp_on_edge[F,3], next_UV_on_cell[C,2])
Like the reference it overwrites graph.x / graph.edge_attr with the normalised features (App. D.8).

Every arithmetic step runs in libfvgn_b200 (normaliser kernels -> fused encoder -> 15 x (node segment-sum, fused cell
kernel, fused edge kernel) -> fused decoder -> face-area BatchNorm -> flux integrator -> de-normalise)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops, parallel
from ..topology import topology_of
from ..utils import normalization
from .EncoderProcesserDecoder import EncoderProcesserDecoder, Intergrator


class FVGN(nn.Module):
    def __init__(self, device=None, model_dir=None, params=None) -> None:
        super().__init__()
        self._device = device
        self.edge_input_size = params.edge_input_size + params.edge_one_hot
        self.cell_input_size = params.cell_input_size + params.cell_one_hot
        self.model_dir = model_dir
        self.dual_edge = params.dual_edge
        self.params = params
        self.model = EncoderProcesserDecoder(
            message_passing_num=params.message_passing_num, cell_input_size=self.cell_input_size, edge_input_size=self.edge_input_size,
            cell_output_size=params.cell_output_size, edge_output_size=params.edge_output_size, drop_out=params.drop_out,
            mp_times=params.mp_times, MultiHead=params.multihead, dual_edge=params.dual_edge, nn_act=params.nn_act).to(device)
        self.intergator = Intergrator(self.edge_input_size, self.cell_input_size, params.cell_output_size, params.loss_cont)
        big = (int(params.dataset_size / params.batch_size) + 1) * params.cumulative_length
        small = (int(1000 / params.batch_size) + 1) * params.cumulative_length
        self._acc_output_normalizer = normalization.Normalizer(size=params.cell_target_input_size, max_accumulations=big, device=device)
        self._face_uvp_flux_output_normalizer = normalization.Normalizer(size=params.face_flux_target_input_size, max_accumulations=small, device=device)
        self._edge_normalizer = normalization.Normalizer(size=params.edge_input_size + params.edge_one_hot, max_accumulations=big, device=device)
        self._cell_normalizer = normalization.Normalizer(size=params.cell_input_size + params.cell_one_hot, max_accumulations=big, device=device)
        self._grid_feature_normalizer = nn.BatchNorm1d(num_features=params.grid_feature_size)
        # default arithmetic of the 128-wide MLPs: the fp32-equivalent tensor-core mode (held to the reference's fp32 bound, rel 1e-5 per
        # step; ~3.7x the strict FFMA path).  set_math_mode("fp32") selects the strict FFMA kernels, "bf16" plain bf16 tiles.
        self.math_mode = "fp32"
        self.set_math_mode("bf16x3")

    def set_math_mode(self, mode):
        """'fp32'   strict fp32 on the FFMA pipes (the reference's evaluation order per row);
        'bf16x3' (default) fp32-equivalent on tcgen05: forward = two-term fp16 split of both operands, 3 MMAs per product (rel ~1e-6 per
                 step, held to the 1e-5 bound; ~5x the FFMA path), backward = the three leading products of 3-way bf16 splits.  Range: the
                 fp16 terms overflow for |activation| >= ~8190 — such a row turns NaN and sets the device status word, which
                 `check_finite()` / GraphedRollout.run(check=True) / GraphedTrainStep.check() turn into a FloatingPointError;
        'bf16'   plain bf16 tiles with fp32 accumulation (stated rollout-error bound, tests/test_gpu_rollout600.py)."""
        self.model.set_math_mode(mode)
        self.math_mode = mode
        return self

    def check_finite(self):
        """synchronises; raises FloatingPointError if a fp32-equivalent forward tile produced a non-finite row since the last check"""
        ops.raise_if_nonfinite("forward")

    # -- feature assembly (FVGN.py:83-138): one-hot + normalise happen in one kernel, nothing is concatenated in HBM
    def update_cell_attr(self, frames, one_hot: int, types: torch.Tensor, accumulation=True, normalizer=None):
        if normalizer != "1":
            raise NotImplementedError("only the normalised path (normalizer='1') is used by the reference forward")
        return self._cell_normalizer(frames, accumulation, type_col=types if one_hot > 0 else None, one_hot=one_hot)

    def update_edge_attr(self, edge_attr, one_hot: int, types: torch.Tensor, accumulation=True, normalizer=None, dual_edge=False):
        if normalizer != "1" or dual_edge:
            raise NotImplementedError("only normalizer='1', dual_edge=False is on the B200 hot path")
        return self._edge_normalizer(edge_attr, accumulation, type_col=types if one_hot > 0 else None, one_hot=one_hot)

    def velocity_to_accelation(self, noised_frames, next_cell_attr):
        return next_cell_attr - noised_frames

    def _face_area(self, graph, graph_edge, topo, dt, want_raw=False):
        bn = self._grid_feature_normalizer
        use_batch_stats = self.training or not bn.track_running_stats
        if use_batch_stats and parallel.is_distributed():
            # data parallel (also under torch.no_grad(): the reference's pre_statistics rounds, train.py:170-185): batch statistics and
            # running-stat updates are those of the WHOLE batch, as in autograd.NetFunction — the ranks' BatchNorm buffers stay identical
            red = self.__dict__.get("_dp_bn_reduced")
            sums = None if red is not None else ops.face_area_sums(graph_edge.x, graph.cell_area, topo.nc, dt)[1]
            gstats, mean, var, n = parallel.global_bn_stats(sums, topo.F, reduced=red)
            self.__dict__["_dp_bn_reduced"] = None
            parallel.update_running_stats_(bn, mean, var, n)
            out, raw, stats = ops.face_area_bn_fwd(graph_edge.x, graph.cell_area, topo.nc, dt, True, bn.weight.detach(), bn.bias.detach(),
                                                   bn.running_mean, bn.running_var, want_raw=want_raw, stats=gstats)
        else:
            out, raw, stats = ops.face_area_bn_fwd(graph_edge.x, graph.cell_area, topo.nc, dt, use_batch_stats, bn.weight, bn.bias,
                                                   bn.running_mean, bn.running_var, want_raw=want_raw)
        if self.training:
            bn.num_batches_tracked += 1
        return out, raw, stats

    def forward(self, graph=None, graph_edge=None, graph_node=None, rho=None, mu=None, dt=None, edge_one_hot=9, cell_one_hot=9,
                device=None, capture=None):
        topo = topology_of(graph, graph_node, graph_edge)
        cells_type = graph.x[:, 0:1]
        faces_type = graph_edge.x[:, 0:1]
        uv = graph.x[:, 2:4]
        norm_before = parallel.snapshot_normalizers(self) if (self.training and parallel.is_distributed()) else None
        self.__dict__["_dp_bn_reduced"] = None
        if norm_before is not None:
            # DP: every rank must normalise with the statistics of the WHOLE batch: accumulate locally (no apply yet),
            # all-reduce the increments — the face-area BatchNorm's batch statistics (a function of the geometry only) ride along in the
            # same collective: one statistic all-reduce per step —, then apply once
            target_acc = self.velocity_to_accelation(uv, graph.y[:, 0:2])
            self._acc_output_normalizer._accumulate(target_acc)
            self._face_uvp_flux_output_normalizer._accumulate(graph_edge.y[:, 0:3])
            self._cell_normalizer._accumulate(uv, cells_type if cell_one_hot > 0 else None, cell_one_hot)
            self._edge_normalizer._accumulate(graph.edge_attr, faces_type if edge_one_hot > 0 else None, edge_one_hot)
            _, bn_sums = ops.face_area_sums(graph_edge.x, graph.cell_area, topo.nc, dt)
            extra = torch.cat([bn_sums, torch.full((1,), float(topo.F), dtype=torch.float64, device=bn_sums.device)])
            self.__dict__["_dp_bn_reduced"] = parallel.sync_normalizer_increments_(self, norm_before, extra=extra)
            target_delta_u = self._acc_output_normalizer(target_acc, False)
            target_face = self._face_uvp_flux_output_normalizer(graph_edge.y[:, 0:3], False)
            x_in = self.update_cell_attr(uv, cell_one_hot, cells_type, False, normalizer="1")
            e_in = self.update_edge_attr(graph.edge_attr, edge_one_hot, faces_type, False, normalizer="1", dual_edge=self.dual_edge)
        else:
            if self.training:
                target_acc = self.velocity_to_accelation(uv, graph.y[:, 0:2])
                target_delta_u = self._acc_output_normalizer(target_acc, True)
                target_face = self._face_uvp_flux_output_normalizer(graph_edge.y[:, 0:3], True)
            x_in = self.update_cell_attr(uv, cell_one_hot, cells_type, self.training, normalizer="1")
            e_in = self.update_edge_attr(graph.edge_attr, edge_one_hot, faces_type, self.training, normalizer="1", dual_edge=self.dual_edge)
        graph.x, graph.edge_attr = x_in, e_in
        if capture is not None:
            capture["cell_in"], capture["edge_in"] = x_in, e_in
        from ..autograd import NetFunction, network_mlps, network_params

        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in network_params(self)):
            # differentiable path: forward and backward both run in libfvgn_b200 kernels (autograd.NetFunction)
            if self.math_mode not in ("fp32", "bf16x3", "bf16"):
                raise NotImplementedError("training modes: 'fp32' (FFMA), 'bf16x3' (fp32-equivalent tensor cores), 'bf16' (bf16 tiles)")
            if self.math_mode != "fp32":  # the tensor-core forward reads packed weight tiles: one launch re-packs every MLP an optimizer step changed
                ops.ensure_packed_many([m.mlp_ref() for m in network_mlps(self)], 2)

            decoded, delta_u, face_area = NetFunction.apply(self, topo, x_in, e_in, graph_edge.x.contiguous(), graph.cell_area.contiguous(),
                                                            graph.unv.contiguous(), rho, dt, *network_params(self))
            loss_cont = 0
        else:
            _, decoded = self.model(graph, graph_node, topo=topo, capture=capture)
            face_area, _, _ = self._face_area(graph, graph_edge, topo, dt)
            loss_cont, delta_u = self.intergator(decoded=decoded, unv=graph.unv, rho=rho, rhs_coef=1.0, face_area=face_area, cells_face_T=topo.cfT)
        if capture is not None:
            capture["decoded"], capture["face_area"], capture["delta_u"] = decoded, face_area, delta_u
        if self.training:
            return loss_cont, delta_u, target_delta_u, decoded, target_face
        next_uv = self._acc_output_normalizer.inverse(delta_u, addend=uv)
        uvp_face = self._face_uvp_flux_output_normalizer.inverse(decoded, width=3)
        return uvp_face, next_uv

    # -- checkpoints: same dict layout as the reference (FVGN.py:328-402)
    def load_checkpoint(self, optimizer=None, scheduler=None, ckpdir=None, device=None, is_traning=True, trian_time_steps=None):
        if ckpdir is None:
            ckpdir = self.model_dir
        dicts = torch.load(ckpdir, map_location=device, weights_only=False)
        self.load_state_dict(dicts["model"])
        train_time_steps = dicts["trian_time_steps"]
        if optimizer is not None:
            for i, o in enumerate(optimizer if isinstance(optimizer, list) else [optimizer]):
                o.load_state_dict(dicts["optimizer{}".format(i)])
        if scheduler is not None:
            for i, s in enumerate(scheduler if isinstance(scheduler, list) else [scheduler]):
                state = dicts["scheduler{}".format(i)]
                if hasattr(s, "load_state_dict"):
                    s.load_state_dict(state)
                for key, value in state.items():
                    if torch.is_tensor(value):
                        value = value.to(device)
                    setattr(s, key, value)
        print("FVGN model and optimizer/scheduler loaded checkpoint %s" % ckpdir)
        return train_time_steps

    def save_checkpoint(self, path=None, optimizer=None, scheduler=None, trian_time_steps: np.ndarray = None, index="final"):
        if path is None:
            path = self.model_dir
        to_save = {"model": self.state_dict(), "trian_time_steps": trian_time_steps}
        for i, o in enumerate(optimizer if isinstance(optimizer, list) else [optimizer]):
            if o is not None:
                to_save["optimizer{}".format(i)] = o.state_dict()
        for i, s in enumerate(scheduler if isinstance(scheduler, list) else [scheduler]):
            if s is not None:
                to_save["scheduler{}".format(i)] = s.get_variable() if hasattr(s, "get_variable") else s.state_dict()
        torch.save(to_save, path)
        print("FVGN model saved at %s" % path)
