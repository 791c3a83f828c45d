"""TEST / BASELINE INFRASTRUCTURE — drives the UNMODIFIED reference (oracle/ref_loader.py: /root/reference in the build
container, the byte-for-byte copy under oracle/_ref/ on the GPU box) through the two loop bodies the benchmark times:

  rollout step   rollout.py:368-425   FVGN.forward (eval) + Data_Pool.create_next_graph
  training step  train.py:150-225     Data_Pool.train_datapreprocessing + FVGN.forward (train) + compute_FVM_loss + backward + AdamW

on `device` = "cpu" (bench.py `--impl reference`, `cpu_baseline`) or "cuda" (bench.py `torch_eager_cuda`: the reference's own
PyTorch path on the B200 — what a user of the reference runs today).  The graphs come from the reference's own Data_Pool
(`require_minibatch_mesh` / `get` + PyG collation) over H5-layout groups whose tables are the oracle's (bit-identical to
extract_mesh_state's, tests/test_oracle_*).  Only bench.py and tests/ may import this file."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

from . import fvgn_oracle as O
from . import ref_loader as RL


def available():
    return RL.available()


_identity_cuda = lambda self, *a, **k: self  # noqa: E731


class _cpu_mode:
    """the reference calls `.cuda()` unconditionally (Load_mesh.py:788-790,1117-1125; loss_compute.py:40-75): while a CPU-device
    harness runs on a box that HAS a GPU, torch.Tensor.cuda is the identity (as ref_loader makes it on a CPU-only box)"""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        if self.on:
            self.saved = torch.Tensor.cuda
            torch.Tensor.cuda = _identity_cuda

    def __exit__(self, *a):
        if self.on:
            torch.Tensor.cuda = self.saved
        return False


class RefHarness:
    def __init__(self, meshes, params, device="cpu", seed=0, loss="global_mean_sum", lr=1e-3, prime=2):
        """meshes: list of synthetic mesh dicts (mesh_pos, cells, node_type, velocity[T,N,2], pressure[T,N,1])"""
        self.ref = RL.load()
        self.device = torch.device(device)
        self._cpu = _cpu_mode(self.device.type == "cpu")
        with self._cpu:
            self._init(meshes, params, seed, loss, lr, prime)

    def _init(self, meshes, params, seed, loss, lr, prime):
        self.params = params
        self.params.loss = loss
        self.params.train_traj_length = meshes[0]["velocity"].shape[0] - 1
        self.params.dataset_size = len(meshes)
        self.groups = []
        for m in meshes:
            tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
            tab["mesh_pos"], tab["node_type"] = m["mesh_pos"], m["node_type"]
            self.groups.append(RL.h5_group_from_tables(tab, m["velocity"], m["pressure"]))
        self.n_graphs = len(meshes)
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(seed)
            self.model = self.ref.FVGN(device=self.device, params=params).to(self.device)
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=lr)  # train.py:45
        self.loss_fn = torch.nn.MSELoss(reduction="mean")  # train.py:133
        self.rho, self.mu, self.dt = params.rho, params.mu, params.dt
        self._pool(False)
        self._rollout_init()
        self._train_init()
        # prime the normalisers like the reference's pre_statistics rounds (train.py:170-185): accumulating forwards, no backward
        self.model.train()
        with torch.no_grad():
            for _ in range(prime):
                self._forward(*self._train_batch()[:3])
        self.C = int(self.gc.x.shape[0])
        self.F = int(self.ge.x.shape[0])

    # ------------------------------------------------------------------ data
    def _pool(self, training):
        dp = RL.ref_data_pool(self.groups, self.params, is_training=training, device=self.device)
        dp._set_status(is_training=training)
        return dp

    def _rollout_init(self):
        """rollout.py:259: datasets.require_minibatch_mesh(...) -> the [N,T,3] rollout graphs on the device"""
        dp = self._pool(False)
        self.dp_valid = dp
        with contextlib.redirect_stdout(io.StringIO()):
            gn, ge, gc, mi, mb = dp.require_minibatch_mesh(start_epoch=0, batch_index=list(range(self.n_graphs)), is_training=False, mode="h5")
        gn, ge, gc = gn.to(self.device), ge.to(self.device), gc.to(self.device)
        self.gn, self.ge, self.gc, self.mb = gn, ge, gc, mb.to(self.device)
        self.cell_type, self.Re, self.rmp = gc.x[:, 0:1].clone(), gc.x[:, 1:2].clone(), gc.edge_attr[:, 2:5].clone()
        self.x_init, self.ea_init = gc.x.clone(), gc.edge_attr.clone()

    def _train_init(self):
        """what one iteration of train.py's DataLoader yields: [Batch(node), Batch(edge), Batch(cell)] of `get(idx)` samples"""
        dp = self._pool(True)
        self.dp_train = dp
        L = self.params.train_traj_length
        samples = [dp.get(i * L) for i in range(self.n_graphs)]
        self.train_cpu = [self.ref.Batch.from_data_list([s[k] for s in samples]) for k in range(3)]
        self.train_dev = [b.clone().to(self.device) for b in self.train_cpu]

    def _train_batch(self):
        """fresh bags over the resident batch (train_datapreprocessing rewrites graph_cell.x / .y / .edge_attr)"""
        graph_list = [b.clone() for b in self.train_dev]
        return self.dp_train.train_datapreprocessing(graph_list, dual_edge=False)

    # ------------------------------------------------------------------ steps
    def _forward(self, gn, ge, gc):
        p = self.params
        return self.model(graph=gc, graph_edge=ge, graph_node=gn, rho=self.rho, mu=self.mu, dt=self.dt, edge_one_hot=p.edge_one_hot,
                          cell_one_hot=p.cell_one_hot, device=self.device)

    def reset_rollout(self):
        self.gc.x, self.gc.edge_attr = self.x_init.clone(), self.ea_init.clone()

    def rollout_step(self):
        """rollout.py:373-425"""
        self.model.eval()
        with torch.no_grad(), self._cpu:
            uvp, nuv = self._forward(self.gn, self.ge, self.gc)
            self.ref.Data_Pool.create_next_graph(self.gc, self.ge, self.mb, nuv, self.cell_type, self.Re, self.rmp)
        return uvp, nuv

    def train_step(self):
        """train.py:150-225"""
        self.model.train()
        with self._cpu:
            self.opt.zero_grad()
            gn, ge, gc, mi, mb = self._train_batch()
            out = self._forward(gn, ge, gc)
            loss = self.ref.compute_FVM_loss(params=self.params, graph_cell=gc, graph_edge=ge, mask_face_interior=mi,
                                             predicted_edge_attr=out[3], target_face_uvp_normalized=out[4], predicted_delta_u=out[1],
                                             target_delta_u=out[2], loss_function=self.loss_fn)
            loss[0].backward()
            self.opt.step()
        return loss[0]
