"""CPU, world_size 2, gloo: host-side logic of the data-parallel path (finite-volume-graph-network_b200/parallel.py):
flat gradient bucket all-reduce, online-normaliser increment sync and cross-rank BatchNorm statistics reproduce the
single-process batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fvgn_b200 as fv
        from fvgn_b200 import parallel
        from tests import helpers as H

        torch.manual_seed(0)
        model = fv.FVGN(device="cpu", params=H.params())
        # ---- 1. gradient bucket: mean over ranks of rank-dependent gradients
        for i, p in enumerate(model.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (1 + (i % 3)))
        bucket = parallel.GradBucket(model.parameters())
        assert bucket.numel == 2210695
        bucket.all_reduce_mean_()
        for i, p in enumerate(model.parameters()):
            assert torch.allclose(p.grad, torch.full_like(p, 1.5 * (1 + (i % 3))))
        # ---- 2. normaliser increments: each rank accumulates its half of a batch (host emulation of the device kernel)
        gen = torch.Generator().manual_seed(5)
        full = torch.randn(40, 14, generator=gen) * 2 + 0.5
        mine = full[rank * 20:(rank + 1) * 20]
        before = parallel.snapshot_normalizers(model)
        n = model._edge_normalizer
        n.acc_sum += mine.sum(0)
        n.acc_sum_squared += (mine ** 2).sum(0)
        n.acc_count += 20.0
        n.num_accumulations += 1.0
        parallel.sync_normalizer_increments_(model, before)
        assert torch.allclose(n.acc_sum, 1.0 + full.sum(0), rtol=1e-6)
        assert torch.allclose(n.acc_sum_squared, 1.0 + (full ** 2).sum(0), rtol=1e-6)
        assert float(n.acc_count) == 41.0 and float(n.num_accumulations) == 2.0
        assert float(model._cell_normalizer.acc_count) == 1.0  # untouched normalisers stay untouched
        # ---- 3. BatchNorm statistics over both ranks == statistics of the concatenated batch
        raw = torch.rand(30, generator=gen, dtype=torch.float64) + 0.1
        part = raw[rank * 15:(rank + 1) * 15]
        stats, mean, var, cnt = parallel.global_bn_stats(torch.stack([part.sum(), (part ** 2).sum()]), 15)
        assert abs(float(mean) - float(raw.mean())) < 1e-12 and abs(float(var) - float(raw.var(unbiased=False))) < 1e-12 and float(cnt) == 30.0
        bn = torch.nn.BatchNorm1d(1)
        ref = torch.nn.BatchNorm1d(1)
        ref.train()
        ref(raw.float().view(-1, 1))
        parallel.update_running_stats_(bn, mean, var, cnt)
        assert torch.allclose(bn.running_mean, ref.running_mean, rtol=1e-5) and torch.allclose(bn.running_var, ref.running_var, rtol=1e-5)
        # ---- 4. direct_mean_loss: the numerators / denominators of the batch means are summed over the ranks BEFORE the log
        #         (utils/loss_compute.py:92-113); host emulation of what ops.loss_fwd(reduce_partials=parallel.all_reduce_sum_) does
        err = torch.rand(40, 2, generator=gen, dtype=torch.float64) * (1.0 + 3.0 * (torch.arange(40) >= 20).double().view(-1, 1))
        mine = err[rank * 20:(rank + 1) * 20]
        tot = torch.tensor([[float((mine ** 2).sum()), 20.0, 0.0, 0.0, 0.0, 0.0]], dtype=torch.float64)
        parallel.all_reduce_sum_(tot)
        assert abs(float(tot[0, 0]) - float((err ** 2).sum())) < 1e-12 and float(tot[0, 1]) == 40.0
        want = torch.log((err ** 2).mean())
        assert abs(float(torch.log(tot[0, 0] / (2 * tot[0, 1]))) - float(want)) < 1e-12
        assert abs(float(torch.log((mine ** 2).mean())) - float(want)) > 1e-2  # the log of a per-rank mean is a different number
        assert parallel.world_size() == 2
        with parallel.single_process():  # the switch bench.py --check uses to compute the single-process result next to the DP one
            assert not parallel.is_distributed() and parallel.world_size() == 1
            t = torch.ones(1)
            parallel.all_reduce_sum_(t)
            assert float(t) == 1.0
        assert parallel.is_distributed()
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback

        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok" and ret.get(1) == "ok", (ret.get(0), ret.get(1))
