"""world_size-2 gloo tests (CPU) of the data-parallel host logic of fetal_net.distributed:
 (1) whole-batch soft Dice under batch sharding: all-reducing the 3 sums and SUM-reducing the local
     gradients reproduces the single-process loss and gradient (metrics.py:11-15 flattens the batch axis);
 (2) patch-sharded sliding-window inference: contiguous shards + one SUM reduce of float64 partial sums
     == single-process patch_wise_prediction, bit for bit; counts are analytic and identical on all ranks.
The device kernels are replaced by the oracle here (no GPU in this tier); the sharding / reduction rules
under test are the ones fetal_net.distributed and fm_patchwise_predict apply.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import prediction_oracle as po
from oracle import unet_oracle as uo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _dice_worker(rank, world, port, out):
    _init(rank, world, port)
    torch.manual_seed(0)
    B = 4
    logits = torch.randn(B, 1, 6, 6, 6, dtype=torch.float64)
    t = (torch.rand(B, 1, 6, 6, 6) < 0.3).double()
    lo, hi = B * rank // world, B * (rank + 1) // world
    z = logits[lo:hi].clone().requires_grad_(True)
    p = torch.sigmoid(z)
    tl = t[lo:hi]
    sums = torch.stack([(tl * p).sum(), tl.sum(), p.sum()]).detach()
    dist.all_reduce(sums)                                                  # forward exchange
    I, S = sums[0], sums[1] + sums[2] + 1.0
    dl_dp = -(2.0 * tl * S - (2.0 * I + 1.0)) / (S * S)                    # closed form with GLOBAL sums
    p.backward(dl_dp)
    # a "parameter" shared by all ranks: scalar gain on the logits -> its gradient is the SUM over ranks
    g_gain = (z.grad * logits[lo:hi]).sum().reshape(1)
    dist.all_reduce(g_gain)
    loss = -(2.0 * I + 1.0) / S
    if rank == 0:
        # single-process reference
        zz = logits.clone().requires_grad_(True)
        gain = torch.ones(1, dtype=torch.float64, requires_grad=True)
        L = uo.dice_coefficient_loss(t, torch.sigmoid(zz * gain))
        L.backward()
        out.put((float(loss), float(L), float(g_gain), float(gain.grad)))
    dist.destroy_process_group()


def test_dice_sum_allreduce_and_gradient_sum():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dice_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    loss, ref_loss, g, ref_g = res
    assert loss == pytest.approx(ref_loss, rel=1e-12)
    assert g == pytest.approx(ref_g, rel=1e-10)


def _infer_worker(rank, world, port, out):
    _init(rank, world, port)
    from fetal_net.distributed import shard_range
    from tests.golden.make_golden import ramp_model
    rng = np.random.default_rng(11)
    vol = rng.standard_normal((1, 40, 36, 20)).astype(np.float32)
    patch = (16, 16, 16)
    fn, oshape = ramp_model(patch, 1)
    d0, fit, _ = po.pad_volume(vol[0], patch, patch)
    idx = po.patch_plan(d0.shape, patch, patch, 0.5)
    lo, hi = shard_range(len(idx), rank, world)
    part = np.zeros(d0.shape + (1,))
    for ci in idx[lo:hi]:                      # ascending patch order inside the shard
        pp = fn(po.extract_patch(d0, patch, ci)[None, None])[0].transpose(1, 2, 3, 0)
        x, y, z = ci
        part[x:x + 16, y:y + 16, z:z + 16] += pp
    cnt = po.count_map(d0.shape, patch, idx)   # analytic, never communicated
    tpart = torch.as_tensor(part)
    dist.reduce(tpart, 0)
    if rank == 0:
        from oracle.ref_harness import FunctionModel
        full = po.patch_wise_prediction(FunctionModel(fn, oshape), vol, patch, overlap_factor=0.5)
        got = tpart.numpy() / cnt[..., None]
        out.put((bool(np.array_equal(got, full)), float(np.abs(got - full).max()), lo, hi, len(idx)))
    dist.destroy_process_group()


def test_patch_sharded_inference_reduce_is_bit_exact():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_infer_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    ok, err, lo, hi, n = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert ok, err
    assert (lo, hi) == (0, n // 2)


def test_shard_ranges_partition():
    from fetal_net.distributed import shard_range
    for n in [1, 7, 49, 147, 625]:
        for world in [1, 2, 3, 4, 8]:
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
