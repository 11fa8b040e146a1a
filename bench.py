#!/usr/bin/env python
"""bench.py — throughput of the hot path on B200 (and the reference's CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload train|infer]

One "step" = one soft-Dice training step (fwd + loss + bwd + Keras-Adam) of the depth-4 / 16-filter 3D
U-Net on a batch of 8 synthetic 64^3 patches per GPU (BASELINE.json configs[1]; weak scaling: global
batch 8N with the Dice-sum + gradient all-reduce over NCCL). `--workload infer` times configs[0]
(patch_wise_prediction of one 256x256x64 volume, patch 64^3, overlap_factor 0.5 -> 49 patches).

JSON keys (one line, rank 0): see the task contract. `value` is timed with CUDA events on the stream the
kernels are launched on, inputs already resident in HBM; `e2e` goes through the reference-facing Python API
(Model.train_on_batch / patch_wise_prediction) with host buffers, H2D + D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "fetal-mri-segmentation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

PATCH = (64, 64, 64)
BATCH = 8
DEPTH, NF = 4, 16
VOLUME = (256, 256, 64)
OVERLAP = 0.5
METRIC = "U-Net train voxels/sec"
UNIT = "voxels/s"

# algorithmic FLOPs: 3D U-Net d4 nf16 @ 64^3
FWD_GF_PER_PATCH = 118.472
FIRST_CONV_GF = 0.226


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe). NVML is polled from a thread
    every few ms (nvidia-smi -lms cannot deliver a sample inside a 50 ms region); the fields are the ones
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device):
        self.device, self.rows, self.thread, self.run = device, [], None, False
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(int(self.device))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.run = True
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        if self.nv is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvml unavailable"], samples=0)
        self.run = False
        self.thread.join(timeout=1)
        sm = [r[0] for r in self.rows]
        reasons = sorted({nm for _, rs in self.rows for bit, nm in self.REASONS.items() if rs & bit})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.max_mhz, reasons=reasons,
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (PyTorch-CPU fp32 restatement of the Keras/TF path) on the
# host cores. The reference itself cannot run here (no TensorFlow/Keras in the image).
# ------------------------------------------------------------------------------------------------
def cpu_train_step_time(batch, steps, warmup, budget_s):
    import torch
    from oracle import unet_oracle as uo
    torch.set_num_threads(os.cpu_count())
    rng = np.random.default_rng(1)
    w = uo.glorot_uniform_weights(uo.unet3d_layers(DEPTH, NF), seed=0)
    x = rng.standard_normal((batch, 1) + PATCH).astype(np.float32)
    t = (np.random.default_rng(2).random(x.shape) < 0.3).astype(np.float32)
    state = {}
    times = []
    t_start = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        uo.unet3d_train_step(x, t, w, state, 1e-4)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    return float(np.median(times)), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    sample_batch = 2     # bounded sample of the batch-8 step: 2 of the 8 patches per step
    sec, n = cpu_train_step_time(sample_batch, args.steps, min(args.warmup, 1), budget_s=150)
    vps = sample_batch * int(np.prod(PATCH)) / sec
    sample = "oracle train step on %d of %d patches (64^3), %d timed steps, torch %d threads" % (
        sample_batch, BATCH, n, torch.get_num_threads())
    line = dict(impl="reference", metric=METRIC, value=vps, unit=UNIT, n_gpus=args.gpus, steps=n,
                warmup=min(args.warmup, 1), ms_per_step=sec * 1e3 * BATCH / sample_batch,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(args.gpus),
                cpu_baseline=dict(value=vps, unit=UNIT, cores=os.cpu_count(), kind="port", sample=sample),
                e2e=dict(value=vps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(n_gpus):
    return dict(workload="unet_model_3d(depth=4,n_base_filters=16) soft-Dice train step, batch 8 x 1x64x64x64 per GPU "
                         "(BASELINE configs[1])",
                global_batch=BATCH * n_gpus, patch=list(PATCH), parallelism="dp%d" % n_gpus,
                l2="working set ~1.9 GB of activations+gradients per step >> 126 MB L2 (no explicit flush needed)")


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from fetal_net import _lib
    from fetal_net.distributed import DataParallelTrainer
    from fetal_net.model import unet_model_3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    ctx = _lib.get_context(local)
    lib = _lib.load()
    model = unet_model_3d(input_shape=(1,) + PATCH, n_base_filters=NF, depth=DEPTH, initial_learning_rate=1e-4,
                          device=local)
    model.init_glorot_uniform(seed=0)
    rng = np.random.default_rng(1 + rank)
    x = rng.standard_normal((BATCH, 1) + PATCH).astype(np.float32)
    t = (np.random.default_rng(2 + rank).random(x.shape) < 0.3).astype(np.float32)
    xd, td = torch.as_tensor(x).cuda(), torch.as_tensor(t).cuda()
    xp, tp = torch.as_tensor(x).pin_memory(), torch.as_tensor(t).pin_memory()
    stream = torch.cuda.ExternalStream(ctx.stream, device="cuda:%d" % local)
    dp = DataParallelTrainer(model) if world > 1 else None
    m4 = np.zeros(4, np.float32)
    lr = 1e-4

    def step_device():
        if dp is None:
            _lib.check(lib.fm_train_step_device(model._h, xd.data_ptr(), td.data_ptr(), BATCH, lr, _lib.fptr(m4)))
        else:
            # device-resident inputs are staged by fm_train_forward from pinned host memory in the DP path;
            # the collective structure (sums all-reduce, bucketed gradient all-reduce) is the real one
            dp.train_on_batch(xp.numpy(), tp.numpy())

    def step_e2e():
        if dp is None:
            return model.train_on_batch(xp.numpy(), tp.numpy())
        return dp.train_on_batch(xp.numpy(), tp.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_device()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    ms_step = ms / args.steps
    vox_step = BATCH * int(np.prod(PATCH)) * world
    value = vox_step / (ms_step * 1e-3)

    for _ in range(2):
        step_e2e()
    # wall-clock around the public API (the user's view), max over ranks
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = dict(value=vox_step / e2e_s, unit=UNIT, h2d_bytes_per_step=int(x.nbytes + t.nbytes),
               d2h_bytes_per_step=64, ms_per_step=e2e_s * 1e3)

    # roofline leg: two extra steps with per-launch CUDA events, dominant kernel by total time
    roof, breakdown = None, None
    pk = peaks()
    if rank == 0:
        ctx.profile(True)
        nprof = 2
        for _ in range(nprof):
            if dp is None:
                _lib.check(lib.fm_train_step_device(model._h, xd.data_ptr(), td.data_ptr(), BATCH, lr, _lib.fptr(m4)))
            else:
                _lib.check(lib.fm_train_forward(model._h, _lib.fptr(x), _lib.fptr(t), BATCH))
                _lib.check(lib.fm_train_backward(model._h))
        recs = ctx.profile_records()
        ctx.profile(False)
        agg = {}
        for name, kms, fl, by in recs:
            a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += kms
            a[2] += fl
            a[3] += by
        total_ms = sum(a[1] for a in agg.values())
        breakdown = {k: dict(launches=a[0] // nprof, ms_per_step=a[1] / nprof, share=a[1] / total_ms,
                             tflops=(a[2] / a[1] / 1e9) if a[1] > 0 and a[2] > 0 else None,
                             gbs=(a[3] / a[1] / 1e6) if a[1] > 0 and a[3] > 0 else None)
                     for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        top = max(agg.items(), key=lambda kv: kv[1][1])
        name, (cnt, kms, fl, by) = top
        # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/)
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_dram_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            key = {"conv3d_wgrad_march": "wgrad/conv3d_wgrad_march_kernel", "conv3d_march_fprop": "fprop/conv3d_march_kernel",
                   "conv3d_march_dgrad": "dgrad/conv3d_march_kernel"}.get(name)
            if key in tj:
                traffic = tj[key]
                traffic_note = "ncu capture of the dec0b-sized launch (32->32 @ 8x64^3; algorithmic bytes 268.4e6)"
        if fl > 0:
            ach = fl / kms / 1e9        # TFLOP/s: algorithmic flops per launch / avg launch duration
            roof = dict(kernel=name, bound="tensor", achieved=ach, peak=pk["tf_sustained"], unit="TFLOP/s",
                        frac=ach / pk["tf_sustained"], traffic=traffic, traffic_note=traffic_note,
                        peak_source=pk["src"] + " (sustained)",
                        launches_per_step=cnt // nprof, avg_launch_ms=kms / cnt,
                        algorithmic_flops_per_launch=fl / cnt, share_of_step=kms / total_ms)
        else:
            ach = by / kms / 1e6
            roof = dict(kernel=name, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                        traffic=None, peak_source=pk["src"], launches_per_step=cnt // nprof,
                        avg_launch_ms=kms / cnt, algorithmic_bytes_per_launch=by / cnt, share_of_step=kms / total_ms)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        sec, n = cpu_train_step_time(2, 2, 1, budget_s=40)
        cpu_base = dict(value=2 * int(np.prod(PATCH)) / sec, unit=UNIT, cores=os.cpu_count(), kind="port",
                        sample="oracle (PyTorch-CPU fp32 restatement) train step on 2 of 8 patches, %d timed steps, "
                               "%d threads" % (n, _t.get_num_threads()))

    if rank == 0:
        step_gf = world * BATCH * (3 * FWD_GF_PER_PATCH - FIRST_CONV_GF)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="bf16", data="synthetic", config=workload_config(world),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu_base,
                    conv_tflops_whole_step=step_gf / ms_step, kernel_breakdown=breakdown,
                    loss=float(m4[0]) if dp is None else None)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_infer(args):
    """configs[0]: patch_wise_prediction of one 256x256x64 volume (49 patches) — extra bench line."""
    import torch
    from fetal_net import _lib
    from fetal_net.model import unet_model_3d
    from fetal_net.prediction import patch_wise_prediction
    ctx = _lib.get_context(0)
    model = unet_model_3d(input_shape=(1,) + PATCH, n_base_filters=NF, depth=DEPTH)
    model.init_glorot_uniform(seed=0)
    vol = np.random.default_rng(0).standard_normal((1,) + VOLUME).astype(np.float32)
    for _ in range(max(args.warmup, 3)):
        out = patch_wise_prediction(model, vol, PATCH, overlap_factor=OVERLAP, batch_size=49)
    l0 = ctx.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = patch_wise_prediction(model, vol, PATCH, overlap_factor=OVERLAP, batch_size=49)
    sec = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    launches = int((ctx.launch_count() - l0) // args.steps)
    nvox = int(np.prod(VOLUME))
    # device leg: the same call with per-launch CUDA events (gather + network + overlap-add, no host copies)
    ctx.profile(True)
    t1 = time.perf_counter()
    out = patch_wise_prediction(model, vol, PATCH, overlap_factor=OVERLAP, batch_size=49)
    prof_wall = time.perf_counter() - t1
    agg = {}
    for name, kms, fl, by in ctx.profile_records():
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += kms
        a[2] += fl
        a[3] += by
    ctx.profile(False)
    dev_ms = sum(a[1] for a in agg.values())
    breakdown = {k: dict(launches=a[0], ms=a[1], tflops=(a[2] / a[1] / 1e9 if a[2] else None),
                         gbs=(a[3] / a[1] / 1e6 if a[3] else None))
                 for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
    breakdown["_sum_kernels_ms"] = dev_ms
    breakdown["_wall_ms_profiled_call"] = prof_wall * 1e3
    pk = peaks()
    top = max(agg.items(), key=lambda kv: kv[1][1])
    roof = dict(kernel=top[0], bound="tensor", achieved=top[1][2] / top[1][1] / 1e9, peak=pk["tf_sustained"],
                unit="TFLOP/s", frac=top[1][2] / top[1][1] / 1e9 / pk["tf_sustained"], traffic=None,
                peak_source=pk["src"] + " (sustained)", launches_per_step=top[1][0], share_of_step=top[1][1] / dev_ms)
    line = dict(metric="U-Net infer voxels/sec", value=nvox / (dev_ms * 1e-3), unit=UNIT, n_gpus=1, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=dev_ms, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload="patch_wise_prediction 1x256x256x64, patch 64^3, overlap_factor 0.5, 49 patches "
                                     "(BASELINE configs[0])", patches=49, batch_size=49,
                            value_is="output voxels / summed kernel time of one call (volume resident in HBM)",
                            l2="49 patches x 64^3 x up to 96 channels of activations per call >> 126 MB L2"),
                clocks=clocks,
                e2e=dict(value=nvox / sec, unit=UNIT, h2d_bytes_per_step=int(vol.nbytes),
                         d2h_bytes_per_step=int(out.nbytes), ms_per_step=sec * 1e3),
                gpu_launches=launches, roofline=roof,
                patch_voxels_per_s=49 * int(np.prod(PATCH)) / sec,
                conv_tflops=49 * FWD_GF_PER_PATCH / dev_ms, out_mean=float(out.mean()), kernel_breakdown=breakdown)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "infer":
        run_infer(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
