"""Data-parallel extension of the hot path (the reference is single-process: training.py:115-117).

One process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch) as plumbing:

  training   batch sharded over ranks, weights replicated. Soft Dice is a WHOLE-batch statistic
             (metrics.py:11-15 flattens the batch axis), so the forward all-reduces the 8 loss sums and
             every rank back-propagates the global Dice; the gradient of that one global scalar is the
             SUM (not the mean) of the ranks' local gradients. The flat fp32 gradient buffer is
             all-reduced bucket by bucket on a side stream while the rest of backward still runs
             (fm_stream_wait_bucket), then every rank applies the identical Adam step.
  inference  the patch list of patch_wise_prediction is split contiguously over ranks; each rank
             overlap-adds its patches into a private float64 partial sum, one reduce, then sum / count.
             Counts are analytic (identical on every rank) and never communicated.
"""
import numpy as np

from . import _lib


def shard_range(n, rank, count):
    """Contiguous share [lo, hi) of n patches for `rank` of `count` (same rule as fm_patchwise_predict)."""
    return n * rank // count, n * (rank + 1) // count


class _DevPtr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=typestr, data=(int(ptr), False), version=2)


def _wrap(ptr, n, typestr, device):
    import torch
    return torch.as_tensor(_DevPtr(ptr, n, typestr), device="cuda:%d" % device)


class DataParallelTrainer:
    """Wraps a native Model for multi-GPU training (call after torch.distributed.init_process_group)."""

    def __init__(self, model, overlap=True):
        import ctypes
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.model = model
        self.lib = _lib.load()
        self.device = model._ctx.device
        self.overlap = overlap
        p, n = _lib.c_u64(0), _lib.c_i64(0)
        _lib.check(self.lib.fm_model_grad_buffer(model._h, ctypes.byref(p), ctypes.byref(n)))
        self.grads = _wrap(p.value, n.value, "<f4", self.device)
        _lib.check(self.lib.fm_model_loss_sums(model._h, ctypes.byref(p)))
        self.sums = _wrap(p.value, 8, "<f8", self.device)
        self.buckets = []
        for b in range(self.lib.fm_model_num_buckets(model._h)):
            off, cnt = _lib.c_i64(0), _lib.c_i64(0)
            _lib.check(self.lib.fm_model_bucket_range(model._h, b, ctypes.byref(off), ctypes.byref(cnt)))
            self.buckets.append((off.value, cnt.value))
        self.compute_stream = torch.cuda.ExternalStream(model._ctx.stream, device="cuda:%d" % self.device)
        self.comm_stream = torch.cuda.Stream(device="cuda:%d" % self.device)

    def broadcast_weights(self, src=0):
        ws = self.model.get_weights()
        t = [self.torch.as_tensor(w).cuda(self.device) for w in ws]
        for x in t:
            self.dist.broadcast(x, src)
        self.model.set_weights([x.cpu().numpy() for x in t])

    def train_on_batch(self, x, y):
        """x, y: this rank's shard of the global batch. Returns the GLOBAL [loss, acc, vod]."""
        torch, dist, lib, h = self.torch, self.dist, self.lib, self.model._h
        x, y = _lib.f32c(x), _lib.f32c(y)
        _lib.check(lib.fm_train_forward(h, _lib.fptr(x), _lib.fptr(y), int(x.shape[0])))
        with torch.cuda.stream(self.compute_stream):
            dist.all_reduce(self.sums)                                   # 8 doubles: global Dice statistics
        _lib.check(lib.fm_train_metrics_async(h))                        # global statistics -> pinned host, async
        _lib.check(lib.fm_train_backward(h))
        # buckets complete in order 0,1,2,... (backward runs in reverse layer order)
        with torch.cuda.stream(self.comm_stream if self.overlap else self.compute_stream):
            for b, (off, cnt) in enumerate(self.buckets):
                if self.overlap:
                    _lib.check(lib.fm_stream_wait_bucket(h, self.comm_stream.cuda_stream, b))
                dist.all_reduce(self.grads[off:off + cnt])               # SUM: one global loss scalar
        m = np.zeros(4, np.float32)
        _lib.check(lib.fm_train_apply(h, float(self.model.optimizer.lr),
                                      self.comm_stream.cuda_stream if self.overlap else 0, None))
        # returns once the statistics are on the host; backward / all-reduce / Adam keep running, ordered before any
        # later call on the model, and the next step's upload (pinned inputs) overlaps them
        _lib.check(lib.fm_train_metrics_wait(h, _lib.fptr(m)))
        return [float(v) for v in m[:len(self.model.metrics_names)]]


def sharded_patch_wise_prediction(model, data, patch_shape, overlap_factor=0, batch_size=5, dst=0):
    """patch_wise_prediction with the patch list sharded over the ranks of the default process group.
    Returns the float64 [X,Y,Z,C] volume on rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from .prediction import patch_wise_prediction
    rank, count = dist.get_rank(), dist.get_world_size()
    if count == 1:
        return patch_wise_prediction(model, data, patch_shape, overlap_factor, batch_size)
    part, cnt = patch_wise_prediction(model, data, patch_shape, overlap_factor, batch_size, shard=(rank, count))
    dev = "cuda:%d" % model._ctx.device if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(part).to(dev)
    dist.reduce(t, dst)                                                  # float64 partial sums
    if rank != dst:
        return None
    assert np.all(cnt > 0), 'Found zeros in count'
    return t.cpu().numpy() / cnt[..., None]
