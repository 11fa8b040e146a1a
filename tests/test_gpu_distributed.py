"""2-GPU tests of fetal_net.distributed (NCCL): data-parallel training step == single-GPU step on the full
batch (global Dice, summed gradients), and patch-sharded inference == single-GPU patch_wise_prediction.
Skipped on boxes with fewer than 2 GPUs (`gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    # bit-reproducible training kernels: per-sample activations are then identical in the 2 x 2 and the 1 x 4 runs and
    # only the fp32 summation order of the weight gradients differs
    os.environ["FETAL_B200_DETERMINISTIC"] = "1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda:%d" % rank))
    from fetal_net.distributed import DataParallelTrainer, sharded_patch_wise_prediction
    from fetal_net.model import unet_model_3d
    from fetal_net.prediction import patch_wise_prediction
    from oracle import unet_oracle as uo
    from tests.test_gpu_model import blob_target, decisive_weights

    w = decisive_weights(uo.unet3d_layers(4, 16), seed=2)
    rng = np.random.default_rng(9)
    x = rng.standard_normal((4, 1, 32, 32, 32)).astype(np.float32)
    t = blob_target(x.shape, rng)
    model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, initial_learning_rate=1e-4, device=rank)
    model.set_named_weights(w)
    dp = DataParallelTrainer(model)
    lo, hi = 4 * rank // world, 4 * (rank + 1) // world
    res = dp.train_on_batch(x[lo:hi], t[lo:hi])
    grads = model.get_gradients()
    vol = rng.standard_normal((1, 72, 40, 32)).astype(np.float32)
    sharded = sharded_patch_wise_prediction(model, vol, (32, 32, 32), overlap_factor=0.5)
    if rank == 0:
        # single-GPU references on the same device
        ref_model = unet_model_3d(input_shape=(1, 32, 32, 32), n_base_filters=16, initial_learning_rate=1e-4, device=0)
        ref_model.set_named_weights(w)
        ref = ref_model.train_on_batch(x, t)
        ref_grads = ref_model.get_gradients()
        cos = []
        for g, r in zip(grads, ref_grads):
            g, r = g.astype(np.float64).ravel(), r.astype(np.float64).ravel()
            cos.append(float(g @ r / max(np.linalg.norm(g) * np.linalg.norm(r), 1e-300)))
        # inference reference with the post-step weights of the DP model (identical on all ranks)
        single = patch_wise_prediction(model, vol, (32, 32, 32), overlap_factor=0.5)
        out.put(dict(res=res, ref=ref, min_cos=min(cos), argmin=int(np.argmin(cos)), cos=[round(v, 5) for v in cos], infer_equal=bool(np.array_equal(sharded, single)),
                     infer_maxdiff=float(np.abs(sharded - single).max())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_step_and_sharded_inference():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    r = q.get(timeout=300)
    [p.join(timeout=120) for p in procs]
    # same global Dice loss / metrics as the single-process step on the full batch
    # (FETAL_B200_DETERMINISTIC=1 in the workers: forward activations are bit-identical per sample)
    assert r["res"][0] == pytest.approx(r["ref"][0], abs=1e-6), r
    assert r["res"][1] == pytest.approx(r["ref"][1], abs=1e-6), r
    # summed gradients == full-batch gradients (only the fp32 red.add order differs)
    assert r["min_cos"] >= 0.9999, r
    assert r["infer_equal"], r
