"""Per-op parity cases shared by tests/test_gpu_ops.py and tools/gpu_diag.py.

Each case runs ONE kernel through the C ABI (fm_op_* hooks, host fp32 channels-last in/out) and
compares with a plain PyTorch-CPU fp32/fp64 reference of the same op on the same bf16-rounded
inputs. Tolerance (stated once): device tensors are bf16 with fp32 accumulation, so an output
element may differ from the exact result by one bf16 rounding of itself plus accumulation-order
noise:  |got - ref| <= 2^-7 * |ref| + 2^-8 * max|ref|   (bf16 has 8 significand bits).
Gradients w.r.t. weights are fp32 end to end: rel. error of the whole tensor <= 2e-3.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from fetal_net import _lib


def bf16_round(a):
    return torch.as_tensor(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def close_bf16(got, ref):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    tol = 2.0 ** -7 * np.abs(ref) + 2.0 ** -8 * max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref)
    worst = float((err / tol).max())
    return worst <= 1.0, worst


def rel_err(got, ref):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def cl_to_cf(x):  # [N,X,Y,Z,C] -> [N,C,X,Y,Z]
    return torch.as_tensor(x).permute(0, 4, 1, 2, 3).contiguous()


def cf_to_cl(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().numpy()


def keras_to_torch_w(w):  # (k,k,k,Cin,Cout) -> (Cout,Cin,k,k,k)
    return torch.as_tensor(w).permute(4, 3, 0, 1, 2).contiguous()


CONV_CASES = [
    # name, N, X, Y, Z, C1, C2, Cout, k
    ("c16_32_8cube", 2, 8, 8, 8, 16, 0, 32, 3),         # SW32 operands (enc0b shape class)
    ("c32_32_16cube", 1, 16, 16, 16, 32, 0, 32, 3),     # SW64
    ("c64_64", 2, 8, 8, 8, 64, 0, 64, 3),               # SW128
    ("c128_256", 1, 8, 8, 8, 128, 0, 256, 3),           # 2 K chunks, 2 N tiles
    ("cat64_32_to32", 1, 16, 16, 16, 64, 32, 32, 3),    # dec0a: two sources, mixed swizzle
    ("cat256_128_to128", 1, 8, 8, 8, 256, 128, 128, 3), # dec2a
    ("ragged_12x10x6", 2, 12, 10, 6, 32, 0, 64, 3),     # partial tiles / clipping
    ("tiny_4cube", 3, 4, 4, 4, 16, 0, 16, 3),           # box larger than the tensor
    ("k1_64_32", 2, 8, 8, 8, 64, 0, 32, 1),             # 1x1x1 (Isensee localisation shape)
    ("cfg_64cube_16_32", 1, 64, 64, 64, 16, 0, 32, 3),  # enc0b at the real patch size
]

# shapes the plane-marching kernel covers (Y % 16 == 0, Z % 8 == 0, filter bank resident in smem)
MARCH_CASES = [
    ("m32_32_16cube", 1, 16, 16, 16, 32, 0, 32, 3),
    ("m16_32_32cube", 2, 32, 32, 32, 16, 0, 32, 3),     # SW32 slabs, ring wrap (X > 16 blocks)
    ("m32_64_16x32x16", 1, 16, 32, 16, 32, 0, 64, 3),   # N = 64 (ring of 8 blocks)
    ("m_oddx_10x16x8", 3, 10, 16, 8, 32, 0, 32, 3),     # X not a multiple of the x-chunk
    ("m_cat64_32_32cube", 2, 32, 32, 32, 64, 32, 32, 3),  # dec0a: two sources, mixed swizzle
    ("m64_16_32cube", 1, 32, 32, 32, 64, 0, 16, 3),     # N = 16 blocks
    ("m_x2", 2, 2, 16, 8, 16, 0, 16, 3),                # only two planes
    ("m_cfg_64cube_32_32", 1, 64, 64, 64, 32, 0, 32, 3),  # dec0b at the real patch size
    # equal plane ranges per CTA that cross column boundaries: segments of 1-4 planes (7 planes per column, 4 per CTA)
    ("m_ragged_7x32x16", 2, 7, 32, 16, 32, 0, 32, 3),
    ("m_ragged_cat_13x32x32", 1, 13, 32, 32, 64, 32, 32, 3),
]


# shapes for the marching wgrad kernel (Cin, Cout multiples of 32)
WGRAD_MARCH_CASES = [
    ("w32_32_16cube", 1, 16, 16, 16, 32, 0, 32, 3),
    ("w32_32_32cube", 2, 32, 32, 32, 32, 0, 32, 3),      # dY ring wrap (X > 8)
    ("w64_32_16x32x16", 1, 16, 32, 16, 64, 0, 32, 3),    # two ci chunks
    ("w32_64_oddx", 3, 10, 16, 8, 32, 0, 64, 3),         # two co chunks, X not multiple of chunk
    ("w128_64_x2", 2, 2, 16, 16, 128, 0, 64, 3),         # 8 pairs, two planes
    ("w_cfg_64cube_32_32", 1, 64, 64, 64, 32, 0, 32, 3),
    ("w16_32_32cube", 2, 32, 32, 32, 16, 0, 32, 3),      # enc0b: Cin = 16 (SWIZZLE_32B X operand, 8 M blocks)
    ("w16_16_32cube", 2, 32, 32, 32, 16, 0, 16, 3),      # Isensee level 0: Cout = 16 (SWIZZLE_32B dY, N = 48)
    ("w32_16_16x16x24", 1, 16, 16, 24, 32, 0, 16, 3),    # Isensee u0_up: 32 -> 16
    ("w_ragged_7x32x16", 2, 7, 32, 16, 32, 0, 32, 3),    # plane ranges crossing column boundaries, 1-plane segments
    ("w_ragged_13x32x32", 1, 13, 32, 32, 64, 0, 32, 3),
]


def conv_fprop_case(ctx, impl, case, seed=0):
    name, N, X, Y, Z, C1, C2, Cout, k = case
    rng = np.random.default_rng(seed)
    x1 = bf16_round(rng.standard_normal((N, X, Y, Z, C1)))
    x2 = bf16_round(rng.standard_normal((N, X, Y, Z, C2))) if C2 else None
    w = bf16_round(rng.standard_normal((k, k, k, C1 + C2, Cout)) / np.sqrt(k ** 3 * (C1 + C2)))
    b = rng.standard_normal(Cout).astype(np.float32)
    y = np.empty((N, X, Y, Z, Cout), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_fprop(ctx.handle, impl, _lib.fptr(x1), _lib.fptr(x2), _lib.fptr(w), _lib.fptr(b),
                                      N, X, Y, Z, C1, C2, Cout, k, 1, _lib.fptr(y)))
    xin = torch.as_tensor(x1 if x2 is None else np.concatenate([x1, x2], -1))
    ref = F.relu(F.conv3d(cl_to_cf(xin).double(), keras_to_torch_w(w).double(), torch.as_tensor(b).double(),
                          padding=k // 2))
    return close_bf16(y, cf_to_cl(ref))


def conv_fprop_raw(ctx, impl, case, seed=0):
    """The raw output of one fprop launch (for run-to-run bit-reproducibility checks)."""
    name, N, X, Y, Z, C1, C2, Cout, k = case
    rng = np.random.default_rng(seed)
    x1 = bf16_round(rng.standard_normal((N, X, Y, Z, C1)))
    x2 = bf16_round(rng.standard_normal((N, X, Y, Z, C2))) if C2 else None
    w = bf16_round(rng.standard_normal((k, k, k, C1 + C2, Cout)) / np.sqrt(k ** 3 * (C1 + C2)))
    b = rng.standard_normal(Cout).astype(np.float32)
    y = np.empty((N, X, Y, Z, Cout), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_fprop(ctx.handle, impl, _lib.fptr(x1), _lib.fptr(x2), _lib.fptr(w), _lib.fptr(b),
                                      N, X, Y, Z, C1, C2, Cout, k, 1, _lib.fptr(y)))
    return y


def conv_dgrad_case(ctx, impl, case, seed=1):
    name, N, X, Y, Z, C1, C2, Cout, k = case
    Cin = C1  # single source
    rng = np.random.default_rng(seed)
    dy = bf16_round(rng.standard_normal((N, X, Y, Z, Cout)))
    w = bf16_round(rng.standard_normal((3, 3, 3, Cin, Cout)) / np.sqrt(27 * Cout))
    act = bf16_round(rng.standard_normal((N, X, Y, Z, Cin)))       # ReLU mask source (act > 0)
    dx = np.empty((N, X, Y, Z, Cin), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_dgrad(ctx.handle, impl, _lib.fptr(dy), _lib.fptr(w), _lib.fptr(act),
                                      N, X, Y, Z, Cin, Cout, _lib.fptr(dx)))
    ref = F.conv_transpose3d(cl_to_cf(dy).double(), keras_to_torch_w(w).double(), padding=1)
    ref = cf_to_cl(ref) * (act > 0)
    return close_bf16(dx, ref)


def conv_wgrad_case(ctx, impl, case, seed=2):
    name, N, X, Y, Z, C1, C2, Cout, k = case
    Cin = C1
    rng = np.random.default_rng(seed)
    x = bf16_round(rng.standard_normal((N, X, Y, Z, Cin)))
    dy = bf16_round(rng.standard_normal((N, X, Y, Z, Cout)))
    dw = np.empty((3, 3, 3, Cin, Cout), np.float32)
    db = np.empty((Cout,), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_wgrad(ctx.handle, impl, _lib.fptr(x), _lib.fptr(dy), N, X, Y, Z, Cin, Cout,
                                      _lib.fptr(dw), _lib.fptr(db)))
    xt = cl_to_cf(x).double().requires_grad_(False)
    wt = torch.zeros(Cout, Cin, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    out = F.conv3d(xt, wt, padding=1)
    out.backward(cl_to_cf(dy).double())
    ref_w = wt.grad.permute(2, 3, 4, 1, 0).numpy()
    ref_b = dy.reshape(-1, Cout).astype(np.float64).sum(0)
    e = max(rel_err(dw, ref_w), rel_err(db, ref_b))
    return e <= 2e-3, e / 2e-3


# decoder convolutions at coarse resolution: name, N, X, Y, Z (fine), Cc (coarse / up source), Cs (skip), Cout
UP_CASES = [
    ("up_dec2a_16cube", 1, 16, 16, 16, 256, 128, 128),   # dec2a of the depth-4 / 16-filter U-Net (2 N tiles in dgrad)
    ("up_dec1a_32cube", 1, 32, 32, 32, 128, 64, 64),     # dec1a
    ("up_dec0a_16cube", 2, 16, 16, 16, 64, 32, 32),      # dec0a channel mix (mixed swizzle widths)
    ("up_ragged_12x20x8", 2, 12, 20, 8, 64, 64, 64),     # coarse extents 6x10x4: partial boxes / clipping
    ("up_small_4cube", 3, 4, 4, 4, 32, 16, 16),          # coarse box larger than the tensor
]


def _up_reference(coarse, skip, w, b, dy=None):
    """torch fp64: conv3d(concatenate([nearest-upsample x2 (coarse), skip])) and its gradients."""
    ct = cl_to_cf(coarse).double().requires_grad_(True)
    up = F.interpolate(ct, scale_factor=2, mode="nearest")
    wt = keras_to_torch_w(w).double().requires_grad_(True)
    out = F.conv3d(torch.cat([up, cl_to_cf(skip).double()], 1), wt, None if b is None else torch.as_tensor(b).double(),
                   padding=1)
    if dy is None:
        return out
    out.backward(cl_to_cf(dy).double())
    return cf_to_cl(ct.grad), wt.grad.permute(2, 3, 4, 1, 0).numpy()


def conv_up_fprop_case(ctx, case, seed=7):
    name, N, X, Y, Z, Cc, Cs, Cout = case
    rng = np.random.default_rng(seed)
    coarse = bf16_round(rng.standard_normal((N, X // 2, Y // 2, Z // 2, Cc)))
    skip = bf16_round(rng.standard_normal((N, X, Y, Z, Cs)))
    w = bf16_round(rng.standard_normal((3, 3, 3, Cc + Cs, Cout)) / np.sqrt(27 * (Cc + Cs)))
    b = rng.standard_normal(Cout).astype(np.float32)
    y = np.empty((N, X, Y, Z, Cout), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_up_fprop(ctx.handle, _lib.fptr(coarse), _lib.fptr(skip), _lib.fptr(w), _lib.fptr(b),
                                         N, X, Y, Z, Cc, Cs, Cout, 1, _lib.fptr(y)))
    ref = F.relu(_up_reference(coarse, skip, w, b)).detach()
    return close_bf16(y, cf_to_cl(ref))


def conv_up_bwd_case(ctx, case, seed=8):
    """Gradient towards the coarse tensor (ReLU-masked) within the bf16 bound, weight gradient of the up-source
    channels within 2e-3 (fp32 accumulation end to end)."""
    name, N, X, Y, Z, Cc, Cs, Cout = case
    rng = np.random.default_rng(seed)
    coarse = bf16_round(rng.standard_normal((N, X // 2, Y // 2, Z // 2, Cc)))
    skip = np.zeros((N, X, Y, Z, Cs), np.float32)
    dy = bf16_round(rng.standard_normal((N, X, Y, Z, Cout)))
    w = bf16_round(rng.standard_normal((3, 3, 3, Cc + Cs, Cout)) / np.sqrt(27 * Cout))
    dc = np.empty_like(coarse)
    dw = np.empty((3, 3, 3, Cc, Cout), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_up_bwd(ctx.handle, _lib.fptr(coarse), _lib.fptr(dy), _lib.fptr(w), N, X, Y, Z, Cc, Cs,
                                       Cout, 1, _lib.fptr(dc), _lib.fptr(dw)))
    ref_dc, ref_dw = _up_reference(coarse, skip, w, None, dy)
    ok, worst = close_bf16(dc, ref_dc * (coarse > 0))
    e = rel_err(dw, ref_dw[:, :, :, :Cc, :])
    return ok and e <= 2e-3, max(worst, e / 2e-3)


FIRST_CASES = [  # name, N, X, Y, Z, Cout : shapes the im2col + tcgen05 first-layer kernels cover (Y % 16 == 0, Z % 8 == 0)
    ("first16_32cube", 2, 32, 32, 32, 16),
    ("first32_16x32x8", 3, 5, 32, 8, 32),      # Isensee / 32-filter variants; single z tile, odd x
    ("first16_64cube", 1, 64, 64, 64, 16),     # the BASELINE patch
    ("first16_ragged_9x12x6", 2, 9, 12, 6, 16),  # not covered by the tensor-core kernel: SIMT route
]


def conv_first_case(ctx, case, seed=9):
    """First conv (fp32 single-channel input) forward within the bf16 bound (the tensor-core route rounds the volume to
    bf16, so the reference does too) and its weight gradient within 2e-3."""
    name, N, X, Y, Z, Cout = case
    rng = np.random.default_rng(seed)
    x = bf16_round(rng.standard_normal((N, X, Y, Z)))
    w = bf16_round(rng.standard_normal((3, 3, 3, 1, Cout)) / np.sqrt(27.0))
    b = rng.standard_normal(Cout).astype(np.float32)
    dy = bf16_round(rng.standard_normal((N, X, Y, Z, Cout)))
    y = np.empty((N, X, Y, Z, Cout), np.float32)
    dw = np.empty((3, 3, 3, 1, Cout), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_conv3d_first(ctx.handle, _lib.fptr(x), _lib.fptr(w), _lib.fptr(b), N, X, Y, Z, Cout, 1,
                                      _lib.fptr(y), _lib.fptr(dy), _lib.fptr(dw)))
    xt = torch.as_tensor(x)[:, None].double()
    wt = keras_to_torch_w(w).double().requires_grad_(True)
    out = F.conv3d(xt, wt, torch.as_tensor(b).double(), padding=1)
    ok, worst = close_bf16(y, cf_to_cl(F.relu(out).detach()))
    out.backward(cl_to_cf(dy).double())
    e = rel_err(dw, wt.grad.permute(2, 3, 4, 1, 0).numpy())
    return ok and e <= 2e-3, max(worst, e / 2e-3)


def maxpool_case(ctx, seed=3):
    N, X, Y, Z, C = 2, 8, 12, 16, 32
    rng = np.random.default_rng(seed)
    x = bf16_round(rng.standard_normal((N, X, Y, Z, C)))
    y = np.empty((N, X // 2, Y // 2, Z // 2, C), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_maxpool3d(ctx.handle, _lib.fptr(x), N, X, Y, Z, C, _lib.fptr(y)))
    ref = cf_to_cl(F.max_pool3d(cl_to_cf(x), 2))
    ok = np.array_equal(y, ref)                                    # exact: max of bf16 values
    # backward with skip-gradient add and ReLU mask
    dy = bf16_round(rng.standard_normal(y.shape))
    dskip = bf16_round(rng.standard_normal(x.shape))
    dx = np.empty_like(x)
    _lib.check(lib.fm_op_maxpool3d_bwd(ctx.handle, _lib.fptr(x), _lib.fptr(dy), _lib.fptr(dskip), N, X, Y, Z, C,
                                       _lib.fptr(dx)))
    xt = cl_to_cf(x).double().requires_grad_(True)
    F.max_pool3d(xt, 2).backward(cl_to_cf(dy).double())
    refb = (cf_to_cl(xt.grad) + dskip) * (x > 0)
    ok2, worst = close_bf16(dx, refb)
    return ok and ok2, worst if ok else float("inf")


def upsample_case(ctx, seed=4):
    N, X, Y, Z, C = 2, 4, 6, 8, 64
    rng = np.random.default_rng(seed)
    x = bf16_round(rng.standard_normal((N, X, Y, Z, C)))
    y = np.empty((N, 2 * X, 2 * Y, 2 * Z, C), np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_upsample3d(ctx.handle, _lib.fptr(x), N, X, Y, Z, C, _lib.fptr(y)))
    ref = np.repeat(np.repeat(np.repeat(x, 2, 1), 2, 2), 2, 3)
    ok = np.array_equal(y, ref)
    dy = bf16_round(rng.standard_normal(y.shape))
    act = bf16_round(rng.standard_normal(x.shape))
    dx = np.empty_like(x)
    _lib.check(lib.fm_op_upsample3d_bwd(ctx.handle, _lib.fptr(dy), _lib.fptr(act), N, X, Y, Z, C, _lib.fptr(dx)))
    refb = dy.astype(np.float64).reshape(N, X, 2, Y, 2, Z, 2, C).sum((2, 4, 6)) * (act > 0)
    ok2, worst = close_bf16(dx, refb)
    return ok and ok2, worst if ok else float("inf")


def dice_case(ctx, seed=5):
    n = 8 * 16 ** 3 + 3           # not a multiple of 4: exercises the scalar tail
    rng = np.random.default_rng(seed)
    p = rng.random(n).astype(np.float32)
    t = (rng.random(n) < 0.3).astype(np.float32)
    sums = np.zeros(8, np.float64)
    g = np.empty(n, np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_dice(ctx.handle, _lib.fptr(p), _lib.fptr(t), n, _lib.dptr(sums), _lib.fptr(g)))
    pb, tb = (p > 0.5), (t > 0.5)
    ref = np.array([(t.astype(np.float64) * p).sum(), t.sum(dtype=np.float64), p.sum(dtype=np.float64),
                    (tb & pb).sum(), tb.sum(), pb.sum(), (t == pb.astype(np.float32)).sum(), n], np.float64)
    e1 = float(np.max(np.abs(sums - ref) / np.maximum(np.abs(ref), 1)))
    I, S = ref[0], ref[1] + ref[2] + 1.0
    gref = -(2.0 * t * S - (2.0 * I + 1.0)) / (S * S)
    e2 = rel_err(g, gref)
    # float32 per-thread partial sums over <= n/151552 elements each, then float64: 1e-6 is ample
    return e1 <= 1e-6 and e2 <= 1e-5, max(e1 / 1e-6, e2 / 1e-5)


def dice_xent_case(ctx, with_mask, seed=15):
    """fm_op_dice_xent (dice_and_xent / dice_and_xent_mask, metrics.py:68-95) against torch autograd in fp64: the ninth
    statistic (sum w * bce) and d(loss)/d(logit) through the sigmoid. Probabilities include saturated ones (the Keras
    clip at 1e-7 is active there: zero cross-entropy gradient)."""
    import torch
    from oracle import unet_oracle as uo
    n = 4 * 16 ** 3 + 2
    rng = np.random.default_rng(seed)
    z = (4.0 * rng.standard_normal(n)).astype(np.float32)
    z[:64] = 40.0 * np.sign(z[:64])                        # saturated: p == 1 or ~4e-18 in fp32
    p = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
    t = (rng.random(n) < 0.3).astype(np.float32)
    mask = (6.0 * rng.random(n)).astype(np.float32) if with_mask else None
    xw, sigma = 0.7, 3.0
    sums = np.zeros(9, np.float64)
    g = np.empty(n, np.float32)
    lib = _lib.load()
    _lib.check(lib.fm_op_dice_xent(ctx.handle, _lib.fptr(p), _lib.fptr(t), _lib.fptr(mask) if with_mask else None, n,
                                   xw, sigma, _lib.dptr(sums), _lib.fptr(g)))
    # reference: the loss as a function of the logit, with p fixed to the SAME fp32 probabilities the kernel saw
    pt = torch.tensor(p.astype(np.float64), requires_grad=True)
    tt = torch.tensor(t.astype(np.float64))
    mt = torch.tensor(mask.astype(np.float64)) if with_mask else None
    loss = uo.dice_and_xent(tt, pt, xw, mt, sigma if with_mask else None)
    loss.backward()
    gref = (pt.grad * pt.detach() * (1 - pt.detach())).numpy()       # chain through the sigmoid
    w = np.exp(-mask.astype(np.float64) / sigma) if with_mask else np.ones(n)
    xent_ref = float((torch.as_tensor(w) * uo.binary_crossentropy(tt, pt.detach())).sum())
    e1 = abs(sums[8] - xent_ref) / abs(xent_ref)
    lref = float(loss.detach())
    e2 = abs((-(2 * sums[0] + 1) / (sums[1] + sums[2] + 1) + xw * sums[8] / sums[7]) - lref) / abs(lref)
    e3 = rel_err(g, gref)
    # fp32 log / exp per voxel, fp32 partial sums then fp64: 2e-6 on the sums; 2e-5 on the gradient
    return e1 <= 2e-6 and e2 <= 2e-6 and e3 <= 2e-5, max(e1 / 2e-6, e2 / 2e-6, e3 / 2e-5)


def adam_case(ctx, seed=6):
    from oracle.unet_oracle import keras_adam_step
    n = 100003
    rng = np.random.default_rng(seed)
    p = rng.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    p2, m2, v2 = p.copy(), m.copy(), v.copy()
    lib = _lib.load()
    worst = 0.0
    for it in range(3):
        g = rng.standard_normal(n).astype(np.float32) * 1e-3
        _lib.check(lib.fm_op_adam(ctx.handle, _lib.fptr(p), _lib.fptr(g), _lib.fptr(m), _lib.fptr(v), n, it, 1e-3))
        keras_adam_step(p2, g, m2, v2, it, 1e-3)
        worst = max(worst, float(np.max(np.abs(p - p2))), float(np.max(np.abs(m - m2))))
    return worst <= 2e-6, worst / 2e-6
