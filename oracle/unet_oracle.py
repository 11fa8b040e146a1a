"""TEST INFRASTRUCTURE ONLY (the product path must never import `oracle/`).

CPU restatement (PyTorch-CPU, fp32 or fp64) of the reference network path.
PARITY UNPINNED for the network: Keras/TensorFlow are not installable in this
image and the reference's own tests hold no numerics for the
builders, so the restatement is anchored on the reference's call sites and the
published Keras-2.2 / TF-1.x / keras_contrib semantics:

  * `unet3d_layers` / `unet3d_forward`  <- unet_model_3d            fetal_net/model/unet3d/unet.py:40-70
                                           create_convolution_block  fetal_net/model/unet3d/unet.py:89-115
                                           get_up_convolution        fetal_net/model/unet3d/unet.py:132-138
  * `isensee3d_forward`                 <- isensee2017_model_3d      fetal_net/model/unet3d/isensee2017.py:39-79,95-111
  * `unet2d_forward`                    <- unet_model_2d             fetal_net/model/unet/unet.py:49-85,91-118
  * `dice_coefficient(_loss)`, `vod_coefficient`  <- fetal_net/metrics.py:11-32
  * `keras_adam_step`                   <- Adam(lr) in unet3d/unet.py:85 (Keras 2.x update rule, App. A.9)
  * `glorot_uniform_weights`            <- Keras default initialisers (App. A.2)

Weights are held in **Keras layout**: kernel (k0,k1,k2,Cin,Cout), bias (Cout,).
Third-party arithmetic restated: Keras>=2 (requirements.txt:7, unpinned), TensorFlow 1.x
(not listed), keras_contrib InstanceNormalization (git HEAD, unpinned).
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# layer tables
# ----------------------------------------------------------------------------------------------

def unet3d_layers(depth=4, n_base_filters=32, in_channels=1, n_labels=1, deconvolution=False):
    """[(name, cin, cout, k)] in Keras creation order (unet3d/unet.py:45-68). deconvolution=True adds the
    Deconvolution3D(filters = channels of the coarse tensor, kernel 2, strides 2) layers "up<d>" (unet.py:57-59,
    132-136), created before the block that consumes them; their Keras kernel is (2,2,2,Cout,Cin)."""
    layers = []
    c = in_channels
    skips = []
    for d in range(depth):
        f1 = n_base_filters * (2 ** d)
        f2 = f1 * 2
        layers.append(("enc%da" % d, c, f1, 3))
        layers.append(("enc%db" % d, f1, f2, 3))
        skips.append(f2)
        c = f2
    for d in range(depth - 2, -1, -1):
        f = skips[d]
        if deconvolution:
            layers.append(("up%d" % d, c, c, 2))
        layers.append(("dec%da" % d, c + f, f, 3))
        layers.append(("dec%db" % d, f, f, 3))
        c = f
    layers.append(("final", c, n_labels, 1))
    return layers


def glorot_uniform_weights(layers, seed=0, ndim=3):
    """Keras glorot_uniform: limit = sqrt(6/(fan_in+fan_out)), fan = k^d*C; zero biases."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, cin, cout, k in layers:
        rf = k ** ndim
        limit = np.sqrt(6.0 / (rf * cin + rf * cout))
        w[name + "/kernel"] = rng.uniform(-limit, limit, size=(k,) * ndim + (cin, cout)).astype(np.float32)
        w[name + "/bias"] = np.zeros((cout,), np.float32)
    return w


def _tw(w, name, dtype):
    """Keras kernel (k..., Cin, Cout) -> torch (Cout, Cin, k...)."""
    k = torch.as_tensor(w[name + "/kernel"]).to(dtype)
    nd = k.dim() - 2
    perm = (nd + 1, nd) + tuple(range(nd))
    return k.permute(*perm).contiguous(), torch.as_tensor(w[name + "/bias"]).to(dtype)


# ----------------------------------------------------------------------------------------------
# plain 3D U-Net
# ----------------------------------------------------------------------------------------------

def _batch_norm(y, w, name, training, bn_updates, eps=1e-3, momentum=0.99):
    """Keras 2.x BatchNormalization(axis=1) (create_convolution_block(batch_normalization=True), unet3d/unet.py:103-104):
    training = batch mean / BIASED variance, y = (x - mean) * rsqrt(var + eps) * gamma + beta, and the moving statistics
    move towards the batch values (the variance with Keras' sample-size correction n / (n - (1 + eps))) - collected in
    `bn_updates`; inference = the moving statistics."""
    dims = (0,) + tuple(range(2, y.dim()))
    shape = (1, -1) + (1,) * (y.dim() - 2)
    dt = y.dtype
    gamma, beta = torch.as_tensor(w[name + "/gamma"]).to(dt), torch.as_tensor(w[name + "/beta"]).to(dt)
    mm, mv = torch.as_tensor(w[name + "/moving_mean"]).to(dt), torch.as_tensor(w[name + "/moving_variance"]).to(dt)
    if training:
        mean, var = y.mean(dim=dims), y.var(dim=dims, unbiased=False)
        if bn_updates is not None:
            n = float(y.numel() // y.shape[1])
            bn_updates[name + "/moving_mean"] = (mm.detach() * momentum + mean.detach() * (1 - momentum)).numpy()
            bn_updates[name + "/moving_variance"] = (mv.detach() * momentum +
                                                    var.detach() * (n / (n - (1.0 + eps))) * (1 - momentum)).numpy()
    else:
        mean, var = mm, mv
    return (y - mean.reshape(shape)) * torch.rsqrt(var.reshape(shape) + eps) * gamma.reshape(shape) + beta.reshape(shape)


def bn_params(layers):
    """Keras initial values of the BatchNormalization behind every 3x3 conv block: gamma 1, beta 0, moving_mean 0,
    moving_variance 1."""
    p = {}
    for name, cin, cout, k in layers:
        if k == 3:
            p[name + "/gamma"] = np.ones((cout,), np.float32)
            p[name + "/beta"] = np.zeros((cout,), np.float32)
            p[name + "/moving_mean"] = np.zeros((cout,), np.float32)
            p[name + "/moving_variance"] = np.ones((cout,), np.float32)
    return p


def unet3d_forward(x, w, depth=4, return_logits=False, quant=None, training=False, bn_updates=None):
    """x: [B,Cin,X,Y,Z] torch tensor. `quant` (optional callable) is applied to every stored
    activation and to the weights - used to model bf16 storage for tolerance studies. Weights holding
    '<layer>/gamma' switch the block to Conv -> BatchNormalization -> ReLU (`training` selects batch / moving
    statistics)."""
    dt = x.dtype
    q = quant if quant is not None else (lambda t: t)

    def cb(t, name):
        k, b = _tw(w, name, dt)
        y = F.conv3d(t, q(k), b, padding=1)
        if (name + "/gamma") in w:
            y = _batch_norm(q(y), w, name, training, bn_updates)
        return q(F.relu(y))

    cur = q(x)
    skips = []
    for d in range(depth):
        cur = cb(cur, "enc%da" % d)
        cur = cb(cur, "enc%db" % d)
        skips.append(cur)
        if d < depth - 1:
            cur = F.max_pool3d(cur, 2)
    for d in range(depth - 2, -1, -1):
        if ("up%d/kernel" % d) in w:                   # Deconvolution3D: Keras kernel (2,2,2,Cout,Cin), bias, no activation
            kt = torch.as_tensor(w["up%d/kernel" % d]).to(dt).permute(4, 3, 0, 1, 2).contiguous()
            up = q(F.conv_transpose3d(cur, q(kt), torch.as_tensor(w["up%d/bias" % d]).to(dt), stride=2))
        else:
            up = cur.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
        cur = torch.cat([up, skips[d]], dim=1)          # unet3d/unet.py:61: [up, skip]
        cur = cb(cur, "dec%da" % d)
        cur = cb(cur, "dec%db" % d)
    k, b = _tw(w, "final", dt)
    logits = F.conv3d(cur, q(k), b)
    if return_logits:
        return logits
    return torch.sigmoid(logits)


# ----------------------------------------------------------------------------------------------
# losses / metrics (metrics.py:11-32)
# ----------------------------------------------------------------------------------------------

def dice_coefficient(t, p, smooth=1.0):
    tf_, pf = t.reshape(-1), p.reshape(-1)
    inter = (tf_ * pf).sum()
    return (2.0 * inter + smooth) / (tf_.sum() + pf.sum() + smooth)


def dice_coefficient_loss(t, p):
    return -dice_coefficient(t, p)


def vod_coefficient(t, p, smooth=1.0):
    tb = (t.reshape(-1) > 0.5).to(p.dtype)
    pb = (p.reshape(-1) > 0.5).to(p.dtype)
    inter = (tb * pb).sum()
    union = tb.sum() + pb.sum() - inter
    return (inter + smooth) / (union + smooth)


def binary_accuracy(t, p):
    # Keras: mean(equal(y_true, round(y_pred)))
    return (t == torch.round(p)).to(p.dtype).mean()


def binary_crossentropy(t, p):
    """Keras K.binary_crossentropy (TF backend, from_logits=False): p clipped to [eps, 1-eps], turned into a logit and
    passed to sigmoid_cross_entropy_with_logits == -(t log p + (1-t) log(1-p)) on the clipped p. Keras forms eps and
    1 - eps in the tensor's dtype, float32: the upper bound is float32(1) - float32(1e-7) = 1 - 2^-23 (0.99999988),
    not 1 - 1e-7 - which caps the per-voxel loss at 15.94, not 16.12."""
    lo = float(np.float32(1e-7))
    hi = float(np.float32(1.0) - np.float32(1e-7))
    pc = p.clamp(lo, hi)
    return -(t * torch.log(pc) + (1.0 - t) * torch.log1p(-pc))


def dice_and_xent(t, p, xent_weight=1.0, weight_mask=None, dist_sigma=None):
    """metrics.py:68-78 (dice_and_xent + weighted_cross_entropy_loss); with `dist_sigma` the mask is the distance map of
    dice_and_xent_mask (metrics.py:89-95): weight = exp(-mask / dist_sigma)."""
    xent = binary_crossentropy(t, p)
    if weight_mask is not None:
        w = torch.exp(-weight_mask / dist_sigma) if dist_sigma is not None else weight_mask
        xent = w.reshape(xent.shape) * xent
    return dice_coefficient_loss(t, p) + xent_weight * xent.mean()


def dice_grad_closed_form(t, p, smooth=1.0):
    """dL/dp for L = -dice: -(2 t S - (2I+smooth)) / S^2, S = sum t + sum p + smooth."""
    I = (t * p).sum()
    S = t.sum() + p.sum() + smooth
    return -(2.0 * t * S - (2.0 * I + smooth)) / (S * S)


# ----------------------------------------------------------------------------------------------
# Keras-2 Adam (App. A.9)
# ----------------------------------------------------------------------------------------------

def keras_adam_step(p, g, m, v, iterations, lr, beta_1=0.9, beta_2=0.999, eps=1e-7):
    """In-place on numpy float32 arrays; `iterations` = number of steps already taken."""
    t = iterations + 1
    lr_t = lr * np.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)
    m[...] = beta_1 * m + (1.0 - beta_1) * g
    v[...] = beta_2 * v + (1.0 - beta_2) * g * g
    p[...] = p - np.float32(lr_t) * m / (np.sqrt(v) + np.float32(eps))


def train_step(forward_fn, x, t, w, adam_state, lr, dtype=torch.float32, loss_fn=None):
    """Generic fwd + loss (soft Dice unless `loss_fn(t, p)` is given) + bwd + Keras-Adam for any of the forward
    restatements."""
    names = sorted(w.keys())
    params = {n: torch.tensor(w[n], dtype=dtype, requires_grad=True) for n in names}
    xt, tt = torch.as_tensor(x).to(dtype), torch.as_tensor(t).to(dtype)
    p = forward_fn(xt, params)
    loss = (loss_fn or dice_coefficient_loss)(tt, p)
    loss.backward()
    out = {"loss": float(loss.detach()), "grads": {}, "pred": p.detach().numpy()}
    it = adam_state.setdefault("iterations", 0)
    for n in names:
        if params[n].grad is None:           # e.g. BatchNormalization moving statistics (the caller applies bn_updates)
            out["grads"][n] = np.zeros_like(w[n])
            continue
        g = params[n].grad.detach().to(torch.float32).numpy()
        out["grads"][n] = g
        keras_adam_step(w[n], g, adam_state.setdefault("m/" + n, np.zeros_like(w[n])),
                        adam_state.setdefault("v/" + n, np.zeros_like(w[n])), it, lr)
    adam_state["iterations"] = it + 1
    return out


def unet3d_train_step(x, t, w, adam_state, lr, depth=4, dtype=torch.float32, quant=None):
    """One fwd + Dice loss + bwd + Keras-Adam update. Mutates w/adam_state. Returns dict of scalars+grads.
    `quant` (e.g. a bf16 round trip) is applied to stored activations/weights in the forward pass; autograd
    treats it as identity, so ReLU masks and max-pool routing are decided on the quantised values."""
    names = sorted(w.keys())
    params = {n: torch.tensor(w[n], dtype=dtype, requires_grad=True) for n in names}
    xt = torch.as_tensor(x).to(dtype)
    tt = torch.as_tensor(t).to(dtype)
    bn_updates = {}
    p = unet3d_forward(xt, params, depth=depth, quant=quant, training=True, bn_updates=bn_updates)
    loss = dice_coefficient_loss(tt, p)
    loss.backward()
    out = {"loss": float(loss.detach()), "binary_accuracy": float(binary_accuracy(tt, p.detach())),
           "vod_coefficient": float(vod_coefficient(tt, p.detach())), "grads": {}, "pred": p.detach().numpy()}
    it = adam_state.setdefault("iterations", 0)
    for n in names:
        if n in bn_updates:                  # non-trainable moving statistics: no gradient, updated by the forward pass
            out["grads"][n] = np.zeros_like(w[n])
            w[n][...] = bn_updates[n].astype(np.float32)
            continue
        g = params[n].grad.detach().to(torch.float32).numpy()
        out["grads"][n] = g
        m = adam_state.setdefault("m/" + n, np.zeros_like(w[n]))
        v = adam_state.setdefault("v/" + n, np.zeros_like(w[n]))
        keras_adam_step(w[n], g, m, v, it, lr)
    adam_state["iterations"] = it + 1
    return out


# ----------------------------------------------------------------------------------------------
# Isensee-2017 3D (isensee2017.py:39-79) — forward only
# ----------------------------------------------------------------------------------------------

def isensee3d_layers(depth=5, n_base_filters=16, n_segmentation_levels=1, in_channels=1, n_labels=1):
    layers = []
    c = in_channels
    filt = []
    for l in range(depth):
        f = n_base_filters * (2 ** l)
        filt.append(f)
        layers += [("l%d_in" % l, c, f, 3), ("l%d_ctx1" % l, f, f, 3), ("l%d_ctx2" % l, f, f, 3)]
        c = f
    for l in range(depth - 2, -1, -1):
        f = filt[l]
        layers += [("u%d_up" % l, c, f, 3), ("u%d_loc1" % l, 2 * f, f, 3), ("u%d_loc2" % l, f, f, 1)]
        c = f
        if l < n_segmentation_levels:
            layers.append(("u%d_seg" % l, f, n_labels, 1))
    return layers


def isensee3d_norm_params(layers):
    p = {}
    for name, cin, cout, k in layers:
        if not name.endswith("_seg"):
            p[name + "/gamma"] = np.ones((cout,), np.float32)
            p[name + "/beta"] = np.zeros((cout,), np.float32)
    return p


def _instance_norm(t, gamma, beta, eps=1e-3):
    # keras_contrib InstanceNormalization(axis=1): stddev = sqrt(var_biased) + eps (App. A.6)
    dims = tuple(range(2, t.dim()))
    mean = t.mean(dim=dims, keepdim=True)
    std = t.var(dim=dims, unbiased=False, keepdim=True).sqrt() + eps
    shape = (1, -1) + (1,) * (t.dim() - 2)
    return (t - mean) / std * gamma.reshape(shape) + beta.reshape(shape)


def instance_norm_lrelu_backward_closed_form(x, gy, gamma, beta, eps=1e-3, slope=0.3, chan_scale=None):
    """Closed form the CUDA kernels implement (csrc/bandwidth.cu instnorm_bwd_*): gradient of
    y = LeakyReLU(gamma * (x - mean) / (sqrt(var) + eps) + beta) [* chan_scale] w.r.t. x, gamma, beta, for
    x, gy of shape [N, C, ...] (statistics per sample and channel over the trailing axes). With s = sigma + eps,
    xh = (x - mean) / s and g = gy * chan_scale * (z > 0 ? 1 : slope):
        dgamma = sum g * xh,  dbeta = sum g,
        dx = (gamma / s) * (g - mean(g) - xh * (s / sigma) * mean(g * xh))
    (eps sits on the standard deviation, so the variance term carries s / sigma instead of 1).
    float64 numpy; checked against autograd in tests/test_oracle_pinning.py."""
    x = np.asarray(x, np.float64)
    gy = np.asarray(gy, np.float64)
    axes = tuple(range(2, x.ndim))
    bshape = (1, -1) + (1,) * (x.ndim - 2)
    gam = np.asarray(gamma, np.float64).reshape(bshape)
    bet = np.asarray(beta, np.float64).reshape(bshape)
    mean = x.mean(axis=axes, keepdims=True)
    sigma = np.sqrt(x.var(axis=axes, keepdims=True))
    s = sigma + eps
    xh = (x - mean) / s
    z = gam * xh + bet
    g = gy * np.where(z > 0, 1.0, slope)
    if chan_scale is not None:
        g = g * np.asarray(chan_scale, np.float64).reshape(x.shape[:2] + (1,) * (x.ndim - 2))
    dgamma = (g * xh).sum(axis=(0,) + axes)
    dbeta = g.sum(axis=(0,) + axes)
    m1 = g.mean(axis=axes, keepdims=True)
    m2 = (g * xh).mean(axis=axes, keepdims=True)
    dx = gam / s * (g - m1 - xh * (s / sigma) * m2)
    return dx, dgamma, dbeta


def upsampled_conv_parity_weights(w):
    """Groundwork for convolving the nearest-upsampled decoder source at COARSE resolution (headroom).
    For x_up = UpSampling3D(2)(x) and a 3x3x3 'same' conv with torch-layout weights w [Cout, Cin, 3, 3, 3], the output
    voxel at fine position 2i + p (parity p in {0,1} per axis) only sees coarse voxels i + o with
        p = 0:  tap 0 -> o = -1,  taps 1, 2 -> o = 0          p = 1:  taps 0, 1 -> o = 0,  tap 2 -> o = +1
    so per parity class the 27 taps collapse into 2x2x2 summed taps. Returns wc [2,2,2][Cout, Cin, 3, 3, 3] indexed by
    parity, laid out over coarse offsets o + 1 in {0,1,2} per axis (zeros where a class does not reach).
    y[..., 2i+p] == conv3d(x, wc[p], padding=1)[..., i] exactly (zero padding agrees at both resolutions)."""
    w = np.asarray(w)
    tap_to_offset = {0: (0, 1, 1), 1: (1, 1, 2)}          # parity -> coarse offset index (o + 1) of taps 0, 1, 2
    wc = np.zeros((2, 2, 2) + w.shape, w.dtype)
    for px in range(2):
        for py in range(2):
            for pz in range(2):
                for tx in range(3):
                    for ty in range(3):
                        for tz in range(3):
                            ox, oy, oz = tap_to_offset[px][tx], tap_to_offset[py][ty], tap_to_offset[pz][tz]
                            wc[px, py, pz, :, :, ox, oy, oz] += w[:, :, tx, ty, tz]
    return wc


def _isensee_forward(x, w, depth, n_segmentation_levels, return_logits, nd, drop=None):
    """The Isensee graph on channels-first tensors, 3D (nd = 3) or 2D (nd = 2). `drop`: {level: scale [B,C]} of the
    SpatialDropout between the two context convs (training with dropout; None = identity)."""
    dt = x.dtype
    conv = F.conv3d if nd == 3 else F.conv2d

    def up2(t):
        for ax in range(2, 2 + nd):
            t = t.repeat_interleave(2, ax)
        return t

    def cb(t, name, stride=1, k=3):
        kk, b = _tw(w, name, dt)
        if k == 3 and stride == 2:
            t = F.pad(t, (0, 1) * nd)              # TF SAME, even input: pad_before 0, pad_after 1 (App. A.3)
            y = conv(t, kk, b, stride=2)
        elif k == 3:
            y = conv(t, kk, b, padding=1)
        else:
            y = conv(t, kk, b)
        g = torch.as_tensor(w[name + "/gamma"]).to(dt)
        be = torch.as_tensor(w[name + "/beta"]).to(dt)
        return F.leaky_relu(_instance_norm(y, g, be), 0.3)

    cur = x
    outs = []
    for l in range(depth):
        inc = cb(cur, "l%d_in" % l, stride=1 if l == 0 else 2)
        c1 = cb(inc, "l%d_ctx1" % l)
        if drop is not None and l in drop:
            c1 = c1 * torch.as_tensor(drop[l]).to(dt).reshape(c1.shape[:2] + (1,) * nd)
        ctx = cb(c1, "l%d_ctx2" % l)                        # dropout: identity at rate 0 / inference
        cur = inc + ctx
        outs.append(cur)
    segs = {}
    for l in range(depth - 2, -1, -1):
        up = cb(up2(cur), "u%d_up" % l)
        cat = torch.cat([outs[l], up], dim=1)           # isensee2017.py:62 / unet/isensee.py:62: [skip, up]
        cur = cb(cb(cat, "u%d_loc1" % l), "u%d_loc2" % l, k=1)
        if l < n_segmentation_levels:
            kk, b = _tw(w, "u%d_seg" % l, dt)
            segs[l] = conv(cur, kk, b)
    out = None
    for l in reversed(range(n_segmentation_levels)):
        out = segs[l] if out is None else out + segs[l]
        if l > 0:
            out = up2(out)
    return out if return_logits else torch.sigmoid(out)


def isensee3d_forward(x, w, depth=5, n_segmentation_levels=1, return_logits=False, drop=None):
    return _isensee_forward(x, w, depth, n_segmentation_levels, return_logits, 3, drop)


def isensee2d_layers(depth=5, n_base_filters=16, n_heads=1, in_channels=5, n_labels=1):
    """isensee2017_model (fetal_net/model/unet/isensee.py:14-86). `n_heads` = heads that reach the output: 1 under the
    reference default summation=False (only segmentation_layers[0] is connected, isensee.py:81-82; Keras keeps no
    weights for the unconnected coarser heads), n_segmentation_levels with summation=True."""
    return isensee3d_layers(depth, n_base_filters, n_heads, in_channels, n_labels)


def isensee2d_forward(x, w, depth=5, n_heads=1, return_logits=False, drop=None):
    """x: [B,H,W,D] (slices-as-channels) -> [B,H,W,1]: Permute((3,1,2)), the 2D graph, Permute((2,3,1))."""
    out = _isensee_forward(x.permute(0, 3, 1, 2), w, depth, n_heads, return_logits, 2, drop)
    return out.permute(0, 2, 3, 1)


# ----------------------------------------------------------------------------------------------
# 2D / 2.5D U-Net (model/unet/unet.py:49-85) — forward only
# ----------------------------------------------------------------------------------------------

def unet2d_layers(depth=4, n_base_filters=32, in_channels=6, n_labels=1, deconvolution=False):
    return unet3d_layers(depth, n_base_filters, in_channels, n_labels, deconvolution)


def library_dropout_scales(n, rate, seed):
    """The library's SpatialDropout keep/scale factors (dropout_scale_kernel, csrc/bandwidth.cu): splitmix64 of
    (seed, index) -> u in [0,1) from the top 24 bits -> keep ? 1/(1-rate) : 0. Keras draws its masks from TF's RNG,
    which cannot be reproduced; this restates OUR generator so that a training step WITH dropout can be checked."""
    i = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * i
        h = (h ^ (h >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        h = (h ^ (h >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        h ^= h >> np.uint64(31)
    u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.where(u >= np.float32(rate), np.float32(1.0) / (np.float32(1.0) - np.float32(rate)), np.float32(0.0)) \
        .astype(np.float32)


def bf16_round(t):
    """Storage model of the CUDA path: every stored activation / weight is rounded to bfloat16 on the way forward and
    the gradient flowing back through the same point is rounded too (the cast is differentiable: autograd casts the
    incoming gradient to bfloat16 and back) - the CUDA path keeps activations AND their gradients in bf16."""
    return t.to(torch.bfloat16).to(t.dtype)


def unet2d_forward(x, w, depth=4, return_logits=False, drop=None, quant=None, training=False, bn_updates=None):
    """x: [B,H,W,D] (slices-as-channels, Keras input layout) -> [B,H,W,n_labels]. `drop` (training with
    SpatialDropout2D, unet/unet.py:60-61,76-77): {'enc<d>' / 'dec<d>': scale [B,C]} applied behind the first block of
    the level."""
    dt = x.dtype
    q = quant if quant is not None else (lambda t: t)
    cur = q(x.permute(0, 3, 1, 2))                   # Permute((3,1,2))
    skips = []

    def cb(t, name):
        k, b = _tw(w, name, dt)
        y = F.conv2d(t, q(k), b, padding=1)
        if (name + "/gamma") in w:
            y = _batch_norm(q(y), w, name, training, bn_updates)
        return q(F.relu(y))

    def dr(t, key):
        if drop is None or key not in drop:
            return t
        return q(t * torch.as_tensor(drop[key]).to(dt)[:, :, None, None])

    for d in range(depth):
        cur = cb(dr(cb(cur, "enc%da" % d), "enc%d" % d), "enc%db" % d)
        skips.append(cur)
        if d < depth - 1:
            cur = F.max_pool2d(cur, 2)
    for d in range(depth - 2, -1, -1):
        if ("up%d/kernel" % d) in w:                   # Deconvolution2D: Keras kernel (2,2,Cout,Cin)
            kt = torch.as_tensor(w["up%d/kernel" % d]).to(dt).permute(3, 2, 0, 1).contiguous()
            up = q(F.conv_transpose2d(cur, q(kt), torch.as_tensor(w["up%d/bias" % d]).to(dt), stride=2))
        else:
            up = cur.repeat_interleave(2, 2).repeat_interleave(2, 3)
        cur = torch.cat([up, skips[d]], dim=1)
        cur = cb(dr(cb(cur, "dec%da" % d), "dec%d" % d), "dec%db" % d)
    k, b = _tw(w, "final", dt)
    logits = F.conv2d(cur, k, b)
    out = logits if return_logits else torch.sigmoid(logits)
    return out.permute(0, 2, 3, 1)                   # Permute((2,3,1))


class OracleModel:
    """Keras-Model duck type over the oracle forward (used by tests and bench's cpu_baseline leg)."""

    def __init__(self, weights, input_shape, depth=4, dtype=torch.float32):
        self.w = weights
        self.depth = depth
        self.dtype = dtype
        self.output_shape = (None, 1) + tuple(input_shape[1:])

    def predict(self, data):
        with torch.no_grad():
            x = torch.as_tensor(np.asarray(data)).to(self.dtype)
            return unet3d_forward(x, self.w, depth=self.depth).to(torch.float32).numpy()
