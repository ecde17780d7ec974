"""CPU oracle, part 1: the TensorFlow-1.14 op semantics the reference's hot path relies on.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.

PARITY UNPINNED AGAINST TENSORFLOW ITSELF: the reference (clovaai/assembled-cnn) ships no tests or
golden vectors for this path and its arithmetic lives in the un-vendored dependency
tensorflow==1.14.0 (README.md:85), which cannot be installed here (Python 3.12, no network).  This
file restates the published TF-1.14 semantics at the reference's own call sites (file:line below).
What it IS pinned against:
  * hand-computed micro-vectors per rule (tests/test_oracle_known_answers.py);
  * implementations written and validated against TensorFlow by others, rule by rule
    (tests/test_oracle_independent_pins_cpu.py: Hugging Face's port of the TF BiT checkpoints for 'SAME'
    padding, ATen's BatchNorm / cross_entropy(label_smoothing) / SGD / avg_pool2d(count_include_pad=
    False) kernels, scipy.ndimage 'mirror' correlation);
  * the reference's OWN model code executed end to end on those third-party kernels in float64
    (tests/golden/tf1_shim no longer imports this file: its convolution padding is Hugging Face's, its
    batch norm ATen's): the golden vectors of tests/golden/make_reference_shim_golden.py hold
    inference- and training-mode logits, moving statistics, loss and gradient digests of 11
    configurations up to ResNet-152 plus DropBlock through the whole model, and oracle/model.py built
    on this file reproduces them to 1e-6 in float64 (tests/test_reference_shim_golden_cpu.py).

All tensors are torch CPU tensors, activations NHWC (the reference's CPU layout,
nets/resnet_model.py:196-198), conv kernels HWIO, dense kernels [in, out].
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nets/model_helper.py:26
BN_MOMENTUM = 0.997    # nets/model_helper.py:26


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


# ------------------------------------------------------------------------------------------
# nets/model_helper.py:40-64  fixed_padding  /  :67-78  conv2d_fixed_padding
# ------------------------------------------------------------------------------------------
def fixed_padding(x, kernel_size):
    """Zero-pad H and W by (k-1)//2 before and k-1-(k-1)//2 after (model_helper.py:53-63)."""
    pad_total = kernel_size - 1
    pad_beg = pad_total // 2
    pad_end = pad_total - pad_beg
    return F.pad(x, (0, 0, pad_beg, pad_end, pad_beg, pad_end))


def _same_pads(size, k, s):
    """TF 'SAME' padding: total = max((ceil(size/s)-1)*s + k - size, 0); extra goes AFTER."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d(x, w_hwio, strides=1, padding="SAME"):
    """tf.layers.conv2d(use_bias=False) on NHWC input with an HWIO kernel."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    if padding == "SAME":
        pt, pb = _same_pads(x.shape[1], kh, strides)
        pl, pr = _same_pads(x.shape[2], kw, strides)
        x = F.pad(x, (0, 0, pl, pr, pt, pb))
    w = w_hwio.permute(3, 2, 0, 1)  # OIHW
    return _nhwc(F.conv2d(_nchw(x), w, stride=strides))


def conv2d_fixed_padding(x, w_hwio, strides):
    """nets/model_helper.py:67-78: explicit pad + VALID when strided, else SAME."""
    k = w_hwio.shape[0]
    if strides > 1:
        x = fixed_padding(x, k)
    return conv2d(x, w_hwio, strides, "SAME" if strides == 1 else "VALID")


# ------------------------------------------------------------------------------------------
# nets/model_helper.py:26-37  batch_norm  (tf.layers.batch_normalization, fused=True)
# ------------------------------------------------------------------------------------------
def batch_norm(x, gamma, beta, moving_mean, moving_var, training, momentum=BN_MOMENTUM,
               eps=BN_EPS):
    """Returns (y, new_moving_mean, new_moving_var).

    Training: normalise with the batch mean and the BIASED batch variance over (N, H, W);
    moving_var is updated with the UNBIASED variance (TF fused kernel), both as
    moving <- moving * momentum + batch * (1 - momentum)  (UPDATE_OPS, optimizer_setting.py:36-37).
    Inference: normalise with the moving statistics.
    """
    if training:
        n = x.numel() // x.shape[-1]
        mean = x.mean(dim=(0, 1, 2))
        var = x.var(dim=(0, 1, 2), unbiased=False)
        y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
        with torch.no_grad():
            unbiased = var * (n / max(n - 1, 1))
            new_mm = moving_mean * momentum + mean * (1 - momentum)
            new_mv = moving_var * momentum + unbiased * (1 - momentum)
        return y, new_mm, new_mv
    y = (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta
    return y, moving_mean, moving_var


# ------------------------------------------------------------------------------------------
# pooling / resampling
# ------------------------------------------------------------------------------------------
def max_pool_same(x, pool_size=3, strides=2):
    """tf.layers.max_pooling2d(padding='SAME') (nets/resnet_model.py:421-424): the odd pad cell
    goes AFTER (0 before / 1 after for 3x3 s2 on even sizes), padded with -inf."""
    pt, pb = _same_pads(x.shape[1], pool_size, strides)
    pl, pr = _same_pads(x.shape[2], pool_size, strides)
    xp = F.pad(_nchw(x), (pl, pr, pt, pb), value=float("-inf"))
    return _nhwc(F.max_pool2d(xp, pool_size, strides))


def avg_pool_resnet_d(x, strides):
    """resnet_d_projection_shortcut (nets/resnet_model.py:123-128): strided: fixed_padding(2) =
    0 before / 1 after zeros, 2x2 VALID (divide by 4); stride 1: 2x2 SAME, TF excludes the padded
    cells from the divisor."""
    if strides > 1:
        xp = fixed_padding(x, 2)
        return _nhwc(F.avg_pool2d(_nchw(xp), 2, strides))
    pt, pb = _same_pads(x.shape[1], 2, 1)
    pl, pr = _same_pads(x.shape[2], 2, 1)
    xs = F.avg_pool2d(F.pad(_nchw(x), (pl, pr, pt, pb)), 2, 1) * 4.0
    ones = torch.ones(1, 1, x.shape[1], x.shape[2], dtype=x.dtype)
    cnt = F.avg_pool2d(F.pad(ones, (pl, pr, pt, pb)), 2, 1) * 4.0
    return _nhwc(xs / cnt)


def avg_pool_bl(x, strides):
    """bl_projection_shortcut (nets/resnet_model.py:133-138): only when strided: fixed_padding(3)
    = 1/1 zeros then 3x3 VALID; the zeros are counted (always / 9)."""
    if strides > 1:
        xp = fixed_padding(x, 3)
        return _nhwc(F.avg_pool2d(_nchw(xp), 3, strides))
    return x


def upsample2x(x):
    """tf.keras.layers.UpSampling2D((2,2)) nearest (nets/resnet_model.py:499)."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


_BINOMIAL = {1: [1.], 2: [1., 1.], 3: [1., 2., 1.], 4: [1., 3., 3., 1.], 5: [1., 4., 6., 4., 1.],
             6: [1., 5., 10., 10., 5., 1.], 7: [1., 6., 15., 20., 15., 6., 1.]}


def anti_aliased_downsample(x, filt_size=3, stride=2):
    """nets/blocks.py:45-107: REFLECT pad int((f-1)/2) each side, depthwise binomial filter
    (outer product / sum), stride, VALID.  Constant filter: no gradient to it."""
    p = int(1.0 * (filt_size - 1) / 2)
    a = torch.tensor(_BINOMIAL[filt_size], dtype=x.dtype)
    if filt_size == 1:
        return x[:, ::stride, ::stride, :]
    filt = a[:, None] * a[None, :]
    filt = filt / filt.sum()
    c = x.shape[-1]
    xp = F.pad(_nchw(x), (p, p, p, p), mode="reflect")
    w = filt[None, None].repeat(c, 1, 1, 1)
    return _nhwc(F.conv2d(xp, w, stride=stride, groups=c))


def global_avg_pool(x):
    """tf.reduce_mean over H, W (nets/resnet_model.py:561; blocks.py:132,170) -> [B, C]."""
    return x.mean(dim=(1, 2))


# ------------------------------------------------------------------------------------------
# nets/blocks.py:110-154  sk_conv2d   /  :156-184  se_block
# ------------------------------------------------------------------------------------------
def sk_attention(u, fc1_w, bn_gamma, bn_beta, bn_mm, bn_mv, fc2_w, training, momentum):
    """The part of sk_conv2d after conv+BN+ReLU (blocks.py:128-152).

    u: [B,H,W,2f] = relu(bn(conv3x3)).  fc1_w [f, d], fc2_w [d, 2f] (1x1 conv kernels squeezed).
    Returns (v [B,H,W,f], new_mm, new_mv).  softmax over the two contiguous channel halves.
    """
    f = u.shape[-1] // 2
    u0, u1 = u[..., :f], u[..., f:]
    s = (u0 + u1).mean(dim=(1, 2))                       # [B, f]
    z = s @ fc1_w                                         # [B, d]
    z, mm, mv = batch_norm(z[:, None, None, :], bn_gamma, bn_beta, bn_mm, bn_mv, training, momentum)
    z = torch.relu(z[:, 0, 0, :])
    a = z @ fc2_w                                         # [B, 2f]
    att = torch.softmax(torch.stack([a[:, :f], a[:, f:]], 0), dim=0)   # [2, B, f]
    v = u0 * att[0][:, None, None, :] + u1 * att[1][:, None, None, :]
    return v, mm, mv


def se_block(x, w1, w2):
    """blocks.py:156-184: x * sigmoid(W2 relu(W1 mean_HW(x))), no bias, no BN; w1 [C, C/16]."""
    q = x.mean(dim=(1, 2))
    e = torch.sigmoid(torch.relu(q @ w1) @ w2)
    return x * e[:, None, None, :]


# ------------------------------------------------------------------------------------------
# losses/cls_losses.py:28-33 ; utils/data_util.py:97-158 ; run_loop_classification.py:166-179 ;
# nets/optimizer_setting.py:23-38
# ------------------------------------------------------------------------------------------
def softmax_cross_entropy(logits, onehot, label_smoothing=0.0):
    """tf.losses.softmax_cross_entropy(weights=1.0): y' = y(1-eps) + eps/num_classes,
    mean over the batch of -sum_c y'_c log_softmax(logits)_c."""
    num_classes = onehot.shape[1]
    if label_smoothing > 0:
        onehot = onehot * (1.0 - label_smoothing) + label_smoothing / num_classes
    return -(onehot * F.log_softmax(logits.float(), dim=1)).sum(dim=1).mean()


def mixup(x, y, lam1, lam2=None, keep_batch_size=True, y_t=None):
    """utils/data_util.py:97-158 with the Beta(0.2,0.2) draws passed in (TF's RNG is not
    reproducible).  x [2B',H,W,C], y [2B',classes]; lam1/lam2 [B'].
    keep_batch_size=False (mixup_type 1): returns B' mixed examples; True (type 2): 2B'.
    With y_t (knowledge-distillation teacher labels) a third value, the mixed teacher labels, is
    returned -- including the reference's own quirk at utils/data_util.py:154, where the second half
    of a type-2 batch mixes y1 (the SUPERVISED labels) instead of y1_t with the reversed y2_t."""
    b = x.shape[0] // 2
    x1, x2 = x[:b], x[b:]
    y1, y2 = y[:b], y[b:]
    l1x = lam1.view(b, 1, 1, 1)
    l1y = lam1.view(b, 1)
    mx = l1x * x1 + (1.0 - l1x) * x2
    my = l1y * y1 + (1.0 - l1y) * y2
    myt = None
    if y_t is not None:
        y1_t, y2_t = y_t[:b], y_t[b:]
        myt = l1y * y1_t + (1.0 - l1y) * y2_t
    if keep_batch_size:
        l2x = lam2.view(b, 1, 1, 1)
        l2y = lam2.view(b, 1)
        x3 = torch.flip(x2, [0])
        y3 = torch.flip(y2, [0])
        mx = torch.cat([mx, l2x * x1 + (1.0 - l2x) * x3], 0)
        my = torch.cat([my, l2y * y1 + (1.0 - l2y) * y3], 0)
        if y_t is not None:
            y3_t = torch.flip(y2_t, [0])
            myt = torch.cat([myt, l2y * y1 + (1.0 - l2y) * y3_t], 0)     # sic: y1, not y1_t
    if y_t is not None:
        return mx.detach(), my.detach(), myt.detach()
    return mx.detach(), my.detach()


def kd_loss(logits, teacher_labels, kd_temp):
    """nets/run_loop_classification.py:156-162: T^2 * softmax_cross_entropy(logits / T, teacher)
    with teacher = softmax(teacher_logits / T) (:90-93), no label smoothing, mean over the batch."""
    return kd_temp * kd_temp * softmax_cross_entropy(logits / kd_temp, teacher_labels, 0.0)


def generalized_mean_pooling(x, p=3):
    """nets/blocks.py:22-42 (NHWC): N^(-1/p) * (max(sum_hw clip(x, 1e-6, 1e12)^p, 1e-6))^(1/p)."""
    n = x.shape[1] * x.shape[2]
    eps = 1e-6
    xp = x.clamp(eps, 1e12) ** p
    s = xp.sum(dim=(1, 2)).clamp_min(eps)
    return (float(n) ** (-1.0 / p)) * s ** (1.0 / p)


def dropblock_keep_mask(u, keep_prob, block_size, gamma_scale, h, w):
    """nets/blocks.py:191-251, the random part: u = the uniform draws of _bernoulli, shape
    [1, h-bs+1, w-bs+1, c] (ONE mask for the whole batch).  Returns the keep mask [1,h,w,c] (1 = kept)
    and the renormalisation factor size / (sum + 1e-8)."""
    bs = block_size
    br = (bs - 1) // 2
    tl = (bs - 1) - br
    gamma = (1.0 - keep_prob) * (w * h) / (bs ** 2) / ((w - bs + 1) * (h - bs + 1)) * gamma_scale
    m = torch.relu(torch.sign(gamma - u))                       # 1 with probability gamma
    m = F.pad(m, (0, 0, tl, br, tl, br))                        # [1,h,w,c]
    m = F.max_pool2d(m.permute(0, 3, 1, 2), bs, 1, bs // 2).permute(0, 2, 3, 1)   # SAME, odd bs
    keep = 1.0 - m
    factor = keep.numel() / (keep.float().sum() + 1e-8)
    return keep, factor


def dropblock(x, keep_prob, block_size, gamma_scale, u):
    """x * keep_mask * size/sum(keep_mask); identity when keep_prob == 1 or gamma_scale == 0."""
    if (isinstance(keep_prob, float) and keep_prob == 1) or gamma_scale == 0 or u is None:
        return x
    keep, factor = dropblock_keep_mask(u.to(x.dtype), keep_prob, block_size, gamma_scale,
                                       x.shape[1], x.shape[2])
    return x * keep * factor.to(x.dtype)


def ece(conf, pred, label, num_thresholds=10):
    """metric/ece_metric.py:171-298 for one batch: per-bin (correct, confidence sum, count) and the
    expected calibration error sum_b cnt_b/sum(cnt) * |correct_b/(eps+cnt_b) - conf_b/(eps+cnt_b)|."""
    eps = 1e-7
    th = [0.0 - eps] + [(i + 1) / num_thresholds for i in range(num_thresholds - 1)] + [1.0 + eps]
    lo = torch.tensor(th[:num_thresholds]).view(-1, 1)
    hi = torch.tensor(th[1:]).view(-1, 1)
    c = conf.float().view(1, -1)
    inb = (c > lo) & (c <= hi)
    ok = (pred.view(1, -1) == label.view(1, -1)) & inb
    correct = ok.float().sum(1)
    csum = (c * inb.float()).sum(1)
    cnt = inb.float().sum(1)
    val = ((cnt / cnt.sum()) * (correct / (eps + cnt) - csum / (eps + cnt)).abs()).sum()
    return val, correct, csum, cnt


def l2_loss(t):
    """tf.nn.l2_loss: sum(t^2) / 2."""
    return (t.float() ** 2).sum() / 2


def momentum_step(w, acc, g, lr, momentum):
    """tf.train.MomentumOptimizer (non-Nesterov): acc = m*acc + g ; w = w - lr*acc."""
    acc = momentum * acc + g
    return w - lr * acc, acc


# ------------------------------------------------------------------------------------------
# functions/model_fns.py:36-95 learning_rate_with_decay ; :26-33 keep_prob_decay (host scalars)
# ------------------------------------------------------------------------------------------
def learning_rate(step, *, decay_type, batch_size, num_images, base_lr, warmup_epochs=0,
                  train_epochs=None, num_epochs_per_decay=2.0, decay_factor=0.94,
                  end_learning_rate=0.0001, boundary_epochs=(30, 60, 80, 90),
                  decay_rates=(1, 0.1, 0.01, 0.001, 1e-4)):
    initial = base_lr * batch_size / batch_size          # batch_denom == batch_size (:211)
    bpe = num_images / batch_size
    decay_steps = int(bpe * num_epochs_per_decay)
    warmup_steps = int(bpe * warmup_epochs)
    g = step - warmup_steps
    if decay_type == "exponential":
        lr = initial * decay_factor ** math.floor(g / decay_steps)
    elif decay_type == "fixed":
        lr = base_lr
    elif decay_type == "polynomial":
        gg = min(g, decay_steps)
        lr = (initial - end_learning_rate) * (1 - gg / decay_steps) + end_learning_rate
    elif decay_type == "piecewise":
        bounds = [int(bpe * e) for e in boundary_epochs]
        vals = [initial * float(d) for d in decay_rates]
        lr = vals[sum(1 for b in bounds if step > b)]
    elif decay_type == "cosine":
        total = int(bpe * train_epochs) - warmup_steps
        gg = min(max(g, 0), total)
        lr = initial * 0.5 * (1 + math.cos(math.pi * gg / total))
    else:
        raise NotImplementedError(decay_type)
    if warmup_steps > 0 and step < warmup_steps:
        return initial * step / warmup_steps
    return lr
