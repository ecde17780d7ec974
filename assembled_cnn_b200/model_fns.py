"""The reference's operator surface for the hot path, executing on B200 through libacnn.so.

Mirrors (same names, argument meaning and error behaviour):
  * functions/model_fns.py:138-198   Model(resnet_size, data_format, num_classes, ...)
  * nets/resnet_model.py:305-310     model(inputs, training, reuse, use_resnet_d, keep_prob, return_embedding)
  * functions/model_fns.py:201-239   model_fn_cls(features, labels, mode, params)
  * nets/run_loop_classification.py:60-234  resnet_model_fn (loss assembly, train op)
  * functions/model_fns.py:36-95     learning_rate_with_decay ; :26-33 keep_prob_decay
plus `build_model(**flags)` (the name BASELINE.json uses; the reference has no such function).

The reference builds a TF graph and lets the Estimator run it; here a call executes on the GPU:
`Model.__call__` runs one forward, `Trainer.train_step` one full step (mixup -> forward -> loss ->
backward -> all-reduce -> SGD).  torch tensors are the device-memory container only.
"""
from __future__ import annotations

import os
import math
from collections import namedtuple

import numpy as np
import torch

from . import _lib, dp
from .hparams import DATASETS, DEFAULTS, params_from_flags
from .plan import BLOCK_SIZES, ModelConfig, build_plan
from .metrics import EvalMetrics
from .runtime import Runtime
from .native import NativeModel, NativeRuntime

DEFAULT_VERSION = 1
# dtype: 'bf16' = production path (bf16 storage, fp32 accumulation -- the analogue of the reference's
# fp16 mode); 'fp32' = the reference's default (nets/resnet_model.py:30-33,
# official/utils/flags/_performance.py:29-32): fp32 storage, 3-way bf16-split tcgen05 GEMMs,
# bit-reproducible reductions -- the mode the 1e-3 parity tests run in.  'fp16' has no B200
# counterpart here (bf16 replaces it) and is rejected.
ALLOWED_TYPES = ("bf16", "fp32")

# tf.estimator.ModeKeys values
TRAIN, EVAL, PREDICT = "train", "eval", "infer"

EstimatorSpec = namedtuple("EstimatorSpec", "mode predictions loss train_op eval_metric_ops")


def get_block_sizes(resnet_size, resnet_version=1):
    """functions/model_fns.py:98-135."""
    choices = BLOCK_SIZES[2 if resnet_version == 2 else 1]
    try:
        return choices[resnet_size]
    except KeyError:
        raise ValueError("Could not find layers for selected Resnet size.\n"
                         "Size received: {}; sizes allowed: {}.".format(resnet_size, choices.keys()))


def per_device_batch_size(batch_size, num_gpus):
    """official/utils/misc/distribution_utils.py:48-76: the global batch must divide over the replicas
    (same ValueError text)."""
    if num_gpus <= 1:
        return batch_size
    remainder = batch_size % num_gpus
    if remainder:
        raise ValueError("When running with multiple GPUs, batch size must be a multiple of the number of "
                         "available GPUs. Found {} GPUs with a batch size of {}; try --batch_size={} instead."
                         .format(num_gpus, batch_size, batch_size - remainder))
    return int(batch_size / num_gpus)


def keep_prob_decay(starter_kp, end_kp, decay_steps):
    """functions/model_fns.py:26-33 (linear polynomial decay, no cycle) as a host function."""
    def fn(global_step):
        s = min(global_step, decay_steps)
        return (starter_kp - end_kp) * (1 - s / decay_steps) + end_kp
    return fn


def learning_rate_with_decay(learning_rate_decay_type, batch_size, batch_denom, num_images,
                             num_epochs_per_decay, learning_rate_decay_factor, end_learning_rate,
                             piecewise_lr_boundary_epochs, piecewise_lr_decay_rates, base_lr,
                             warmup_epochs=0, train_epochs=None):
    """functions/model_fns.py:36-95; returns learning_rate_fn(global_step) -> python float."""
    initial = base_lr * batch_size / batch_denom
    bpe = num_images / batch_size
    decay_steps = int(bpe * num_epochs_per_decay)

    def learning_rate_fn(global_step):
        warmup_steps = int(bpe * warmup_epochs)
        g = global_step - warmup_steps
        if learning_rate_decay_type == "exponential":
            lr = initial * learning_rate_decay_factor ** math.floor(g / decay_steps)
        elif learning_rate_decay_type == "fixed":
            lr = base_lr
        elif learning_rate_decay_type == "polynomial":
            gg = min(g, decay_steps)
            lr = (initial - end_learning_rate) * (1 - gg / decay_steps) + end_learning_rate
        elif learning_rate_decay_type == "piecewise":
            bounds = [int(bpe * e) for e in piecewise_lr_boundary_epochs]
            vals = [initial * float(d) for d in piecewise_lr_decay_rates]
            lr = vals[sum(1 for b in bounds if global_step > b)]
        elif learning_rate_decay_type == "cosine":
            total = int(bpe * train_epochs) - warmup_steps
            gg = min(max(g, 0), total)
            lr = initial * 0.5 * (1 + math.cos(math.pi * gg / total))
        else:
            raise NotImplementedError
        if warmup_steps > 0 and global_step < warmup_steps:
            return initial * global_step / warmup_steps
        return lr
    return learning_rate_fn


def _truncated_normal(shape, std, gen):
    w = torch.empty(shape)
    torch.nn.init.trunc_normal_(w, 0.0, std, -2 * std, 2 * std, generator=gen)
    return w


class Model:
    """functions/model_fns.py:138-198 `Model` with ImageNet defaults (64 filters, 7x7/2 stem,
    3x3/2 pool, bottleneck blocks) + nets/resnet_model.py:166-249 argument checks."""

    def __init__(self, resnet_size, data_format=None, num_classes=None,
                 resnet_version=DEFAULT_VERSION, dtype="bf16", no_downsample=False,
                 zero_gamma=False, use_se_block=False, use_sk_block=False, bn_momentum=0.997,
                 embedding_size=0, anti_alias_filter_size=0, anti_alias_type="", pool_type="gap",
                 loss_type="softmax", bl_alpha=2, bl_beta=4, *, seed=42, device="cuda:0",
                 deterministic=None, native=None):
        if data_format not in (None, "channels_last"):
            raise ValueError("this implementation is NHWC only (data_format='channels_last')")
        if dtype not in ALLOWED_TYPES:
            raise ValueError("dtype must be one of: {}".format(ALLOWED_TYPES))
        self.resnet_size = int(resnet_size)
        self.num_classes = num_classes if num_classes is not None else 1001
        self.cfg_kwargs = dict(
            resnet_size=self.resnet_size, num_classes=self.num_classes,
            resnet_version=int(resnet_version), no_downsample=no_downsample, zero_gamma=zero_gamma,
            use_se_block=use_se_block, use_sk_block=use_sk_block, bn_momentum=bn_momentum,
            embedding_size=embedding_size, anti_alias_filter_size=anti_alias_filter_size,
            anti_alias_type=anti_alias_type, pool_type=pool_type, loss_type=loss_type,
            bl_alpha=bl_alpha, bl_beta=bl_beta)
        ModelConfig(**self.cfg_kwargs).validate()         # ValueError / NotImplementedError
        self.block_sizes = get_block_sizes(self.resnet_size, int(resnet_version))
        self.dtype = dtype
        self.data_format = "channels_last"
        self.device = device
        self.seed = seed
        # None: bit-reproducible steps in the fp32 mode only; True: also in bf16 (wgrad and the SE
        # fc GEMMs run without split-K -- slower); every other reduction is ordered in both modes
        self.deterministic = deterministic
        # True (default): the layer plan is built and executed inside libacnn.so through the model-level
        # C ABI (include/acnn_model.h, native.NativeRuntime); False (ACNN_NATIVE_PLAN=0): the Python plan
        # + per-op ctypes executor the lockstep parity tests drive (same plan, text-for-text)
        self.native = (os.environ.get("ACNN_NATIVE_PLAN", "1") != "0") if native is None else bool(native)
        self._runtimes = {}          # (B, H, W, training, use_resnet_d, mixup, ls) -> Runtime
        self._primary = {}           # use_resnet_d -> Runtime owning the parameters
        self._pending_weights = None

    # ---------------------------------------------------------------- runtimes / parameters
    def runtime(self, batch, height, width, *, training, use_resnet_d=False, mixup_type=0,
                label_smoothing=0.0, with_loss=False, use_dropblock=False, kd_temp=0.0) -> Runtime:
        key = (batch, height, width, bool(training), bool(use_resnet_d), mixup_type,
               float(label_smoothing), bool(with_loss or training), bool(use_dropblock and training),
               float(kd_temp) if training else 0.0)
        rt = self._runtimes.get(key)
        if rt is None:
            cfg = ModelConfig(use_resnet_d=bool(use_resnet_d), **self.cfg_kwargs)
            step = dict(training=training, mixup_type=mixup_type, label_smoothing=label_smoothing,
                        with_loss=with_loss, dtype=self.dtype, use_dropblock=use_dropblock,
                        kd_temp=kd_temp)
            prim = self._primary.get(bool(use_resnet_d))
            if self.native:
                rt = NativeRuntime(NativeModel(cfg, batch, height, width,
                                               deterministic=self.deterministic, **step),
                                   self.device, share=prim)
            else:
                rt = Runtime(build_plan(cfg, batch, height, width, **step), self.device, share=prim,
                             deterministic=self.deterministic)
            if prim is None:
                self._primary[bool(use_resnet_d)] = rt
                if self._pending_weights is not None:
                    rt.set_weights(self._pending_weights)
                else:
                    self.init_weights(rt)
            self._runtimes[key] = rt
        return rt

    def init_weights(self, rt: Runtime):
        """The reference's initializers: variance_scaling (truncated normal, fan-in) for conv /
        SK / SE kernels (nets/model_helper.py:77), glorot-uniform dense kernel, zero bias
        (nets/resnet_model.py:595-597), gamma 1 (0 with zero_gamma on block-final BNs), beta 0."""
        gen = torch.Generator().manual_seed(self.seed)
        vals = {}
        for name, p in rt.plan.params.items():
            if p.kind == "conv_kernel":
                kh, kw, cin, cout = p.tf_shape
                std = math.sqrt(1.0 / (kh * kw * cin)) / 0.87962566103423978
                vals[name] = _truncated_normal(p.tf_shape, std, gen)
            elif p.kind == "dense_kernel":
                cin, cout = p.tf_shape
                lim = math.sqrt(6.0 / (cin + cout))
                vals[name] = (torch.rand(cin, cout, generator=gen) * 2 - 1) * lim
            elif p.kind == "gamma":
                vals[name] = torch.zeros(p.tf_shape) if p.zero_init else torch.ones(p.tf_shape)
            else:
                vals[name] = torch.zeros(p.tf_shape)
        for name, p in rt.plan.state.items():
            vals[name] = torch.ones(p.tf_shape) if p.kind == "moving_variance" \
                else torch.zeros(p.tf_shape)
        rt.set_weights(vals)

    def set_weights(self, tf_vars):
        """Load variables given in the reference's names and layouts (HWIO kernels, [in,out]
        dense kernel); applies to every runtime of this model."""
        self._pending_weights = tf_vars
        for rt in self._primary.values():
            rt.set_weights(tf_vars)

    def get_weights(self, use_resnet_d=None):
        if use_resnet_d is None:
            use_resnet_d = getattr(self, "use_resnet_d", False)
        rt = self._primary[bool(use_resnet_d)]
        names = list(rt.plan.params) + list(rt.plan.state)
        return {n: rt.get_tf(n).detach().float().cpu().clone() for n in names}

    # ---------------------------------------------------------------- forward
    def __call__(self, inputs, training, reuse=False, use_resnet_d=None, keep_prob=1.0,
                 return_embedding=False):
        """nets/resnet_model.py:305-599.  inputs: float32 [N,H,W,3] NHWC (CPU or CUDA tensor).
        Returns logits [N, num_classes] fp32 on the GPU (or the pooled embedding [N, C])."""
        # keep_prob == 1.0 (a Python float) is the reference's "DropBlock off" (nets/blocks.py:209);
        # anything else runs the DropBlock plan in training mode (inference ignores it, :205)
        use_db = bool(training) and not (isinstance(keep_prob, float) and keep_prob == 1.0)
        if use_resnet_d is None:
            # the reference's call-time default is False (nets/resnet_model.py:308); build_model()
            # records the flag it was given as the default of this model's calls
            use_resnet_d = getattr(self, "use_resnet_d", False)
        inputs = torch.as_tensor(inputs)
        if inputs.dim() != 4 or inputs.shape[-1] != 3:
            raise ValueError("inputs must be [N, H, W, 3] (NHWC)")
        n, h, w, _ = inputs.shape
        rt = self.runtime(n, h, w, training=False, use_resnet_d=use_resnet_d) if not training \
            else self.runtime(n, h, w, training=True, use_resnet_d=use_resnet_d,
                              use_dropblock=use_db)
        m = rt.plan.meta
        rt.t[m["images"]].copy_(inputs.to(torch.float32), non_blocking=True)
        if training:
            if use_db:
                self._fwd_calls = getattr(self, "_fwd_calls", 0) + 1
                rt.set_hparams(keep_prob=float(keep_prob), step=self._fwd_calls)
            rt.zero_step_buffers()
            fwd = [op for op in rt.plan.forward
                   if op.kind not in ("mix_labels", "softmax_ce", "kd_teacher")]
            rt.run(fwd)
        else:
            rt.run_forward()
        if return_embedding:
            # nets/resnet_model.py:586-590: the (BN-normalised) embedding when embedding_size > 0,
            # else the pooled features
            feat = rt.t[m["embedding"]] if "embedding" in m else rt.t[m["pooled"]]
            return feat.reshape(n, -1).float()
        return rt.t[m["logits"]][:, :self.num_classes]


def build_model(**flags) -> Model:
    """`build_model()` of BASELINE.json's north_star: the Model for a flag set
    (resnet_size, resnet_version, use_sk_block, use_se_block, anti_alias_type, ...).
    `use_resnet_d` is a call-time argument in the reference (nets/resnet_model.py:308); it is
    remembered here as the default of the returned model's calls via `model.use_resnet_d`."""
    use_resnet_d = flags.pop("use_resnet_d", False)
    ctor = {k: flags.pop(k) for k in list(flags) if k in (
        "resnet_size", "data_format", "num_classes", "resnet_version", "dtype", "no_downsample",
        "zero_gamma", "use_se_block", "use_sk_block", "bn_momentum", "embedding_size",
        "anti_alias_filter_size", "anti_alias_type", "pool_type", "loss_type", "bl_alpha",
        "bl_beta", "seed", "device", "deterministic", "native")}
    if flags:
        raise TypeError("build_model: unknown flag(s) %s" % sorted(flags))
    ctor.setdefault("resnet_size", DEFAULTS["resnet_size"])
    model = Model(**ctor)
    model.use_resnet_d = bool(use_resnet_d)
    return model


class Trainer:
    """One data-parallel replica of the training step (resnet_model_fn's TRAIN branch +
    get_train_op): owns the step's static buffers, the LR schedule and the CUDA graphs."""

    def __init__(self, model: Model, params: dict, height=224, width=224, *, use_cuda_graph=True,
                 lam_seed=7):
        p = params
        if p.get("cls_loss_type", "softmax") != "softmax":
            raise NotImplementedError("only cls_loss_type='softmax' is on the hot path")
        self.model = model
        self.p = p
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.local_batch = per_device_batch_size(p["batch_size"], self.world)
        self.mixup_type = int(p.get("mixup_type", 0))
        self.kd_temp = float(p.get("kd_temp", 0) or 0)
        self.use_dropblock = bool(p.get("use_dropblock", False))
        self.rt = model.runtime(self.local_batch, height, width, training=True,
                                use_resnet_d=p.get("use_resnet_d", False),
                                mixup_type=self.mixup_type,
                                label_smoothing=float(p.get("label_smoothing", 0.0)),
                                use_dropblock=self.use_dropblock, kd_temp=self.kd_temp)
        ds = DATASETS[p.get("dataset_name") or "imagenet"]
        # functions/model_fns.py:221-228: keep_prob decays linearly over the whole run
        self.keep_prob_fn = None
        if self.use_dropblock:
            kp0, kp1 = p["dropblock_kp"]
            bpe = ds["num_images"]["train"] / p["batch_size"]
            self.keep_prob_fn = keep_prob_decay(kp0, kp1, int(p["train_epochs"] * bpe))
        self.learning_rate_fn = learning_rate_with_decay(
            learning_rate_decay_type=p["learning_rate_decay_type"], batch_size=p["batch_size"],
            batch_denom=p["batch_size"], num_images=ds["num_images"]["train"],
            num_epochs_per_decay=p["num_epochs_per_decay"],
            learning_rate_decay_factor=p["learning_rate_decay_factor"],
            end_learning_rate=p["end_learning_rate"],
            piecewise_lr_boundary_epochs=p["piecewise_lr_boundary_epochs"],
            piecewise_lr_decay_rates=p["piecewise_lr_decay_rates"],
            base_lr=p["base_learning_rate"], train_epochs=p["train_epochs"],
            warmup_epochs=p["lr_warmup_epochs"])
        self.global_step = 0
        self.rng = np.random.default_rng(lam_seed)
        self.loss_scale = float(p.get("loss_scale", 1) or 1)
        self.rt.loss_scale = self.loss_scale
        self.use_graph = use_cuda_graph
        self._graphs = None
        m = self.rt.plan.meta
        self.images_buf = self.rt.t[m["images"]]
        self.labels_buf = self.rt.t[m["labels"]]
        self.lam1_buf = self.rt.t[m["lam1"]] if "lam1" in m else None
        self.lam2_buf = self.rt.t[m["lam2"]] if "lam2" in m else None
        self.teacher_buf = self.rt.t[m["teacher_logits"]] if "teacher_logits" in m else None
        # hyper-parameters reach the device through a RING of pinned host buffers (one per step in
        # flight, guarded by an event): the host may run several steps ahead of the GPU, and a
        # single buffer would be overwritten before its asynchronous copy has executed
        self._hp_ring = [torch.zeros(8, dtype=torch.float32).pin_memory() for _ in range(4)]
        self._hp_events = [None] * len(self._hp_ring)
        self._loss_slot = self.rt.slot_view(m["loss"])[:3 if self.kd_temp > 0 else 2]
        self._buckets = self._segments = None
        if self.world > 1:
            self._buckets = dp.grad_buckets(self.rt.plan)
            self._segments = dp.backward_segments(self.rt.plan, self._buckets)
        # input double-buffering: prefetch() copies the NEXT batch host->device on a side stream
        # while the current step computes; train_step() then takes it with a device-side copy
        self._copy_stream = torch.cuda.Stream(self.rt.dev)
        self._staged = None
        self._consumed = None
        self._stage_imgs = None
        self._stage_labs = None

    @property
    def input_batch(self):
        """Examples the input pipeline must deliver per replica and step (2x for mixup type 1,
        functions/input_fns.py:98-100)."""
        return self.rt.plan.meta["input_batch"]

    def _capture(self):
        rt = self.rt
        # warm-up launch outside capture (cudaFuncSetAttribute, driver entry points, ...); it is
        # a real forward/backward, so the BN moving statistics it updated are put back
        state_backup = rt.state.clone()
        self._fwd_bwd()
        torch.cuda.synchronize()
        rt.state.copy_(state_backup)
        s = torch.cuda.Stream(rt.dev)
        s.wait_stream(torch.cuda.current_stream(rt.dev))
        # one graph for forward + backward on a single GPU; with data parallelism the backward is cut
        # at the gradient-bucket boundaries (the all-reduces are issued between the graph replays)
        fns = [self._fwd_bwd]
        if self.world > 1:
            bwd = rt.plan.backward
            fns = []
            for k, (a, b) in enumerate(self._segments):
                if k == 0:
                    fns.append(lambda ops=bwd[a:b]: (rt.run_forward(), rt.run(ops)))
                else:
                    fns.append(lambda ops=bwd[a:b]: rt.run(ops))
        graphs = []
        with torch.cuda.stream(s):
            for fn in fns + [lambda: rt.run(rt.plan.update)]:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    fn()
                graphs.append(g)
        torch.cuda.current_stream(rt.dev).wait_stream(s)
        self._graphs = graphs

    def _fwd_bwd(self):
        rt = self.rt
        rt.run_forward()
        # (a second stream for the wgrad GEMMs was measured to give nothing under a CUDA graph:
        # every kernel already spans the GPU; Runtime.run(..., overlap_wgrad=True) keeps the option)
        rt.run(rt.plan.backward, overlap_wgrad=os.environ.get("ACNN_OVERLAP_WGRAD", "0") == "1")

    def prefetch(self, images, labels):
        """Start the host->device copy of the next step's inputs (pinned host tensors) on a side
        stream; the following train_step(None, None) consumes them.  Lets the PCIe transfer of
        step i+1 overlap the compute of step i."""
        if self._stage_imgs is None:
            self._stage_imgs = torch.empty_like(self.images_buf)
            self._stage_labs = torch.empty_like(self.labels_buf)
        cs = self._copy_stream
        # the staging buffers may still be read by the previous step's device-side copy (and only
        # by that: waiting for the whole main stream would serialise the transfer behind the step)
        if self._consumed is not None:
            cs.wait_event(self._consumed)
        with torch.cuda.stream(cs):
            self._stage_imgs.copy_(images, non_blocking=True)
            self._stage_labs.copy_(labels, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._staged = ev

    def train_step(self, images=None, labels=None, lam1=None, lam2=None, teacher_logits=None,
                   keep_prob=None):
        """images fp32 [input_batch,H,W,3] and int32 labels [input_batch] (pinned host or device),
        or None to consume the batch given to prefetch().  teacher_logits fp32 [input_batch, classes]
        when kd_temp > 0 (the second half of the reference's KD label tensor).  keep_prob overrides
        the DropBlock schedule for this step.
        Returns the device tensor [cross_entropy, l2_loss(, kd_loss)] of this replica (read it with
        .tolist() -- that read is the only host sync of the step)."""
        rt = self.rt
        if images is None:
            if self._staged is None:
                raise ValueError("train_step() without inputs needs a preceding prefetch()")
            torch.cuda.current_stream(rt.dev).wait_event(self._staged)
            self._staged = None
            images, labels = self._stage_imgs, self._stage_labs
        self.images_buf.copy_(images, non_blocking=True)
        self.labels_buf.copy_(labels, non_blocking=True)
        if images is self._stage_imgs:
            self._consumed = torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(rt.dev))
        if self.mixup_type:
            n = self.input_batch // 2
            if lam1 is None:
                lam1 = torch.from_numpy(self.rng.beta(0.2, 0.2, n).astype(np.float32))
            self.lam1_buf.copy_(torch.as_tensor(lam1, dtype=torch.float32), non_blocking=True)
            if self.mixup_type == 2:
                if lam2 is None:
                    lam2 = torch.from_numpy(self.rng.beta(0.2, 0.2, n).astype(np.float32))
                self.lam2_buf.copy_(torch.as_tensor(lam2, dtype=torch.float32), non_blocking=True)
        if self.kd_temp > 0:
            if teacher_logits is None:
                raise ValueError("kd_temp > 0: train_step needs teacher_logits")
            self.teacher_buf.copy_(torch.as_tensor(teacher_logits, dtype=torch.float32),
                                   non_blocking=True)
        lr = self.learning_rate_fn(self.global_step)
        slot = self.global_step % len(self._hp_ring)
        if self._hp_events[slot] is not None:
            self._hp_events[slot].synchronize()          # its copy of 4 steps ago has executed
        hp = self._hp_ring[slot]
        hp[0] = lr
        hp[1] = self.p["momentum"]
        hp[2] = self.p["weight_decay"]
        hp[3] = 1.0 / (self.world * self.loss_scale)
        if keep_prob is None:
            keep_prob = self.keep_prob_fn(self.global_step) if self.keep_prob_fn else 1.0
        hp[4] = keep_prob
        hp.view(torch.int32)[5] = self.global_step & 0x7fffffff      # Philox counter of the masks
        rt.hp.copy_(hp, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(rt.dev))
        self._hp_events[slot] = ev
        if self.use_graph and self._graphs is None:
            self._capture()
        if self.world == 1:
            if self.use_graph:
                self._graphs[0].replay()
            else:
                self._fwd_bwd()
        else:
            self._fwd_bwd_allreduce()
        if self.use_graph:
            self._graphs[-1].replay()
        else:
            rt.run(rt.plan.update)
        if self.world > 1:
            self.sync_moving_statistics()
        self.global_step += 1
        self.last_lr = lr
        self.last_keep_prob = keep_prob
        return self._loss_slot

    # ---------------------------------------------------------------- data parallelism
    def sync_moving_statistics(self):
        """MirroredStrategy keeps the BN moving statistics as mirrored variables whose per-replica
        updates are MEAN-aggregated (official/utils/misc/distribution_utils.py:24-45)."""
        dp.average_moving_statistics(self.rt.state, self.world)

    def _fwd_bwd_allreduce(self):
        """forward + backward with the gradient all-reduce overlapped: the backward is cut into
        segments (each its own CUDA graph), and as soon as a segment has produced the last gradient
        of a bucket that bucket is all-reduced asynchronously on NCCL's stream while the next
        segment computes (assembled_cnn_b200/dp.py)."""
        rt = self.rt
        works, k = [], 0
        for ev in dp.schedule(self._buckets, self._segments):
            if ev[0] == "run":
                if self.use_graph:
                    self._graphs[k].replay()
                else:
                    if k == 0:
                        rt.run_forward()
                    rt.run(rt.plan.backward[ev[1]:ev[2]])
                k += 1
            else:
                works.append(dp.all_reduce_bucket(rt.grads, ev[1], ev[2], async_op=True))
        for w in works:
            w.wait()            # stream-level wait: the update runs after every bucket


_TRAINERS = {}


def l2_loss(rt: Runtime, weight_decay: float):
    """weight_decay * sum over the decayed variables of |v|^2 / 2 (run_loop_classification.py:
    166-177), from the flat master buffer and its per-256-element decay flags.  TRAIN gets the same
    number from the SGD kernel; this is the EVAL-mode path (a reduction over 42 M floats, once per
    evaluation batch)."""
    w = rt.params.view(-1, 256)
    return 0.5 * weight_decay * (w[rt.decay_flags.bool()[:w.shape[0]]].double() ** 2).sum().float()


def predictions_of(logits):
    """nets/run_loop_classification.py:126-130: the `predictions` dict of the EstimatorSpec."""
    return {"classes": logits.argmax(dim=1), "probabilities": torch.softmax(logits, dim=1),
            "probabilities_sigmoid": torch.sigmoid(logits)}


def model_fn_cls(features, labels, mode, params):
    """functions/model_fns.py:201-239 -> nets/run_loop_classification.py:60-234.

    features: {'image': float32 [N,H,W,3]} (or the tensor itself in PREDICT mode), labels: int32 [N]
    -- or, with kd_temp > 0, the reference's float KD label tensor [N, 2*num_classes] = one-hot
    labels ++ teacher logits (nets/run_loop_classification.py:86-93).
    TRAIN executes one training step and returns the spec with its loss; EVAL returns loss
    (cross-entropy + L2, as the reference) and predictions of a forward pass with moving statistics
    plus the accuracy / top-5 / ECE metrics; PREDICT only predictions.
    """
    if int(params["resnet_size"]) < 50:
        assert not params.get("use_dropblock")
        assert not params.get("use_se_block")
        assert not params.get("use_sk_block")
        assert not params.get("use_resnet_d")
    p = params_from_flags(**{k: v for k, v in params.items() if k in DEFAULTS})
    ds = DATASETS[p.get("dataset_name") or "imagenet"]
    images = features["image"] if isinstance(features, dict) else features
    images = torch.as_tensor(images)
    key = tuple(sorted((k, str(v)) for k, v in p.items())) + (tuple(images.shape[1:3]),)
    entry = _TRAINERS.get(key)
    if entry is None:
        model = Model(p["resnet_size"], p["data_format"], num_classes=ds["num_classes"],
                      resnet_version=p["resnet_version"], zero_gamma=p["zero_gamma"],
                      use_se_block=p["use_se_block"], use_sk_block=p["use_sk_block"],
                      no_downsample=p["no_downsample"],
                      anti_alias_filter_size=p["anti_alias_filter_size"],
                      anti_alias_type=p["anti_alias_type"], bn_momentum=p["bn_momentum"],
                      embedding_size=p["embedding_size"], pool_type=p["pool_type"],
                      bl_alpha=p["bl_alpha"], bl_beta=p["bl_beta"], dtype=p["dtype"],
                      loss_type=p["cls_loss_type"])
        entry = _TRAINERS[key] = {"model": model, "trainer": None}
    model = entry["model"]

    if mode == PREDICT:
        logits = model(images, False, False, use_resnet_d=p["use_resnet_d"])
        return EstimatorSpec(mode, predictions_of(logits), None, None, None)
    teacher_logits = None
    if mode != PREDICT and p.get("kd_temp", 0) > 0:
        kd_labels = torch.as_tensor(labels, dtype=torch.float32)
        if kd_labels.dim() != 2 or kd_labels.shape[1] != 2 * ds["num_classes"]:
            raise ValueError("kd_temp > 0: labels must be [N, 2*num_classes] (one-hot ++ teacher "
                             "logits, nets/run_loop_classification.py:86-93)")
        onehot, teacher_logits = kd_labels.split(ds["num_classes"], dim=1)
        labels = onehot.argmax(dim=1).to(torch.int32)
    if mode == TRAIN:
        if entry["trainer"] is None:
            entry["trainer"] = Trainer(model, p, images.shape[1], images.shape[2])
        tr = entry["trainer"]
        loss = tr.train_step(images, torch.as_tensor(labels, dtype=torch.int32),
                             teacher_logits=teacher_logits)
        m = tr.rt.plan.meta
        logits = tr.rt.t[m["logits"]][:, :model.num_classes]
        return EstimatorSpec(mode, predictions_of(logits), loss.sum(), tr, None)
    if mode == EVAL:
        n, h, w, _ = images.shape
        rt = model.runtime(n, h, w, training=False, use_resnet_d=p["use_resnet_d"],
                           label_smoothing=p["label_smoothing"], with_loss=True)
        m = rt.plan.meta
        rt.t[m["images"]].copy_(images.to(torch.float32), non_blocking=True)
        rt.t[m["labels"]].copy_(torch.as_tensor(labels, dtype=torch.int32), non_blocking=True)
        rt.run_forward()
        logits = rt.t[m["logits"]][:, :model.num_classes]
        lab = rt.t[m["labels"]].long()
        pred = predictions_of(logits)
        # eval_metric_ops are streaming in the reference (tf.metrics.*): they accumulate over the
        # batches of one evaluation; reset with model_fn_cls.reset_metrics(params)
        em = entry.setdefault("metrics", EvalMetrics(device=logits.device))
        metrics = em.update(logits, lab)
        # loss = cross_entropy + l2_loss in EVAL too (nets/run_loop_classification.py:166-179)
        loss = rt.slot_view(m["loss"])[0] + l2_loss(rt, p["weight_decay"])
        if teacher_logits is not None:
            t = torch.softmax(teacher_logits.to(logits.device) / p["kd_temp"], dim=1)
            loss = loss + p["kd_temp"] ** 2 * -(t * torch.log_softmax(logits / p["kd_temp"], 1)
                                                ).sum(1).mean()
        return EstimatorSpec(mode, pred, loss, None, metrics)
    raise ValueError("unknown mode %r" % (mode,))


def reset_metrics(params=None):
    """Start a new evaluation: drop the streaming accuracy / ECE accumulators."""
    for entry in _TRAINERS.values():
        entry.pop("metrics", None)


model_fn_cls.reset_metrics = reset_metrics
