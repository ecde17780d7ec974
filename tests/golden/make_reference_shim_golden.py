#!/usr/bin/env python
"""Golden vectors produced by EXECUTING the reference's own model code (nets/resnet_model.py,
nets/blocks.py, nets/model_helper.py under /root/reference) through the TF-1.14 API stand-in in
tests/golden/tf1_shim/ (TensorFlow itself cannot be installed here).  Run in the build container:

    python tests/golden/make_reference_shim_golden.py            # rewrites reference_shim_golden.json

What a golden entry holds, per model configuration: the variable names / shapes / initializer kinds in
the reference's creation order (a digest + the count), and for seeded inputs and seeded variable
values the logits in training and in inference mode plus digests of the updated BN moving statistics.
tests/test_reference_shim_golden_cpu.py replays the oracle on the same seeds against this file; it
does not need /root/reference.

The wrapper arguments below restate functions/model_fns.py:141-203 (class Model: num_filters=64,
kernel_size=7, conv_stride=2, first_pool 3/2, block sizes by depth, block strides by version);
importing that module itself would pull in the Estimator run loop.
"""
import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "reference_shim_golden.json")

# configuration name -> (constructor flags of functions/model_fns.py:Model, use_resnet_d, batch, size)
CONFIGS = {
    "assemble_r50_rv2_sk_sconv": (dict(resnet_size=50, resnet_version=2, use_sk_block=True,
                                       anti_alias_type="sconv", anti_alias_filter_size=3), False, 4, 64),
    "assemble_r50_rv2_sk_sconv_resnet_d": (dict(resnet_size=50, resnet_version=2, use_sk_block=True,
                                                anti_alias_type="sconv", anti_alias_filter_size=3), True, 2, 64),
    "vanilla_r50_rv1": (dict(resnet_size=50, resnet_version=1), False, 2, 64),
    "r50_rv1_resnet_d_sk_aa_zero_gamma": (dict(resnet_size=50, resnet_version=1, use_sk_block=True,
                                               anti_alias_type="sconv", anti_alias_filter_size=3,
                                               zero_gamma=True), True, 2, 64),
    "r50_rv1_se_proj_aa5": (dict(resnet_size=50, resnet_version=1, use_se_block=True,
                                 anti_alias_type="proj_sconv", anti_alias_filter_size=5), False, 2, 64),
    "r50_rv2_se_no_downsample": (dict(resnet_size=50, resnet_version=2, use_se_block=True,
                                      no_downsample=True), False, 2, 64),
    "assemble_r152_rv2_sk_sconv": (dict(resnet_size=152, resnet_version=2, use_sk_block=True,
                                        anti_alias_type="sconv", anti_alias_filter_size=3), False, 2, 64),
    "r101_rv1_resnet_d": (dict(resnet_size=101, resnet_version=1), True, 2, 64),
}
BLOCK_SIZES = {1: {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3], 200: [3, 24, 36, 3]}}


def block_sizes(resnet_size, resnet_version):
    """functions/model_fns.py:96-135 get_block_sizes (version 2 = BigLittle table)."""
    if resnet_version == 2:
        return {50: [3, 4, 6, 3], 101: [4, 8, 18, 3], 152: [5, 12, 30, 3]}[resnet_size]
    return BLOCK_SIZES[1][resnet_size]


def seeded_value(index, name, shape):
    """Deterministic value of the index-th variable (shared with the test): well-conditioned
    statistics so that training- and inference-mode outputs are both O(1)."""
    g = torch.Generator().manual_seed(1000003 * (index + 1) + len(name))
    leaf = name.rsplit("/", 1)[1]
    if leaf == "moving_variance":
        return 0.5 + torch.rand(shape, generator=g)
    if leaf == "gamma":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf in ("beta", "moving_mean", "bias"):
        return 0.1 * torch.randn(shape, generator=g)
    fan_in = 1
    for d in shape[:-1]:
        fan_in *= d
    return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5


def seeded_input(batch, size):
    return torch.randn(batch, size, size, 3, generator=torch.Generator().manual_seed(20240923))


def digest(t):
    t = t.detach().double().flatten()
    return {"sum": float(t.sum()), "abs_sum": float(t.abs().sum()), "first": float(t[0]), "last": float(t[-1])}


def names_digest(order):
    h = hashlib.sha256()
    for name, shape, kind, trainable in order:
        h.update(("%s|%s|%s|%d\n" % (name, ",".join(map(str, shape)), kind, trainable)).encode())
    return h.hexdigest()


def run_reference(flags, use_resnet_d, batch, size):
    sys.path.insert(0, os.path.join(HERE, "tf1_shim"))
    sys.path.insert(0, "/root/reference")
    import tensorflow as tf          # the shim
    from nets import resnet_model   # the REAL reference code
    f = dict(flags)
    rv = f.get("resnet_version", 1)
    strides = [2, 2, 1, 2] if rv == 2 else [1, 2, 2, 2]
    if f.pop("no_downsample", False):
        strides[-1] = 1
    size_ = f.pop("resnet_size")

    def make():
        return resnet_model.Model(resnet_size=size_, bottleneck=True, num_classes=1001, num_filters=64,
                                  kernel_size=7, conv_stride=2, first_pool_size=3, first_pool_stride=2,
                                  block_sizes=block_sizes(size_, rv), block_strides=strides, **f)

    # pass 1: discover the variables (names, shapes, initializers, order)
    tf.reset()
    make()(tf.Tensor(seeded_input(batch, size)), training=False, use_resnet_d=use_resnet_d)
    order = list(tf.variables.order)
    values = {n: seeded_value(i, n, s) for i, (n, s, _, _) in enumerate(order)}
    out = {"num_variables": len(order), "names_sha256": names_digest(order),
           "first_names": [o[0] for o in order[:3]], "last_names": [o[0] for o in order[-2:]],
           "zero_init_gammas": sum(1 for o in order if o[2] == "zeros" and o[0].endswith("/gamma")),
           "batch": batch, "size": size, "use_resnet_d": use_resnet_d}
    x = seeded_input(batch, size)
    # pass 2: inference mode with the seeded values
    tf.reset(values)
    y = make()(tf.Tensor(x), training=False, use_resnet_d=use_resnet_d)
    out["eval_logits"] = digest(y.t)
    out["eval_logits_row0_head"] = [float(v) for v in y.t[0, :8]]
    # pass 3: training mode (batch statistics + moving-average updates)
    tf.reset(values)
    y = make()(tf.Tensor(x), training=True, use_resnet_d=use_resnet_d)
    out["train_logits"] = digest(y.t)
    mm = torch.cat([tf.variables.vars[o[0]].t.flatten() for o in order if o[0].endswith("moving_mean")])
    mv = torch.cat([tf.variables.vars[o[0]].t.flatten() for o in order if o[0].endswith("moving_variance")])
    out["moving_mean_after"] = digest(mm)
    out["moving_variance_after"] = digest(mv)
    return out, order


if __name__ == "__main__":
    gold = {}
    for name, (flags, d, b, s) in CONFIGS.items():
        gold[name], _ = run_reference(flags, d, b, s)
        print(name, gold[name]["num_variables"], gold[name]["eval_logits"]["abs_sum"])
    json.dump(gold, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)
