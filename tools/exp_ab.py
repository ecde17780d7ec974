#!/usr/bin/env python
"""EXPERIMENT harness: A/B the c3 training step under one library tuning knob (an `acnn_set_*`
function of include/acnn.h; none of them changes results beyond fp32 summation order).  One Trainer
(= one captured CUDA graph) per value, then the graphs are replayed INTERLEAVED (A B A B ...), so the
slow drift of a power-capped GPU's clocks cancels instead of biasing one arm.

    python tools/exp_ab.py --knob acnn_set_wgrad_overhead_stages --values 0,16 [--rounds 6 --steps 10]
    python tools/exp_ab.py --configs "acnn_set_conv_mtiles=1,acnn_set_conv_split_epilogue=2;acnn_set_conv_mtiles=-1"
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from assembled_cnn_b200 import _lib
from assembled_cnn_b200.hparams import params_from_flags
from assembled_cnn_b200.model_fns import Model, Trainer
from bench import MODEL_FLAGS, TRAIN_FLAGS, synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--knob", default="")
ap.add_argument("--values", default="")
ap.add_argument("--configs", default="", help="';'-separated arms, each a ','-separated list of knob=value")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
lib = _lib.load()
if args.configs:
    arms = [[(kv.split("=")[0], int(kv.split("=")[1])) for kv in arm.split(",")]
            for arm in args.configs.split(";")]
else:
    arms = [[(args.knob, int(v))] for v in args.values.split(",")]
values = [" ".join("%s(%d)" % kv for kv in arm) for arm in arms]
params = params_from_flags(batch_size=args.batch, **MODEL_FLAGS, **TRAIN_FLAGS)
trainers = []
for v, arm in zip(values, arms):
    prevs = [(k, getattr(lib, k)(val)) for k, val in arm]
    model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv",
                  anti_alias_filter_size=3)
    tr = Trainer(model, params, 224, 224, use_cuda_graph=True)     # captured under the knob
    x, y = synth_batch(tr.input_batch, 224, 1234)
    x, y = x.cuda(), y.cuda()
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
    trainers.append((v, tr, x, y))
    for k, pv in reversed(prevs):
        getattr(lib, k)(pv)
times = {v: [] for v in values}
for r in range(args.rounds):
    for v, tr, x, y in trainers:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tr.train_step(x, y)
        e1.record()
        torch.cuda.synchronize()
        times[v].append(e0.elapsed_time(e1) / args.steps)
for v in values:
    t = times[v]
    print("%s: median %.3f ms/step  (min %.3f max %.3f over %d interleaved rounds of %d steps)"
          % (v, statistics.median(t), min(t), max(t), len(t), args.steps), flush=True)
