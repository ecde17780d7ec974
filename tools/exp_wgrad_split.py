#!/usr/bin/env python
"""EXPERIMENT: training-step time (c3, CUDA graph) against the wgrad split-K rule: the round-1 "two
waves of CTAs" rule (0) vs the cost model with a per-CTA fixed cost of N pipeline stages
(acnn_set_wgrad_overhead_stages).  Results are unchanged beyond fp32 summation order.

    python tools/exp_wgrad_split.py [--steps 15]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from assembled_cnn_b200 import _lib
from assembled_cnn_b200.hparams import params_from_flags
from assembled_cnn_b200.model_fns import Model, Trainer
from bench import MODEL_FLAGS, TRAIN_FLAGS, synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=15)
ap.add_argument("--values", default="0,4,8,12,16,24,32,0")
args = ap.parse_args()
lib = _lib.load()
params = params_from_flags(batch_size=args.batch, **MODEL_FLAGS, **TRAIN_FLAGS)
model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv",
              anti_alias_filter_size=3)
for v in [int(t) for t in args.values.split(",")]:
    lib.acnn_set_wgrad_overhead_stages(v)
    tr = Trainer(model, params, 224, 224, use_cuda_graph=True)     # fresh capture under the knob
    x, y = synth_batch(tr.input_batch, 224, 1234)
    x, y = x.cuda(), y.cuda()
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.train_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    print("wgrad_overhead_stages=%2d : %.3f ms/step   loss %s"
          % (v, e0.elapsed_time(e1) / args.steps, tr.train_step(x, y).tolist()), flush=True)
    del tr
