#!/usr/bin/env python
"""EXPERIMENT: step time of the BASELINE c3 training step under the library's tuning knobs (none of
them changes results): programmatic dependent launch off / light kernels only / all, fused SK attention chains on / off,
wgrad pixels per stage 64 / 128.

    python tools/exp_knobs.py [--steps 15]
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
args = ap.parse_args()
lib = _lib.load()
B = args.batch
params = params_from_flags(batch_size=B, **MODEL_FLAGS, **TRAIN_FLAGS)
model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv",
              anti_alias_filter_size=3)


def step_ms():
    tr = Trainer(model, params, 224, 224, use_cuda_graph=True)     # fresh capture under the knobs
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
    loss = tr.train_step(x, y).tolist()
    return e0.elapsed_time(e1) / args.steps, loss


for pdl in (0, 2, 1):
    lib.acnn_set_pdl(pdl)
    ms, loss = step_ms()
    print("pdl=%d : %.3f ms/step   loss %s" % (pdl, ms, loss), flush=True)
lib.acnn_set_pdl(0)
for fused, pix in ((0, 64), (0, 128), (1, 64)):
    lib.acnn_set_sk_fc_fused(fused)
    lib.acnn_set_wgrad_pixels(pix)
    ms, loss = step_ms()
    print("sk_fc_fused=%d wgrad_pixels=%3d : %.3f ms/step   loss %s" % (fused, pix, ms, loss),
          flush=True)
lib.acnn_set_sk_fc_fused(-1)
lib.acnn_set_wgrad_pixels(0)
