#!/usr/bin/env python
"""EXPERIMENT: step time under acnn_set_conv_mtiles (-1 auto, 1, 2) -- checks the size rules."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from assembled_cnn_b200 import _lib
from assembled_cnn_b200.hparams import params_from_flags
from assembled_cnn_b200.model_fns import Model, Trainer
from bench import MODEL_FLAGS, TRAIN_FLAGS, synth_batch

lib = _lib.load()
params = params_from_flags(batch_size=256, **MODEL_FLAGS, **TRAIN_FLAGS)
model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv",
              anti_alias_filter_size=3)
for mode in (-1, 1, 2, -1):
    lib.acnn_set_conv_mtiles(mode)
    tr = Trainer(model, params, 224, 224, use_cuda_graph=True)
    x, y = synth_batch(tr.input_batch, 224, 1234)
    x, y = x.cuda(), y.cuda()
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(15):
        tr.train_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    print("conv_mtiles=%2d : %.3f ms/step" % (mode, e0.elapsed_time(e1) / 15), flush=True)
lib.acnn_set_conv_mtiles(-1)
