#!/usr/bin/env python
"""EXPERIMENT: in-graph cost of every op kind.  Re-captures the training-step CUDA graph with all
ops of one kind removed (results are garbage, timing is not: no kernel has data-dependent work)
and reports the step-time delta against the full graph.  Unlike the eager per-op event timings of
tools/profile_step.py this includes exactly the launch / dependency latency the op costs inside
the graph, which is what fusing it away would recover.

    python tools/exp_skip.py [--steps 20]
"""
import argparse
import os
os.environ.setdefault('ACNN_NATIVE_PLAN', '0')   # per-op hooks live in the Python executor (same plan, same launches)
import sys
from collections import Counter

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from assembled_cnn_b200.hparams import params_from_flags
from assembled_cnn_b200.model_fns import Model, Trainer
from bench import MODEL_FLAGS, TRAIN_FLAGS, synth_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
B = args.batch
params = params_from_flags(batch_size=B, **MODEL_FLAGS, **TRAIN_FLAGS)
model = Model(50, num_classes=1001, resnet_version=2, use_sk_block=True, anti_alias_type="sconv",
              anti_alias_filter_size=3)


def step_ms(skip):
    tr = Trainer(model, params, 224, 224, use_cuda_graph=True)
    rt = tr.rt
    saved = {}
    for k in skip:
        saved[k] = getattr(rt, "op_" + k)
        setattr(rt, "op_" + k, lambda op: None)
    x, y = synth_batch(tr.input_batch, 224, 1234)
    x, y = x.cuda(), y.cuda()
    try:
        for _ in range(3):
            tr.train_step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tr.train_step(x, y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps
    finally:
        for k in skip:
            delattr(rt, "op_" + k)


tr0 = Trainer(model, params, 224, 224, use_cuda_graph=False)
plan = tr0.rt.plan
counts = Counter(op.kind for op in plan.forward + plan.backward + plan.update)
base = step_ms(())
print("full graph: %.3f ms/step" % base)
rows = []
for kind in sorted(counts):
    if kind in ("prep_weights", "pack_input", "mix_labels", "softmax_ce", "sgd"):
        continue
    ms = step_ms((kind,))
    rows.append((base - ms, kind, counts[kind], ms))
    print("  without %-18s (%3d ops): %.3f ms  -> in-graph cost %.3f ms" % (kind, counts[kind], ms, base - ms), flush=True)
print("--- sorted")
for d, kind, n, ms in sorted(rows, reverse=True):
    print("%-18s %3d ops  %.3f ms  (%.1f us/op)" % (kind, n, d, 1e3 * d / n))
print("sum of in-graph costs %.2f ms of %.2f" % (sum(r[0] for r in rows), base))
