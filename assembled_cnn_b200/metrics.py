"""Streaming evaluation metrics of the reference's EVAL mode (nets/run_loop_classification.py:
207-227): accuracy, top-5 accuracy (tf.metrics.accuracy / tf.metrics.mean(in_top_k)) and the
expected calibration error of metric/ece_metric.py:171-298 (10 confidence bins, per-bin running
sums).  They consume the logits the CUDA forward produced; a few tiny torch reductions per batch --
host-side bookkeeping, not part of the hot path."""
from __future__ import annotations

import torch


class EceMetric:
    """metric/ece_metric.py: accumulators accuracy_per_bin / confidence_per_bin / count_per_bin."""

    def __init__(self, num_thresholds=10, device="cpu"):
        eps = 1e-7
        th = [0.0 - eps] + [(i + 1) / num_thresholds for i in range(num_thresholds - 1)] + [1.0 + eps]
        self.eps = eps
        self.lo = torch.tensor(th[:num_thresholds], device=device).view(-1, 1)
        self.hi = torch.tensor(th[1:], device=device).view(-1, 1)
        self.correct = torch.zeros(num_thresholds, device=device)
        self.conf = torch.zeros(num_thresholds, device=device)
        self.cnt = torch.zeros(num_thresholds, device=device)

    def update(self, conf, pred, label):
        c = conf.float().view(1, -1)
        inb = (c > self.lo) & (c <= self.hi)
        ok = (pred.view(1, -1) == label.view(1, -1)) & inb
        self.correct += ok.float().sum(1)
        self.conf += (c * inb.float()).sum(1)
        self.cnt += inb.float().sum(1)
        return self.result()

    def result(self):
        acc = self.correct / (self.eps + self.cnt)
        avg = self.conf / (self.eps + self.cnt)
        return ((self.cnt / self.cnt.sum()) * (acc - avg).abs()).sum()


class EvalMetrics:
    """{'accuracy', 'accuracy_top_5', 'ece'} accumulated over the batches of an evaluation."""

    def __init__(self, device="cpu"):
        self.n = 0
        self.top1 = 0.0
        self.top5 = 0.0
        self.ece = EceMetric(device=device)

    def update(self, logits, labels):
        labels = labels.to(logits.device).long()
        prob = torch.softmax(logits.float(), dim=1)
        conf, pred = prob.max(dim=1)
        self.n += labels.numel()
        self.top1 += float((pred == labels).sum())
        self.top5 += float((logits.topk(5, dim=1).indices == labels[:, None]).any(1).sum())
        self.ece.update(conf, pred, labels)
        return self.result()

    def result(self):
        n = max(self.n, 1)
        return {"accuracy": self.top1 / n, "accuracy_top_5": self.top5 / n,
                "ece": float(self.ece.result())}
