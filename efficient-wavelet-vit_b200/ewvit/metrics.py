"""On-device evaluation metrics: what ``eval.py:174-192`` computes with scikit-learn on host lists,
computed by one kernel on the per-video scores that already live on the GPU."""
import torch

from . import ops
from ._lib import EwvitError, check, load

KEYS = ("auc", "eer", "eer_threshold", "accuracy", "precision", "recall", "f1", "ap")


def binary_metrics_tensor(scores, labels, out=None):
    """scores [n] fp32 CUDA probabilities, labels [n] (0 real / 1 fake) -> fp32 CUDA tensor [12] (layout: include/ewvit.h).
    Asynchronous on the current stream: no host synchronisation."""
    ops._require_cuda(scores, "scores")
    scores = scores.reshape(-1)
    if scores.dtype != torch.float32 or not scores.is_contiguous():
        raise EwvitError("binary_metrics: scores must be a contiguous fp32 CUDA tensor")
    labels = labels.reshape(-1).to(device=scores.device, dtype=torch.int32).contiguous()
    if labels.numel() != scores.numel():
        raise EwvitError("binary_metrics: one label per score")
    if out is None:
        out = torch.empty(12, dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        check(load().ewvit_binary_metrics_fwd(scores.data_ptr(), labels.data_ptr(), scores.numel(), out.data_ptr(), ops._stream()),
              "ewvit_binary_metrics_fwd")
    return out


def binary_metrics(scores, labels):
    """The dict ``eval.evaluate`` builds (eval.py:176-189), from device tensors; one device -> host copy of 48 bytes."""
    v = binary_metrics_tensor(scores, labels).cpu().tolist()
    m = dict(zip(KEYS, v[:8]))
    tn, fp, fn, tp = (int(round(x)) for x in v[8:12])
    m["conf_matrix"] = [[tn, fp], [fn, tp]]
    return m
