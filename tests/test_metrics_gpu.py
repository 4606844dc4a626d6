"""GPU parity of the on-device evaluation metrics (row f-4) against scikit-learn, the reference's own implementation
(eval.py:79-94 `calculate_eer`, eval.py:174-192).  Floating point: the kernel accumulates in double like numpy and returns
fp32, tolerance 1e-6 absolute; thresholds and the integer confusion matrix must match exactly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sklearn_reference(labels, scores):
    from sklearn.metrics import (accuracy_score, average_precision_score, confusion_matrix, f1_score, precision_score, recall_score,
                                 roc_auc_score, roc_curve)
    fpr, tpr, thresholds = roc_curve(labels, scores)                       # eval.py:87-92
    fnr = 1 - tpr
    i = np.nanargmin(np.absolute(fnr - fpr))
    binary = [1 if p >= 0.5 else 0 for p in scores]                        # eval.py:172
    return {"auc": roc_auc_score(labels, scores), "eer": fpr[i], "eer_threshold": thresholds[i],
            "accuracy": accuracy_score(labels, binary), "precision": precision_score(labels, binary, zero_division=0),
            "recall": recall_score(labels, binary, zero_division=0), "f1": f1_score(labels, binary, zero_division=0),
            "ap": average_precision_score(labels, scores), "conf_matrix": confusion_matrix(labels, binary, labels=[0, 1]).tolist()}


@pytest.mark.parametrize("n,seed,ties,sep", [(64, 0, False, 1.0), (64, 1, True, 1.0), (300, 2, False, 0.3), (1000, 3, True, 2.0),
                                             (8192, 4, False, 0.5), (7, 5, False, 1.0), (2, 6, False, 5.0), (513, 7, True, 0.0)])
def test_binary_metrics_match_sklearn(n, seed, ties, sep):
    from ewvit.metrics import binary_metrics
    g = torch.Generator().manual_seed(seed)
    labels = (torch.rand(n, generator=g) < 0.45).int()
    labels[0], labels[-1] = 0, 1                                           # both classes present
    logits = torch.randn(n, generator=g) + sep * (labels.float() * 2 - 1)
    if ties:
        logits = (logits * 4).round() / 4                                  # many equal scores
    scores = torch.sigmoid(logits)
    got = binary_metrics(scores.cuda(), labels.cuda())
    ref = _sklearn_reference(labels.numpy(), scores.numpy())
    assert got["conf_matrix"] == ref["conf_matrix"]
    for k in ("auc", "eer", "accuracy", "precision", "recall", "f1", "ap"):
        assert abs(got[k] - float(ref[k])) <= 1e-6, (k, got[k], ref[k])
    if np.isinf(ref["eer_threshold"]):
        assert np.isinf(got["eer_threshold"])
    else:
        assert got["eer_threshold"] == float(np.float32(ref["eer_threshold"])), (got["eer_threshold"], ref["eer_threshold"])


def test_single_class_is_flagged_and_sizes_are_checked():
    from ewvit import EwvitError
    from ewvit.metrics import binary_metrics
    m = binary_metrics(torch.rand(16).cuda(), torch.ones(16).int().cuda())
    assert np.isnan(m["auc"]) and np.isnan(m["eer"]) and np.isnan(m["ap"]) and m["recall"] >= 0
    with pytest.raises(EwvitError):
        binary_metrics(torch.rand(9000).cuda(), torch.zeros(9000).int().cuda())
    with pytest.raises(EwvitError):
        binary_metrics(torch.rand(8).cuda(), torch.zeros(7).int().cuda())
