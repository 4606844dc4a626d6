"""Training-step plumbing for BASELINE.json configs[4] (host side): the reference's losses
(``config/focal_loss.py:5-52``, ``train.py:55-91``) and one optimizer step with gradient accumulation under
``DistributedDataParallel`` (one process per GPU, NCCL all-reduce of the gradients once per ``accum_steps`` micro-steps --
the reference's single-process ``nn.DataParallel``, ``train.py:249-251``, has no equivalent of ``no_sync``).

The forward/backward arithmetic of a training step is the PyTorch composition of the drop-in modules (autograd, BatchNorm
batch statistics per chunk, dropout) with the native Haar kernel and its hand-written adjoint; native backward kernels
are the next row, not part of this file.
"""
import contextlib

import torch
import torch.nn.functional as F


def binary_focal_loss(logits, target, alpha=0.25, gamma=2.0, reduction="mean"):
    """``BinaryFocalLoss.forward`` (config/focal_loss.py:23-52): alpha_t * (1 - p_t)^gamma * BCE(sigmoid(logits), target),
    evaluated from the logits (``binary_cross_entropy_with_logits``) so that it stays finite for |logit| > 16 where the
    reference's ``binary_cross_entropy(sigmoid(x))`` clamps its log at -100."""
    target = target.to(logits.dtype)
    p = torch.sigmoid(logits)
    ce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")
    p_t = p * target + (1 - p) * (1 - target)
    loss = (alpha * target + (1 - alpha) * (1 - target)) * (1 - p_t) ** gamma * ce
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    return loss


def orthogonal_loss(space, freq):
    """``train.py:55-67``: squared Frobenius norm of the off-diagonal cross-covariance of the L2-normalised features."""
    d = space.shape[1]
    cov = F.normalize(space, p=2, dim=1).t() @ F.normalize(freq, p=2, dim=1)
    off = cov * (1 - torch.eye(d, device=cov.device, dtype=cov.dtype))
    return off.pow(2).sum() / (d * (d - 1))


def combined_loss(outputs, labels, epoch, max_epochs, alpha=0.25, gamma=2.0):
    """``train.py:69-91`` with the focal criterion: classification loss, plus the ramped orthogonality term after 20 % of
    the epochs."""
    labels = labels.view(-1, 1).float()
    cls = binary_focal_loss(outputs["logits"], labels, alpha, gamma)
    if epoch < 0.2 * max_epochs or "space" not in outputs:      # sfe_only / sfe_mwt return no branch features: classification loss only
        return cls                                              # (ablation.py:70-76, eval.py:158-160)
    lam = min(1.0, (epoch - 0.2 * max_epochs) / (0.5 * max_epochs))
    return cls + lam * orthogonal_loss(outputs["space"], outputs["freq"])


def freeze_unused_(model, ablation):
    """Switch off ``requires_grad`` for every parameter outside the ablation mode's path (dynamic: ``dama`` + ``classifier``;
    sfe_only: ``sfe_cls``; sfe_mwt: ``sfe`` + ``mwt`` + ``fusion_gate`` + ``classifier`` -- model.py:83-161), keeping the
    reference's frozen backbone prefix (sfe.py:115-119).  DistributedDataParallel then registers exactly the tensors that
    receive gradients, so no ``find_unused_parameters`` graph walk is needed per step.  Returns the number of trainable
    scalars (= fp32 gradient elements all-reduced per optimizer step)."""
    used = {"dynamic": ("dama.", "classifier."), "sfe_only": ("sfe_cls.",),
            "sfe_mwt": ("sfe.", "mwt.", "fusion_gate.", "classifier.")}[ablation]
    total = 0
    for name, p in model.named_parameters():
        on_path = name.startswith(used)
        if ablation == "sfe_only" and name.startswith("sfe_cls.feat_map."):
            on_path = False                      # cls mode returns mlp_head(token 0): feat_map is never evaluated (sfe.py:163-166)
        if ablation in ("dynamic", "sfe_mwt") and (name.startswith("dama.sfe.mlp_head.") or name.startswith("sfe.mlp_head.")):
            on_path = False                      # feature-map mode never evaluates mlp_head (sfe.py:168-173)
        if ".efficient_net._fc." in name:
            on_path = False                      # b0's ImageNet classifier: extract_features stops before it (sfe.py:148)
        if not on_path:
            p.requires_grad_(False)
        if p.requires_grad:
            total += p.numel()
    return total


def train_step(model, micro_batches, optimizer, ablation="dynamic", batch_size=8, epoch=0, max_epochs=1, alpha=0.25, gamma=2.0):
    """One optimizer step over ``len(micro_batches)`` accumulated micro-steps (reference loop ``train.py:93-115`` with
    ``accum_steps = len(micro_batches)``).  ``micro_batches`` is a sequence of ``(frames[B,K,3,H,W], labels[B])``.
    With a ``DistributedDataParallel`` model the gradient all-reduce runs on the LAST micro-step only (``no_sync`` on the
    others).  Returns the mean loss of the micro-steps as a float (one host sync per optimizer step, not per micro-step)."""
    accum = len(micro_batches)
    if accum == 0:
        raise ValueError("train_step needs at least one micro-batch")
    optimizer.zero_grad(set_to_none=True)
    total = None
    for i, (frames, labels) in enumerate(micro_batches):
        sync = i == accum - 1
        ctx = contextlib.nullcontext() if sync or not hasattr(model, "no_sync") else model.no_sync()
        with ctx:
            out = model(frames, batch_size, ablation)
            loss = combined_loss(out, labels, epoch, max_epochs, alpha, gamma) / accum
            loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    optimizer.step()
    return float(total)
