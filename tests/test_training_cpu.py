"""Host side of BASELINE configs[4] (training step): the reference's losses restated, and gradient accumulation under
DistributedDataParallel over gloo (world_size 2) -- all-reduce on the last micro-step only, replicas stay identical."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from ewvit.training import binary_focal_loss, combined_loss, orthogonal_loss, train_step


def _reference_focal(logits, target, alpha=0.25, gamma=2, reduction="mean"):
    """config/focal_loss.py:23-52, restated line by line"""
    p = torch.sigmoid(logits)
    ce = F.binary_cross_entropy(p, target, reduction="none")
    p_t = p * target + (1 - p) * (1 - target)
    loss = (alpha * target + (1 - alpha) * (1 - target)) * (1 - p_t) ** gamma * ce
    return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss


@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_focal_loss_matches_reference_formula(reduction):
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(16, 1, generator=g) * 3
    target = (torch.rand(16, 1, generator=g) > 0.5).float()
    a = binary_focal_loss(logits, target, reduction=reduction)
    b = _reference_focal(logits, target, reduction=reduction)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    big = torch.tensor([[40.0], [-40.0]])
    assert torch.isfinite(binary_focal_loss(big, torch.tensor([[0.0], [1.0]]))).all()


def test_orthogonal_and_combined_loss():
    g = torch.Generator().manual_seed(4)
    s, f = torch.randn(6, 8, generator=g), torch.randn(6, 8, generator=g)
    sn, fn = F.normalize(s, dim=1), F.normalize(f, dim=1)
    cov = sn.t() @ fn
    ref = torch.norm(cov * (1 - torch.eye(8)), p="fro") ** 2 / (8 * 7)          # train.py:55-67
    assert torch.allclose(orthogonal_loss(s, f), ref, rtol=1e-5)
    out = {"logits": torch.randn(6, 1, generator=g), "space": s, "freq": f}
    y = torch.tensor([0, 1, 1, 0, 1, 0])
    early = combined_loss(out, y, epoch=0, max_epochs=10)
    late = combined_loss(out, y, epoch=7, max_epochs=10)                         # lambda = min(1, (7 - 2) / 5) = 1
    assert torch.allclose(late - early, orthogonal_loss(s, f), rtol=1e-5, atol=1e-7)


class _Tiny(torch.nn.Module):
    """stands in for DeepfakeDetector: same call signature and output dict, an unused branch like model.mwt/sfe_cls"""

    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Linear(12, 8)
        self.freq = torch.nn.Linear(12, 8)
        self.head = torch.nn.Linear(8, 1)
        self.unused = torch.nn.Linear(4, 4)

    def forward(self, x, batch_size=8, ablation="dynamic"):
        v = x.flatten(2).mean(dim=1)
        s, f = self.enc(v), self.freq(v)
        return {"logits": self.head(s + f), "space": s, "freq": f, "fused": s + f}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _micro(rank, step):
    g = torch.Generator().manual_seed(100 + 10 * rank + step)
    return torch.randn(2, 3, 12, generator=g), (torch.rand(2, generator=g) > 0.5).long()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.parallel.DistributedDataParallel(_Tiny(), find_unused_parameters=True)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        loss = train_step(model, [_micro(rank, 0), _micro(rank, 1)], opt, epoch=5, max_epochs=10)
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        q.put((rank, loss, flat.tolist(), model.module.unused.weight.grad is None))
    finally:
        dist.destroy_process_group()


def test_ddp_accumulation_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, l0, w0, u0), (r1, l1, w1, u1) = res
    assert w0 == w1, "replicas diverged: the gradients were not all-reduced"
    assert u0 and u1, "the unused branch must not receive gradients"
    # single-process reference: average of the two ranks' accumulated gradients == one step on all four micro-batches
    torch.manual_seed(0)
    ref = _Tiny()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1)
    opt.zero_grad()
    for rank in range(2):
        for step in range(2):
            x, y = _micro(rank, step)
            (combined_loss(ref(x), y, 5, 10) / 2 / 2).backward()
    opt.step()
    flat = torch.cat([p.detach().flatten() for p in ref.parameters()])
    assert torch.allclose(torch.tensor(w0), flat, atol=1e-6)
