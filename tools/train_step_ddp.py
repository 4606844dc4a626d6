"""BASELINE.json configs[4]: one training step (forward + backward, BinaryFocalLoss + orthogonality term, accum_steps = 2,
Adam) under DistributedDataParallel, one process per GPU, for the three ablation modes.

    python tools/train_step_ddp.py [--frames 8] [--videos 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/train_step_ddp.py

Per rank x = randn(videos, frames, 3, 224, 224) per micro-step (K reduced from 300 -- train-mode
activations are ~85 MB per frame), batch_size 8.  The arithmetic of a training step is the PyTorch composition of the drop-in
modules (autograd, BatchNorm batch statistics per chunk, dropout) with the native Haar kernel and its adjoint
(ewvit/training.py; row f-2 is 'partial': no native backward kernels).  What this tool measures is the multi-GPU side of
configs[4]: step time (CUDA events, max over ranks), the same step with the gradient all-reduce suppressed on every micro-step
(`no_sync`), and their difference = exposed communication time of the bucketed NCCL all-reduce over NVLink."""
import argparse
import contextlib
import json
import os
import sys

os.environ.setdefault("EWVIT_ALLOW_RANDOM_BACKBONE", "1")

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
os.chdir(os.path.join(REPO, "efficient-wavelet-vit_b200"))

from ewvit.training import combined_loss, freeze_unused_, train_step  # noqa: E402
from network.model import DeepfakeDetector  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=8)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    results = {}
    for mode in ("dynamic", "sfe_only", "sfe_mwt"):
        torch.manual_seed(42)
        base = DeepfakeDetector(3, 128, batch_size=8).to(dev).train()
        n_train = freeze_unused_(base, mode)                 # parameters outside the mode's path get no gradient
        model = torch.nn.parallel.DistributedDataParallel(base, device_ids=[local]) if world > 1 else base
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        micro = [(torch.randn(args.videos, args.frames, 3, 224, 224, device=dev, generator=g),
                  torch.randint(0, 2, (args.videos,), device=dev, generator=g)) for _ in range(2)]

        def step(sync=True):
            if sync:
                return train_step(model, micro, opt, ablation=mode, batch_size=8, epoch=5, max_epochs=10)
            opt.zero_grad(set_to_none=True)                  # same arithmetic, gradient all-reduce suppressed everywhere
            for frames, labels in micro:
                with (model.no_sync() if world > 1 else contextlib.nullcontext()):
                    (combined_loss(model(frames, 8, mode), labels, 5, 10) / 2).backward()
            opt.step()
            return 0.0

        def timed(sync):
            step(sync)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                loss = step(sync)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()), loss

        ms_sync, loss = timed(True)
        ms_nosync, _ = timed(False)
        frames = 2 * args.videos * args.frames * world
        results[mode] = {"ms_per_step": ms_sync, "ms_per_step_no_allreduce": ms_nosync,
                         "exposed_allreduce_ms": ms_sync - ms_nosync, "exposed_allreduce_frac": (ms_sync - ms_nosync) / ms_sync,
                         "train_frames_per_s": frames / ms_sync * 1e3, "trainable_params": n_train, "loss": loss,
                         "grad_bytes_allreduced": 4 * n_train, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        del model, base, opt, micro
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"workload": f"configs[4]: training step, {args.videos} videos x {args.frames} frames per rank per micro-step, accum_steps 2, "
                                      "batch_size 8, BinaryFocalLoss + orthogonality, Adam; DDP (NCCL) one process per GPU",
                          "n_gpus": world, "arithmetic": "PyTorch composition (autograd/cuDNN/cuBLAS fp32) + native Haar fwd/adjoint: row f-2 partial",
                          "modes": results}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
