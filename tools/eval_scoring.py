"""BASELINE.json configs[3]: eval-style video scoring -- 64 synthetic videos x 300 frames, batches of 8 videos,
batch_size=8 (reference eval.py:135-194 with the DataLoader replaced by on-device synthetic videos), videos sharded over the
ranks of a torchrun launch with ONE NCCL all_gather of the per-video logits at the end.

    python tools/eval_scoring.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/eval_scoring.py

Prints one JSON line on rank 0: frames/s over the whole job (device time, max over ranks) and the gathered decisions."""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
os.chdir(os.path.join(REPO, "efficient-wavelet-vit_b200"))

from ewvit.distributed import score_videos_sharded, shard_range  # noqa: E402
from network.model import DeepfakeDetector  # noqa: E402

VIDEOS, FRAMES, PER_CALL, BATCH_SIZE = 64, 300, 8, 8


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    model = DeepfakeDetector(3, 128, batch_size=BATCH_SIZE).to(dev).eval()

    def make_videos(ids):            # seed 42 + video id, generated on the device (38.5 GB would not fit pinned host memory)
        out = torch.empty(len(ids), FRAMES, 3, 224, 224, device=dev)
        for k, vid in enumerate(ids):
            g = torch.Generator(device=dev).manual_seed(42 + vid)
            out[k] = torch.randn(FRAMES, 3, 224, 224, device=dev, generator=g)
        return out

    def score(x):
        return model(x, BATCH_SIZE, "dynamic")["logits"]

    with torch.no_grad():
        score(make_videos([0])[:, :64])                      # warm-up (kernel attributes, workspaces)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        logits = score_videos_sharded(score, make_videos, VIDEOS, PER_CALL, device=dev)
        e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    lo, hi = shard_range(VIDEOS, rank, world)
    if rank == 0:
        probs = torch.sigmoid(logits.float().cpu())
        print(json.dumps({"workload": "configs[3]: 64 videos x 300 frames, 8 videos per call, batch_size=8 (37 chunks of 64 + 1 of 32 frames per call)",
                          "n_gpus": world, "frames": VIDEOS * FRAMES, "ms": float(ms.item()),
                          "frames_per_s": VIDEOS * FRAMES / (float(ms.item()) / 1e3),
                          "note": "includes on-device synthetic frame generation (torch.randn) for every call",
                          "videos_on_rank0": hi - lo, "fake_decisions": int((probs >= 0.5).sum()), "logits_finite": bool(torch.isfinite(logits).all())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
