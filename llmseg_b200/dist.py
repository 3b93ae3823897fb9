"""Data-parallel sharding of the inference forward: one process per GPU, full weight replica per
rank (15.4 GB bf16 ≪ 180 GB), images split in contiguous blocks, and ONE collective per forward —
an all-gather of the per-image mask logits (the reference only all-reduces 3-float
meters, utils/utils.py:76-97).  The payload is `[B_local, 2*K_max + 2]` fp32 per rank
(similarity | iou | best index | K), ≤ 16 KiB: pure NVLink latency.

The collective goes through torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns (rank, world_size, local_rank); a single process without the env runs un-distributed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns [lo, hi) with sizes differing by at most one."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_logits(sim: Tensor, iou: Tensor, best: Tensor, ks: List[int], k_max: int, b_max: int) -> Tensor:
    """[b_max, 2*k_max+2] fp32: similarity (pad -inf) | iou (pad 0) | best | K; rows beyond the local
    batch are marked K = -1 so ragged shards survive the fixed-size all-gather."""
    dev = sim.device
    out = torch.zeros((b_max, 2 * k_max + 2), dtype=torch.float32, device=dev)
    out[:, :k_max] = float("-inf")
    out[:, 2 * k_max + 1] = -1.0
    b, k = sim.shape
    out[:b, :k] = sim
    out[:b, k_max:k_max + k] = iou
    out[:b, 2 * k_max] = best.to(torch.float32)
    out[:b, 2 * k_max + 1] = torch.tensor(ks, dtype=torch.float32, device=dev)
    return out


def unpack_logits(gathered: Tensor, k_max: int):
    """Inverse of pack_logits over the concatenated shards -> lists (sim [1,K], iou [1,K]), best list."""
    sims, ious, best = [], [], []
    for row in gathered:
        k = int(row[2 * k_max + 1].item())
        if k < 0:
            continue
        sims.append(row[:k].unsqueeze(0))
        ious.append(row[k_max:k_max + k].unsqueeze(0))
        best.append(int(row[2 * k_max].item()))
    return sims, ious, best


def all_gather_logits(local: Tensor) -> Tensor:
    """The single collective of the forward: concatenate every rank's packed logits (rank order)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class DataParallelLisa:
    """Shards a global batch over the ranks, runs the local forward and all-gathers the logits."""

    def __init__(self, model, k_max: int = 128):
        self.model = model
        self.k_max = k_max
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def forward(self, global_inputs: dict) -> dict:
        B = global_inputs["images"].shape[0]
        lo, hi = shard_range(B, self.rank, self.world)
        b_max = -(-B // self.world)
        loc = dict(global_inputs)
        for key in ("images", "images_clip", "input_ids", "labels", "attention_masks"):
            if loc.get(key) is not None:
                loc[key] = loc[key][lo:hi]
        for key in ("sam_segs_list", "masks_list", "label_list", "resize_list"):
            if loc.get(key) is not None:
                loc[key] = loc[key][lo:hi]
        loc["offset"] = torch.arange(hi - lo + 1)
        out = self.model.model_forward(**loc)
        ks = [int(s.shape[1]) for s in out["pred_similarity"]]
        packed = pack_logits(out["similarity_padded"], out["iou_padded"], out["best_index"], ks, self.k_max, b_max)
        sims, ious, best = unpack_logits(all_gather_logits(packed), self.k_max)
        return {"pred_similarity": sims, "pred_iou": ious, "best_index": best, "gt_masks": global_inputs.get("masks_list")}
