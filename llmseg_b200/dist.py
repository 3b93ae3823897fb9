"""Data-parallel sharding of the inference forward: one process per GPU, full weight replica per
rank (15.4 GB bf16 ≪ 180 GB), images split in contiguous blocks, and ONE collective per forward —
an all-gather of the per-image mask logits (the reference only all-reduces 3-float
meters, utils/utils.py:76-97).  The payload is `[B_local, 2*K_max + 3]` fp32 per rank
(similarity | iou | best index | K | C), ≤ 16 KiB: pure NVLink latency.

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


_PACK_TEMPLATES: dict = {}


def packed_width(k_max: int, c_cap: int = 1) -> int:
    return (c_cap + 1) * k_max + 3


def pack_logits(sim: Tensor, iou: Tensor, best: Tensor, ks: List[int], k_max: int, b_max: int,
                convs: Optional[List[int]] = None, c_cap: int = 1, device=None) -> Tensor:
    """One row per local image, [b_max, (c_cap+1)*k_max + 3] fp32:
        similarity of conversation 0..c_cap-1 (k_max each, pad -inf) | iou (pad 0) | best | K | C
    `sim` holds one row per CONVERSATION in image order (`convs[i]` rows for image i; default one each), `iou` and
    `best` one row per image.  Rows beyond the local batch are marked K = -1, so ragged — and empty — shards survive
    the fixed-size all-gather (pass sim=None with ks=[] for a rank that owns no image)."""
    dev = device if sim is None else sim.device
    w = packed_width(k_max, c_cap)
    b = len(ks)
    convs = [1] * b if convs is None else [int(c) for c in convs]
    # the padding / meta columns depend only on the shapes: built once per shape and cloned, so that a step issues no
    # host-to-device copy (a pageable H2D copy blocks the CPU until the stream has drained: one forward of run-ahead lost)
    key = (tuple(ks), tuple(convs), k_max, b_max, c_cap, str(dev))
    tmpl = _PACK_TEMPLATES.get(key)
    if tmpl is None:
        if b > b_max or (b and max(ks) > k_max):
            raise ValueError(f"pack_logits: {b} images / K={max(ks)} exceed the packed shape (b_max={b_max}, k_max={k_max})")
        if b and max(convs) > c_cap:
            raise ValueError(f"pack_logits: conversations per image {convs} exceed c_cap={c_cap}")
        t = torch.zeros((b_max, w), dtype=torch.float32)
        t[:, :c_cap * k_max] = float("-inf")
        t[:, w - 2] = -1.0
        if b:
            t[:b, w - 2] = torch.tensor(ks, dtype=torch.float32)
            t[:b, w - 1] = torch.tensor(convs, dtype=torch.float32)
        if len(_PACK_TEMPLATES) > 64:
            _PACK_TEMPLATES.clear()
        tmpl = _PACK_TEMPLATES[key] = t.to(dev)
    out = tmpl.clone()
    if b == 0:
        return out
    if sum(convs) != sim.shape[0]:
        raise ValueError(f"pack_logits: conversations per image {convs} vs {sim.shape[0]} similarity rows")
    k = min(sim.shape[1], k_max)
    if convs == [1] * b:
        out[:b, :k] = sim[:, :k]
    else:
        r = 0
        for i, c in enumerate(convs):
            for j in range(c):
                out[i, j * k_max:j * k_max + k] = sim[r, :k]
                r += 1
    out[:b, c_cap * k_max:c_cap * k_max + k] = iou[:, :k]
    out[:b, w - 3] = best.to(torch.float32)
    return out


def unpack_logits(gathered: Tensor, k_max: int, c_cap: int = 1):
    """Inverse of pack_logits over the concatenated shards -> lists (sim [C,K], iou [1,K]), best list."""
    sims, ious, best = [], [], []
    w = packed_width(k_max, c_cap)
    meta = gathered[:, w - 3:].cpu()
    for i, row in enumerate(gathered):
        k, c = int(meta[i, 1]), int(meta[i, 2])
        if k < 0:
            continue
        sims.append(torch.stack([row[j * k_max:j * k_max + k] for j in range(c)], dim=0))
        ious.append(row[c_cap * k_max:c_cap * k_max + k].unsqueeze(0))
        best.append(int(meta[i, 0]))
    return sims, ious, best


def all_gather_logits(local: Tensor, async_op: bool = False):
    """The single collective of the forward: concatenate every rank's packed logits (rank order).
    async_op: returns (out, work) with the collective in flight on the communicator's stream — call `work.wait()`
    (None when un-distributed) before reading `out`; the caller's stream is not blocked in between, so a rank can
    start its next forward while slower ranks are still finishing this one."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return (local, None) if async_op else local
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(out, local.contiguous(), async_op=async_op)
    return (out, work) if async_op else out


class DataParallelLisa:
    """Shards a global batch over the ranks by IMAGE (an image's conversations stay together, located through
    `offset`), runs the local forward and all-gathers the logits.  A rank whose shard is empty (global batch
    smaller than the world size: the ragged last batch of a validation epoch) skips the forward and contributes
    an all-padding block, so the collective still matches up on every rank."""

    def __init__(self, model, k_max: int = 128, device=None):
        self.model = model
        self.k_max = k_max
        self.device = device if device is not None else getattr(model, "device", None)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def shard(self, global_inputs: dict):
        """-> (local input dict or None for an empty shard, b_max, c_cap)."""
        B = global_inputs["images"].shape[0]
        offset = global_inputs.get("offset")
        off = list(range(B + 1)) if offset is None else [int(v) for v in offset.tolist()]
        n_conv = global_inputs["input_ids"].shape[0]
        if len(off) != B + 1 or off[0] != 0 or off[-1] != n_conv:
            raise ValueError(f"offset {off} does not partition {n_conv} conversations over {B} images")
        lo, hi = shard_range(B, self.rank, self.world)
        b_max = -(-B // self.world)
        c_cap = max([off[i + 1] - off[i] for i in range(B)] + [1])
        if hi == lo:
            return None, b_max, c_cap
        loc = dict(global_inputs)
        for key in ("images", "images_clip"):
            loc[key] = loc[key][lo:hi]
        for key in ("input_ids", "labels", "attention_masks"):
            if loc.get(key) is not None:
                loc[key] = loc[key][off[lo]:off[hi]]
        for key in ("sam_segs_list", "masks_list", "label_list", "resize_list", "sam_ious_list", "sam_iops_list"):
            if loc.get(key) is not None:
                loc[key] = loc[key][lo:hi]
        loc["offset"] = torch.tensor([o - off[lo] for o in off[lo:hi + 1]])
        return loc, b_max, c_cap

    def forward(self, global_inputs: dict) -> dict:
        loc, b_max, c_cap = self.shard(global_inputs)
        if loc is None:
            packed = pack_logits(None, None, None, [], self.k_max, b_max, c_cap=c_cap, device=self.device)
        else:
            out = self.model.model_forward(**loc)
            ks = [int(s.shape[1]) for s in out["pred_similarity"]]
            convs = [int(s.shape[0]) for s in out["pred_similarity"]]
            packed = pack_logits(out.get("similarity_all", out["similarity_padded"]), out["iou_padded"],
                                 out["best_index"], ks, self.k_max, b_max, convs=convs, c_cap=c_cap)
        sims, ious, best = unpack_logits(all_gather_logits(packed), self.k_max, c_cap)
        return {"pred_similarity": sims, "pred_iou": ious, "best_index": best, "gt_masks": global_inputs.get("masks_list")}
