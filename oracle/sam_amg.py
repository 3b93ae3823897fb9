"""ORACLE (test infrastructure, never on the product path).

Torch restatement of SAM-Everything proposal generation as LLM-Seg uses it:

  * prompt encoder for point prompts            reference model/segment_anything/modeling/prompt_encoder.py:62-90,171-238
  * mask decoder (two-way transformer, upscaling, hyper-networks, IoU head)
                                                reference modeling/mask_decoder.py:71-164, modeling/transformer.py:63-242
  * `SamPredictor.predict_torch` + `Sam.postprocess_masks` for an already square 1024 x 1024 input
                                                reference predictor.py:166-241, modeling/sam.py:135-166
  * `SamAutomaticMaskGenerator._process_batch / _process_crop / generate` with one crop layer
                                                reference automatic_mask_generator.py:189-322, utils/amg.py:156-176,303-346
  * LLM-Seg's consumer side: keep the 50 largest masks (reference utils/sam_mask_reader.py:69-83) and resize them to
    256 x 256 soft masks with an antialiased bilinear filter (reference utils/dataset.py:620-622)

Weights use the reference's state-dict names below `model.visual_model.` (`prompt_encoder.*`, `mask_decoder.*`).
Pinned by tests/golden/sam_amg_*.pt: outputs of the reference's own `PromptEncoder`, `MaskDecoder` and
`SamAutomaticMaskGenerator` classes on seeded inputs (oracle/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

EMBED = 256
HEADS = 8
IMG = 1024
GRID = 64


# ---- prompt encoder ---------------------------------------------------------------------------
def pe_encoding(coords01: Tensor, gauss: Tensor) -> Tensor:
    """`PositionEmbeddingRandom._pe_encoding` (prompt_encoder.py:186-195): coords in [0,1]^2 -> [..., 256]."""
    c = 2 * coords01 - 1
    c = c.to(gauss.dtype) @ gauss
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd: Dict[str, Tensor], size: int = GRID) -> Tensor:
    """`get_dense_pe` (prompt_encoder.py:62-71,197-210): [1, 256, size, size]."""
    g = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    grid = torch.ones((size, size), device=g.device, dtype=g.dtype)
    y = (grid.cumsum(dim=0) - 0.5) / size
    x = (grid.cumsum(dim=1) - 0.5) / size
    return pe_encoding(torch.stack([x, y], dim=-1), g).permute(2, 0, 1).unsqueeze(0)


def embed_points(points_xy: Tensor, sd: Dict[str, Tensor], img: int = IMG) -> Tensor:
    """One foreground point per prompt + the padding point `boxes is None` adds (prompt_encoder.py:73-90,151-164):
    points_xy [P, 2] (x, y) in input-frame pixels -> sparse embeddings [P, 2, 256]."""
    g = sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    pts = (points_xy.to(torch.float32) + 0.5) / img
    e = pe_encoding(pts, g) + sd["prompt_encoder.point_embeddings.1.weight"]          # label 1
    pad = sd["prompt_encoder.not_a_point_embed.weight"].expand(points_xy.shape[0], -1)
    return torch.stack([e.to(pad.dtype), pad], dim=1)


# ---- two-way transformer ----------------------------------------------------------------------------
def attention(q: Tensor, k: Tensor, v: Tensor, sd: Dict[str, Tensor], prefix: str, heads: int = HEADS) -> Tensor:
    """`Attention.forward` (modeling/transformer.py:222-242) incl. the down-scaled internal dimension."""
    q = F.linear(q, sd[prefix + "q_proj.weight"], sd[prefix + "q_proj.bias"])
    k = F.linear(k, sd[prefix + "k_proj.weight"], sd[prefix + "k_proj.bias"])
    v = F.linear(v, sd[prefix + "v_proj.weight"], sd[prefix + "v_proj.bias"])

    def split(x):
        b, n, c = x.shape
        return x.reshape(b, n, heads, c // heads).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1]), dim=-1)
    o = (att @ v).transpose(1, 2)
    o = o.reshape(o.shape[0], o.shape[1], -1)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def _ln(x: Tensor, sd: Dict[str, Tensor], name: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def two_way_block(queries, keys, query_pe, key_pe, sd, prefix: str, skip_first_layer_pe: bool):
    """`TwoWayAttentionBlock.forward` (modeling/transformer.py:151-184)."""
    if skip_first_layer_pe:
        queries = attention(queries, queries, queries, sd, prefix + "self_attn.")
    else:
        q = queries + query_pe
        queries = queries + attention(q, q, queries, sd, prefix + "self_attn.")
    queries = _ln(queries, sd, prefix + "norm1")
    q, k = queries + query_pe, keys + key_pe
    queries = _ln(queries + attention(q, k, keys, sd, prefix + "cross_attn_token_to_image."), sd, prefix + "norm2")
    h = F.linear(F.relu(F.linear(queries, sd[prefix + "mlp.lin1.weight"], sd[prefix + "mlp.lin1.bias"])),
                 sd[prefix + "mlp.lin2.weight"], sd[prefix + "mlp.lin2.bias"])
    queries = _ln(queries + h, sd, prefix + "norm3")
    q, k = queries + query_pe, keys + key_pe
    keys = _ln(keys + attention(k, q, queries, sd, prefix + "cross_attn_image_to_token."), sd, prefix + "norm4")
    return queries, keys


def two_way_transformer(src: Tensor, pos: Tensor, tokens: Tensor, sd: Dict[str, Tensor], prefix: str):
    """`TwoWayTransformer.forward` (modeling/transformer.py:63-107): src/pos [B,256,h,w], tokens [B,N,256]."""
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pos.flatten(2).permute(0, 2, 1)
    queries = tokens
    for i in range(2):
        queries, keys = two_way_block(queries, keys, tokens, key_pe, sd, f"{prefix}layers.{i}.", i == 0)
    q, k = queries + tokens, keys + key_pe
    queries = _ln(queries + attention(q, k, keys, sd, prefix + "final_attn_token_to_image."), sd, prefix + "norm_final_attn")
    return queries, keys


def _mlp3(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    for i in range(3):
        x = F.linear(x, sd[f"{prefix}layers.{i}.weight"], sd[f"{prefix}layers.{i}.bias"])
        if i < 2:
            x = F.relu(x)
    return x


def layer_norm_2d(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def mask_decoder(image_embedding: Tensor, sparse: Tensor, sd: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
    """`MaskDecoder.predict_masks` (modeling/mask_decoder.py:116-164) with the dense embedding of `masks=None`
    (prompt_encoder.py:231-235): image_embedding [1,256,64,64], sparse [P,2,256] -> masks [P,4,256,256], iou [P,4]."""
    p = "mask_decoder."
    P = sparse.shape[0]
    out_tok = torch.cat([sd[p + "iou_token.weight"], sd[p + "mask_tokens.weight"]], dim=0)
    tokens = torch.cat([out_tok.unsqueeze(0).expand(P, -1, -1), sparse], dim=1)
    src = image_embedding.expand(P, -1, -1, -1) + sd["prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1)
    pos = dense_pe(sd, image_embedding.shape[-1]).to(src.dtype).expand(P, -1, -1, -1)
    b, c, h, w = src.shape
    hs, src2 = two_way_transformer(src, pos, tokens, sd, p + "transformer.")
    iou_tok, mask_toks = hs[:, 0, :], hs[:, 1:5, :]
    x = src2.transpose(1, 2).reshape(b, c, h, w)
    x = F.conv_transpose2d(x, sd[p + "output_upscaling.0.weight"], sd[p + "output_upscaling.0.bias"], stride=2)
    x = F.gelu(layer_norm_2d(x, sd[p + "output_upscaling.1.weight"], sd[p + "output_upscaling.1.bias"]))
    x = F.gelu(F.conv_transpose2d(x, sd[p + "output_upscaling.3.weight"], sd[p + "output_upscaling.3.bias"], stride=2))
    hyper = torch.stack([_mlp3(mask_toks[:, i, :], sd, f"{p}output_hypernetworks_mlps.{i}.") for i in range(4)], dim=1)
    b, c, h, w = x.shape
    masks = (hyper @ x.view(b, c, h * w)).view(b, 4, h, w)
    return masks, _mlp3(iou_tok, sd, p + "iou_prediction_head.")


def predict_points(image_embedding: Tensor, points_xy: Tensor, sd: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
    """`predict_torch(point_coords[:,None], labels=1, multimask_output=True)` (predictor.py:216-232):
    low-res mask logits [P,3,256,256] and IoU predictions [P,3]."""
    masks, iou = mask_decoder(image_embedding, embed_points(points_xy, sd), sd)
    return masks[:, 1:], iou[:, 1:]


def upsample_logits(low_res: Tensor, size: int = IMG) -> Tensor:
    """`postprocess_masks` for a square input of the encoder's own size (modeling/sam.py:155-166)."""
    return F.interpolate(low_res.float(), (size, size), mode="bilinear", align_corners=False)


# ---- automatic mask generator ----------------------------------------------------------------------
def point_grid(n_per_side: int, size: int = IMG) -> Tensor:
    """`build_point_grid` scaled to the image (utils/amg.py:179-186, automatic_mask_generator.py:240-241): [n^2, 2] (x, y)."""
    off = 1 / (2 * n_per_side)
    one = torch.linspace(off, 1 - off, n_per_side, dtype=torch.float64)
    x = one[None, :].expand(n_per_side, -1)
    y = one[:, None].expand(-1, n_per_side)
    return (torch.stack([x, y], dim=-1).reshape(-1, 2) * size).to(torch.float32)


def stability_score(logits: Tensor, thr: float = 0.0, off: float = 1.0) -> Tensor:
    """utils/amg.py:156-176."""
    inter = (logits > thr + off).flatten(-2).sum(-1).to(torch.float32)
    union = (logits > thr - off).flatten(-2).sum(-1).to(torch.float32)
    return inter / union


def mask_to_box(masks: Tensor) -> Tensor:
    """`batched_mask_to_box` (utils/amg.py:303-346): XYXY (inclusive pixel indices), [0,0,0,0] for an empty mask."""
    h, w = masks.shape[-2:]
    rows = masks.any(dim=-1)
    cols = masks.any(dim=-2)
    ar_h = torch.arange(h, device=masks.device)
    ar_w = torch.arange(w, device=masks.device)
    bottom = (rows * ar_h).max(-1).values
    top = (rows * ar_h + h * (~rows)).min(-1).values
    right = (cols * ar_w).max(-1).values
    left = (cols * ar_w + w * (~cols)).min(-1).values
    empty = (right < left) | (bottom < top)
    out = torch.stack([left, top, right, bottom], dim=-1)
    return out * (~empty).unsqueeze(-1)


def box_iou(a: Tensor, b: Tensor) -> Tensor:
    """torchvision.ops.box_iou semantics (area = (x2-x1)*(y2-y1) on the raw coordinates)."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def nms(boxes: Tensor, scores: Tensor, thr: float) -> Tensor:
    """Greedy box NMS, scores descending (stable), suppress IoU > thr: what `batched_nms` with a single category
    does (automatic_mask_generator.py:256-262).  Returns kept indices in score order."""
    order = torch.argsort(scores, descending=True, stable=True)
    iou = box_iou(boxes[order].float(), boxes[order].float())
    keep, dead = [], torch.zeros(len(order), dtype=torch.bool)
    for i in range(len(order)):
        if dead[i]:
            continue
        keep.append(int(order[i]))
        dead |= (iou[i] > thr).cpu()
    return torch.tensor(keep, dtype=torch.long)


def generate(image_embedding: Tensor, sd: Dict[str, Tensor], *, points_per_side: int = 32, points_per_batch: int = 64,
             pred_iou_thresh: float = 0.88, stability_score_thresh: float = 0.95, stability_score_offset: float = 1.0,
             box_nms_thresh: float = 0.7, size: int = IMG, decoded: Optional[Tuple[Tensor, Tensor]] = None) -> dict:
    """`SamAutomaticMaskGenerator.generate` (automatic_mask_generator.py:141-322) with crop_n_layers = 0 and
    min_mask_region_area = 0 (the class defaults) on the features of one square image:
    -> {"masks" bool [M,size,size], "boxes" XYXY, "iou_preds", "stability", "points", "areas", "candidates"} after the
    box NMS ("candidates": index 3 * prompt + mask of every record).  decoded = (low-res logits [P,3,256,256], IoU
    predictions [P,3]) replaces the mask decoder (tests: post-processing of another implementation's logits)."""
    pts = point_grid(points_per_side, size).to(image_embedding.device if decoded is None else decoded[0].device)
    acc: Dict[str, List[Tensor]] = {k: [] for k in ("masks", "iou_preds", "stability", "points", "candidates")}
    for i in range(0, pts.shape[0], points_per_batch):
        p = pts[i:i + points_per_batch]
        if decoded is None:
            low, iou = predict_points(image_embedding, p, sd)
        else:
            low, iou = decoded[0][i:i + points_per_batch], decoded[1][i:i + points_per_batch]
        logits = upsample_logits(low, size).flatten(0, 1)                  # [3P, size, size]
        iou = iou.flatten(0, 1).float()
        rep = p.repeat_interleave(3, dim=0)       # MaskData repeats the points per mask (np.repeat, :283)
        cid = torch.arange(3 * i, 3 * i + logits.shape[0], device=logits.device)
        keep = iou > pred_iou_thresh if pred_iou_thresh > 0.0 else torch.ones_like(iou, dtype=torch.bool)
        logits, iou, rep, cid = logits[keep], iou[keep], rep[keep], cid[keep]
        stab = stability_score(logits, 0.0, stability_score_offset)
        if stability_score_thresh > 0.0:
            keep = stab >= stability_score_thresh
            logits, iou, rep, stab, cid = logits[keep], iou[keep], rep[keep], stab[keep], cid[keep]
        acc["masks"].append(logits > 0.0)
        acc["iou_preds"].append(iou)
        acc["stability"].append(stab)
        acc["points"].append(rep)
        acc["candidates"].append(cid)
    data = {k: torch.cat(v, dim=0) for k, v in acc.items()}
    data["boxes"] = mask_to_box(data["masks"])
    keep = nms(data["boxes"], data["iou_preds"], box_nms_thresh).to(data["boxes"].device)
    data = {k: v[keep] for k, v in data.items()}
    data["areas"] = data["masks"].flatten(1).sum(-1)
    return data


def llmseg_proposals(data: dict, top_k: int = 50, out: int = 256) -> Tuple[Tensor, Tensor]:
    """What LLM-Seg feeds its selector from the generator's records: the `top_k` largest masks (stable sort by area,
    descending — utils/sam_mask_reader.py:75-83) resized with an antialiased bilinear filter to `out` x `out` soft
    masks (utils/dataset.py:620-622).  -> (soft masks fp32 [K,out,out], indices into `data`)."""
    order = torch.argsort(data["areas"], descending=True, stable=True)[:top_k]
    m = data["masks"][order].float()
    soft = F.interpolate(m.unsqueeze(0), size=(out, out), mode="bilinear", align_corners=False, antialias=True).squeeze(0)
    return soft, order


def aa_downsample_weights(in_size: int = IMG, out_size: int = 256):
    """The separable filter `F.interpolate(mode='bilinear', antialias=True)` applies when shrinking by s = in/out:
    output i averages inputs [lo_i, lo_i + n_i) with triangle weights of half-width s centred at (i + 0.5) * s
    (ATen `_compute_indices_weights_aa`).  -> (lo int64 [out], weights fp64 [out, max_n])."""
    s = in_size / out_size
    support = s
    max_n = int(math.ceil(support)) * 2 + 1
    lo = torch.zeros(out_size, dtype=torch.int64)
    w = torch.zeros(out_size, max_n, dtype=torch.float64)
    for i in range(out_size):
        center = s * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        ws = []
        for j in range(xmax - xmin):
            x = (j + xmin - center + 0.5) / s
            ws.append(max(0.0, 1.0 - abs(x)))
        tot = sum(ws)
        lo[i] = xmin
        for j, v in enumerate(ws):
            w[i, j] = v / tot
    return lo, w


# ---- synthetic weights ---------------------------------------------------------------------------------
def random_state_dict(seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Prompt-encoder + mask-decoder weights with the reference's names and shapes (under `model.visual_model.`)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def rn(*shape, std=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).to(dtype)

    def lin(name, out_f, in_f, std=None):
        sd[name + ".weight"] = rn(out_f, in_f, std=std if std is not None else in_f ** -0.5)
        sd[name + ".bias"] = rn(out_f, std=0.05)

    def ln(name, dim):
        sd[name + ".weight"] = rn(dim, std=0.1, mean=1.0)
        sd[name + ".bias"] = rn(dim, std=0.1)

    sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"] = rn(2, EMBED // 2, std=1.0)
    for i in range(4):
        sd[f"prompt_encoder.point_embeddings.{i}.weight"] = rn(1, EMBED, std=0.5)
    sd["prompt_encoder.not_a_point_embed.weight"] = rn(1, EMBED, std=0.5)
    sd["prompt_encoder.no_mask_embed.weight"] = rn(1, EMBED, std=0.5)
    t = "mask_decoder.transformer."
    for i in range(2):
        p = f"{t}layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj"):
            lin(p + "self_attn." + proj, EMBED, EMBED)
        lin(p + "self_attn.out_proj", EMBED, EMBED)
        for att in ("cross_attn_token_to_image", "cross_attn_image_to_token"):
            for proj in ("q_proj", "k_proj", "v_proj"):
                lin(p + att + "." + proj, EMBED // 2, EMBED)
            lin(p + att + ".out_proj", EMBED, EMBED // 2)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            ln(p + n, EMBED)
        lin(p + "mlp.lin1", 2048, EMBED)
        lin(p + "mlp.lin2", EMBED, 2048)
    for proj in ("q_proj", "k_proj", "v_proj"):
        lin(t + "final_attn_token_to_image." + proj, EMBED // 2, EMBED)
    lin(t + "final_attn_token_to_image.out_proj", EMBED, EMBED // 2)
    ln(t + "norm_final_attn", EMBED)
    d = "mask_decoder."
    sd[d + "iou_token.weight"] = rn(1, EMBED, std=0.5)
    sd[d + "mask_tokens.weight"] = rn(4, EMBED, std=0.5)
    sd[d + "output_upscaling.0.weight"] = rn(EMBED, 64, 2, 2, std=EMBED ** -0.5)
    sd[d + "output_upscaling.0.bias"] = rn(64, std=0.05)
    ln(d + "output_upscaling.1", 64)
    sd[d + "output_upscaling.3.weight"] = rn(64, 32, 2, 2, std=64 ** -0.5)
    sd[d + "output_upscaling.3.bias"] = rn(32, std=0.05)
    for i in range(4):
        for j, (o, k) in enumerate(((EMBED, EMBED), (EMBED, EMBED), (32, EMBED))):
            # last layer x8: mask logits of +-20 like a trained decoder's, so that the stability score (IoU of the
            # masks at thresholds +1 / -1) spreads over (0, 1) instead of sitting near 0
            lin(f"{d}output_hypernetworks_mlps.{i}.layers.{j}", o, k, std=(8.0 if j == 2 else 1.0) * k ** -0.5)
    for j, (o, k) in enumerate(((EMBED, EMBED), (EMBED, EMBED), (4, EMBED))):
        lin(f"{d}iou_prediction_head.layers.{j}", o, k)
    return sd
