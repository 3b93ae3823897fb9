"""ORACLE (test infrastructure, never on the product path).

CPU/torch restatement of the image-feature branch the reference's checked-in `model_forward` actually
takes ("variant B"):

    image_embeddings = get_dinov2_visual_embs(images)           reference model/LISA.py:186-199
    image_embeddings = self.model.lisa_dino_conv(image_embeddings)   model/LISA.py:244-245, :92

`visual_model_dinov2` is `torch.hub.load('facebookresearch/dinov2', 'dinov2_vitl14')` (LISA.py:48):
a THIRD-PARTY dependency at an un-pinned `main`, NOT under /root/reference and not fetchable offline.
The restatement follows the published algorithm of `DinoVisionTransformer.forward_features`
(dinov2/models/vision_transformer.py, dinov2/layers/{block,attention,mlp,layer_scale,patch_embed}.py):

    x = patch_embed(img)  (conv 14x14 / stride 14, bias)         -> [B, g*g, D]
    x = cat(cls_token, x) + interpolate_pos_encoding(pos_embed)  (bicubic from the 37x37 training grid)
    for blk: x = x + ls1 * proj(softmax((q*hd^-.5) k^T) v) with q,k,v = qkv(norm1(x))
             x = x + ls2 * fc2(gelu_erf(fc1(norm2(x))))          (LayerNorm eps 1e-6)
    x_norm_patchtokens = norm(x)[:, 1:]                          (ViT-L/14 has no register tokens)

then the reference reshapes [1, g*g, 1024] -> [1, 1024, g, g] and applies the 1x1 conv 1024 -> 256.

State-dict names are the hub module's (`cls_token`, `pos_embed`, `patch_embed.proj.*`,
`blocks.{i}.{norm1,attn.qkv,attn.proj,ls1.gamma,norm2,mlp.fc1,mlp.fc2,ls2.gamma}`, `norm.*`) under the
reference prefix `model.visual_model_dinov2.`; the conv is `model.lisa_dino_conv.{weight,bias}`.

PARITY PINNING: the hub sources are absent, so the restatement is pinned against the INSTALLED
transformers `Dinov2Model` (eager; the same architecture re-implemented from that repo) by
oracle/make_golden.py -> tests/golden/dinov2_tiny.pt.  One detail stays "parity unpinned": hub `main`
of late 2023 resampled the position table with `scale_factor=(g+0.1)/37` (`interpolate_offset=0.1`),
transformers uses `size=(g, g)`; `interpolate_offset` selects either (0.1 = hub default, 0.0 = the
pinned transformers behaviour).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class Dinov2Config:
    img_size: int = 896          # 64x64 patches: `reshape(1, 1024, 64, 64)` at LISA.py:195
    patch_size: int = 14
    embed_dim: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_ratio: float = 4.0
    train_grid: int = 37         # 518 / 14: the grid the checkpoint's pos_embed is stored at
    ln_eps: float = 1e-6
    out_chans: int = 256         # lisa_dino_conv
    interpolate_offset: float = 0.1

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size


def interpolate_pos_embed(pos_embed: Tensor, grid: int, offset: float = 0.1) -> Tensor:
    """[1, 1+M*M, D] -> [1, 1+grid*grid, D]; bicubic, align_corners=False, no antialias
    (DinoVisionTransformer.interpolate_pos_encoding).  A weight-only transform, done once at load."""
    n = pos_embed.shape[1] - 1
    M = int(round(math.sqrt(n)))
    assert M * M == n
    if M == grid:
        return pos_embed
    D = pos_embed.shape[-1]
    cls_pos, patch_pos = pos_embed[:, :1], pos_embed[:, 1:]
    p = patch_pos.float().reshape(1, M, M, D).permute(0, 3, 1, 2)
    if offset:
        s = float(grid + offset) / M
        p = F.interpolate(p, scale_factor=(s, s), mode="bicubic", align_corners=False)
    else:
        p = F.interpolate(p, size=(grid, grid), mode="bicubic", align_corners=False)
    assert p.shape[-2:] == (grid, grid)
    p = p.permute(0, 2, 3, 1).reshape(1, grid * grid, D).to(pos_embed.dtype)
    return torch.cat([cls_pos, p], dim=1)


def forward_features(images: Tensor, sd: Dict[str, Tensor], cfg: Dinov2Config) -> Tensor:
    """[B,3,S,S] -> x_norm_patchtokens [B, g*g, D]."""
    B = images.shape[0]
    D, H = cfg.embed_dim, cfg.num_heads
    hd = D // H
    x = F.conv2d(images, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd["cls_token"].expand(B, -1, -1), x], dim=1)
    x = x + interpolate_pos_embed(sd["pos_embed"], cfg.grid, cfg.interpolate_offset)
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)
        qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        qkv = qkv.reshape(B, -1, 3, H, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
        att = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, -1, D)
        x = x + sd[p + "ls1.gamma"] * F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        x = x + sd[p + "ls2.gamma"] * F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    x = F.layer_norm(x, (D,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)
    return x[:, 1:]


def image_embeddings(images: Tensor, sd_dino: Dict[str, Tensor], conv_w: Tensor, conv_b: Tensor,
                     cfg: Dinov2Config) -> Tensor:
    """get_dinov2_visual_embs + lisa_dino_conv (LISA.py:186-199,244-245) -> [B, out_chans, g, g]."""
    g = cfg.grid
    tok = forward_features(images, sd_dino, cfg)
    nchw = tok.permute(0, 2, 1).reshape(images.shape[0], cfg.embed_dim, g, g)
    return F.conv2d(nchw, conv_w, conv_b)


def random_state_dict(cfg: Dinov2Config, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Hub-named random weights (LayerScale gammas O(1) rather than the 1e-5 init, so the branch matters)."""
    g = torch.Generator().manual_seed(seed)
    D, p = cfg.embed_dim, cfg.patch_size
    mlp = int(D * cfg.mlp_ratio)

    def rn(*s, std=0.02):
        return (torch.randn(*s, generator=g) * std).to(dtype)

    sd = {
        "cls_token": rn(1, 1, D, std=0.5),
        "pos_embed": rn(1, 1 + cfg.train_grid ** 2, D, std=0.3),
        "mask_token": rn(1, D),
        "patch_embed.proj.weight": rn(D, 3, p, p, std=(3 * p * p) ** -0.5),
        "patch_embed.proj.bias": rn(D, std=0.1),
        "norm.weight": 1 + rn(D, std=0.1), "norm.bias": rn(D, std=0.1),
    }
    for i in range(cfg.depth):
        b = f"blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[b + n + ".weight"] = 1 + rn(D, std=0.1)
            sd[b + n + ".bias"] = rn(D, std=0.1)
        sd[b + "attn.qkv.weight"] = rn(3 * D, D, std=D ** -0.5)
        sd[b + "attn.qkv.bias"] = rn(3 * D, std=0.1)
        sd[b + "attn.proj.weight"] = rn(D, D, std=D ** -0.5)
        sd[b + "attn.proj.bias"] = rn(D)
        sd[b + "ls1.gamma"] = 0.5 + rn(D, std=0.1)
        sd[b + "mlp.fc1.weight"] = rn(mlp, D, std=D ** -0.5)
        sd[b + "mlp.fc1.bias"] = rn(mlp)
        sd[b + "mlp.fc2.weight"] = rn(D, mlp, std=mlp ** -0.5)
        sd[b + "mlp.fc2.bias"] = rn(D)
        sd[b + "ls2.gamma"] = 0.5 + rn(D, std=0.1)
    return sd


def to_hf_names(sd: Dict[str, Tensor], cfg: Dinov2Config) -> Dict[str, Tensor]:
    """hub names -> transformers `Dinov2Model` names (for the pinning run in make_golden)."""
    D = cfg.embed_dim
    out = {
        "embeddings.cls_token": sd["cls_token"], "embeddings.position_embeddings": sd["pos_embed"],
        "embeddings.mask_token": sd["mask_token"],
        "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
        "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
        "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"],
    }
    for i in range(cfg.depth):
        b, h = f"blocks.{i}.", f"encoder.layer.{i}."
        for j, n in enumerate(("query", "key", "value")):
            out[h + f"attention.attention.{n}.weight"] = sd[b + "attn.qkv.weight"][j * D:(j + 1) * D]
            out[h + f"attention.attention.{n}.bias"] = sd[b + "attn.qkv.bias"][j * D:(j + 1) * D]
        out[h + "attention.output.dense.weight"] = sd[b + "attn.proj.weight"]
        out[h + "attention.output.dense.bias"] = sd[b + "attn.proj.bias"]
        out[h + "layer_scale1.lambda1"] = sd[b + "ls1.gamma"]
        out[h + "layer_scale2.lambda1"] = sd[b + "ls2.gamma"]
        for n in ("norm1", "norm2", "mlp.fc1", "mlp.fc2"):
            out[h + n + ".weight"] = sd[b + n + ".weight"]
            out[h + n + ".bias"] = sd[b + n + ".bias"]
    return out
