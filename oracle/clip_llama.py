"""ORACLE (test infrastructure, never on the product path).

CPU/torch restatement of the two third-party encoders the reference calls through
`transformers==4.29.0` (reference requirements.txt:276; NOT under /root/reference):

  * CLIP ViT-L/14 vision tower, `hidden_states[-2]` without CLS
        call sites: reference model/llava/model/multimodal_encoder/clip_encoder.py:31-60
        algorithm : transformers 4.29 `CLIPVisionTransformer` (pre_layrnorm, 24 pre-LN layers,
                    q scaled by hd^-0.5 before QKᵀ, quick_gelu MLP, LN eps 1e-5)
  * LLaMA-7B decoder stack on input embeddings, last hidden state
        call sites: reference model/llava/model/language_model/llava_llama.py:93-102,124-127
        algorithm : transformers 4.29 `LlamaModel` (RMSNorm with fp32 variance, rotate-half RoPE
                    θ=1e4, causal ∧ key-padding additive mask, fp32 softmax, SwiGLU MLP)

PARITY PINNING: the pinned 4.29.0 sources are not available offline.  The restatement is pinned
against the INSTALLED transformers (5.5.0) eager modules by oracle/make_golden.py (same recipe at
fp32; CLIP differs only in where the 1/sqrt(d) scale is applied, which is exact in fp32) and by the
golden vectors tests/golden/clip_tiny.pt / llama_tiny.pt generated there.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# CLIP ViT
# ----------------------------------------------------------------------------------------------
@dataclass
class ClipConfig:
    image_size: int = 224
    patch_size: int = 14
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    eps: float = 1e-5
    select_layer: int = -2  # LLaVA `mm_vision_select_layer` (reference LISA.py:134-135)

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


def clip_patch_features(images: Tensor, sd: Dict[str, Tensor], cfg: ClipConfig, prefix: str = "vision_model.") -> Tensor:
    """[N,3,224,224] -> [N,256,hidden]: hidden_states[select_layer][:, 1:] (clip_encoder.py:31-39,53-57)."""
    N = images.shape[0]
    x = F.conv2d(images, sd[prefix + "embeddings.patch_embedding.weight"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[prefix + "embeddings.class_embedding"].expand(N, 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[prefix + "embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (cfg.hidden,), sd[prefix + "pre_layrnorm.weight"], sd[prefix + "pre_layrnorm.bias"], cfg.eps)
    # hidden_states = (x0, x1, ..., xL); select_layer -2 == output of layer L-1
    n_run = cfg.layers + 1 + cfg.select_layer if cfg.select_layer < 0 else cfg.select_layer
    hd = cfg.hidden // cfg.heads
    for i in range(n_run):
        p = f"{prefix}encoder.layers.{i}."
        h = F.layer_norm(x, (cfg.hidden,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], cfg.eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]) * hd ** -0.5
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        sp = lambda t: t.reshape(N, -1, cfg.heads, hd).transpose(1, 2)
        att = torch.softmax(sp(q) @ sp(k).transpose(-1, -2), dim=-1)
        o = (att @ sp(v)).transpose(1, 2).reshape(N, -1, cfg.hidden)
        x = x + F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (cfg.hidden,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], cfg.eps)
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu
        x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x[:, 1:]


def clip_random_state_dict(cfg: ClipConfig, seed: int = 0, dtype=torch.float32, prefix: str = "vision_model.") -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    D = cfg.hidden

    def rn(*s, std=0.02):
        return (torch.randn(*s, generator=g) * std).to(dtype)

    sd = {
        prefix + "embeddings.class_embedding": rn(D, std=0.5),
        prefix + "embeddings.patch_embedding.weight": rn(D, 3, cfg.patch_size, cfg.patch_size, std=(3 * cfg.patch_size ** 2) ** -0.5),
        prefix + "embeddings.position_embedding.weight": rn(cfg.tokens, D, std=0.3),
        prefix + "pre_layrnorm.weight": 1 + rn(D, std=0.1), prefix + "pre_layrnorm.bias": rn(D, std=0.1),
    }
    for i in range(cfg.layers):
        p = f"{prefix}encoder.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            sd[p + n + ".weight"] = 1 + rn(D, std=0.1)
            sd[p + n + ".bias"] = rn(D, std=0.1)
        for n in ("q_proj", "k_proj", "v_proj"):
            sd[p + f"self_attn.{n}.weight"] = rn(D, D, std=D ** -0.5)
            sd[p + f"self_attn.{n}.bias"] = rn(D, std=0.1)
        sd[p + "self_attn.out_proj.weight"] = rn(D, D, std=0.5 * D ** -0.5)
        sd[p + "self_attn.out_proj.bias"] = rn(D)
        sd[p + "mlp.fc1.weight"] = rn(cfg.mlp, D, std=D ** -0.5)
        sd[p + "mlp.fc1.bias"] = rn(cfg.mlp)
        sd[p + "mlp.fc2.weight"] = rn(D, cfg.mlp, std=0.5 * cfg.mlp ** -0.5)
        sd[p + "mlp.fc2.bias"] = rn(D)
    return sd


# ----------------------------------------------------------------------------------------------
# LLaMA
# ----------------------------------------------------------------------------------------------
@dataclass
class LlamaConfig:
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    mlp: int = 11008
    vocab: int = 32003
    eps: float = 1e-6
    rope_theta: float = 10000.0

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


def rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """transformers 4.29 LlamaRMSNorm: fp32 variance, cast to weight dtype, then scale."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    xn = x.float() * torch.rsqrt(var + eps)
    return w * xn.to(w.dtype)


def rope_tables(seq: int, head_dim: int, theta: float, dtype, device=None):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=device) / head_dim))
    fr = torch.outer(torch.arange(seq, dtype=torch.float32, device=device), inv)
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rot_half(x: Tensor) -> Tensor:
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def llama_last_hidden(embeds: Tensor, attention_mask: Optional[Tensor], sd: Dict[str, Tensor],
                      cfg: LlamaConfig, prefix: str = "") -> Tensor:
    """[N,T,hidden] input embeddings -> final-norm hidden states [N,T,hidden].

    position_ids = arange(T) for every row; additive mask = causal + key padding, as in 4.29
    `_prepare_decoder_attention_mask`."""
    N, T, D = embeds.shape
    H, hd = cfg.heads, cfg.head_dim
    dt = embeds.dtype
    cos, sin = rope_tables(T, hd, cfg.rope_theta, dt, embeds.device)
    neg = torch.finfo(dt).min
    mask = torch.full((T, T), neg, dtype=dt, device=embeds.device).triu(1)[None, None].expand(N, 1, T, T).clone()
    if attention_mask is not None:
        pad = (~attention_mask.bool())[:, None, None, :]
        mask = mask.masked_fill(pad, neg)
    x = embeds
    for i in range(cfg.layers):
        p = f"{prefix}layers.{i}."
        h = rms_norm(x, sd[p + "input_layernorm.weight"], cfg.eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).reshape(N, T, H, hd).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).reshape(N, T, H, hd).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).reshape(N, T, H, hd).transpose(1, 2)
        q = q * cos + _rot_half(q) * sin
        k = k * cos + _rot_half(k) * sin
        att = q @ k.transpose(-1, -2) / math.sqrt(hd) + mask
        att = torch.max(att, torch.tensor(neg, dtype=dt, device=att.device))
        att = torch.softmax(att, dim=-1, dtype=torch.float32).to(dt)
        o = (att @ v).transpose(1, 2).reshape(N, T, D)
        x = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], cfg.eps)
        h = F.silu(F.linear(h, sd[p + "mlp.gate_proj.weight"])) * F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(h, sd[p + "mlp.down_proj.weight"])
    return rms_norm(x, sd[prefix + "norm.weight"], cfg.eps)


def llama_random_state_dict(cfg: LlamaConfig, seed: int = 0, dtype=torch.float32, prefix: str = "") -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    D = cfg.hidden

    def rn(*s, std=0.02):
        return (torch.randn(*s, generator=g) * std).to(dtype)

    sd = {prefix + "embed_tokens.weight": rn(cfg.vocab, D, std=1.0), prefix + "norm.weight": 1 + rn(D, std=0.1)}
    for i in range(cfg.layers):
        p = f"{prefix}layers.{i}."
        sd[p + "input_layernorm.weight"] = 1 + rn(D, std=0.1)
        sd[p + "post_attention_layernorm.weight"] = 1 + rn(D, std=0.1)
        for n in ("q_proj", "k_proj", "v_proj"):
            sd[p + f"self_attn.{n}.weight"] = rn(D, D, std=D ** -0.5)
        sd[p + "self_attn.o_proj.weight"] = rn(D, D, std=0.5 * D ** -0.5)
        sd[p + "mlp.gate_proj.weight"] = rn(cfg.mlp, D, std=D ** -0.5)
        sd[p + "mlp.up_proj.weight"] = rn(cfg.mlp, D, std=D ** -0.5)
        sd[p + "mlp.down_proj.weight"] = rn(D, cfg.mlp, std=0.5 * cfg.mlp ** -0.5)
    return sd
