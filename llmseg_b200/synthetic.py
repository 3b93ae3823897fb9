"""Seeded synthetic weights and inputs for the LLM-Seg forward: there are no
checkpoints or datasets offline, so benchmarks and parity tests use random-init weights with the
reference's parameter names/shapes and ReasonSeg-shaped inputs.

Weights are generated directly on the target device in bf16 (a 7B-parameter state dict takes
seconds on a GPU, minutes on the CPU).  This is set-up code, not part of the measured path.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .lisa import IMAGE_TOKEN_INDEX, ClipCfg, DinoCfg, LisaCfg, LlamaCfg, SamCfg

Tensor = torch.Tensor
BF16 = torch.bfloat16


class _Gen:
    def __init__(self, seed: int, device):
        self.g = torch.Generator(device=device).manual_seed(seed)
        self.device = device

    def rn(self, *shape, std=0.02, mean=0.0) -> Tensor:
        t = torch.randn(*shape, generator=self.g, device=self.device, dtype=torch.float32)
        return (t * std + mean).to(BF16)


def sam_state_dict(cfg: SamCfg, seed: int, device, prefix: str = "model.visual_model.image_encoder.") -> Dict[str, Tensor]:
    G = _Gen(seed, device)
    D, hd = cfg.embed_dim, cfg.embed_dim // cfg.num_heads
    mlp, p, g = int(D * cfg.mlp_ratio), cfg.patch_size, cfg.grid
    sd = {
        prefix + "patch_embed.proj.weight": G.rn(D, 3, p, p, std=(3 * p * p) ** -0.5),
        prefix + "patch_embed.proj.bias": G.rn(D),
        prefix + "pos_embed": G.rn(1, g, g, D),
        prefix + "neck.0.weight": G.rn(cfg.out_chans, D, 1, 1, std=D ** -0.5),
        prefix + "neck.1.weight": G.rn(cfg.out_chans, std=0.1, mean=1.0), prefix + "neck.1.bias": G.rn(cfg.out_chans, std=0.1),
        prefix + "neck.2.weight": G.rn(cfg.out_chans, cfg.out_chans, 3, 3, std=(9 * cfg.out_chans) ** -0.5),
        prefix + "neck.3.weight": G.rn(cfg.out_chans, std=0.1, mean=1.0), prefix + "neck.3.bias": G.rn(cfg.out_chans, std=0.1),
    }
    for i in range(cfg.depth):
        b = f"{prefix}blocks.{i}."
        S = g if i in cfg.global_attn_indexes else cfg.window_size
        sd.update({
            b + "norm1.weight": G.rn(D, std=0.1, mean=1.0), b + "norm1.bias": G.rn(D, std=0.1),
            b + "norm2.weight": G.rn(D, std=0.1, mean=1.0), b + "norm2.bias": G.rn(D, std=0.1),
            b + "attn.qkv.weight": G.rn(3 * D, D, std=D ** -0.5), b + "attn.qkv.bias": G.rn(3 * D, std=0.1),
            b + "attn.proj.weight": G.rn(D, D, std=0.5 * D ** -0.5), b + "attn.proj.bias": G.rn(D),
            # the reference zero-initialises these; N(0, .) so the rel-pos path is exercised
            b + "attn.rel_pos_h": G.rn(2 * S - 1, hd, std=0.1), b + "attn.rel_pos_w": G.rn(2 * S - 1, hd, std=0.1),
            b + "mlp.lin1.weight": G.rn(mlp, D, std=D ** -0.5), b + "mlp.lin1.bias": G.rn(mlp),
            b + "mlp.lin2.weight": G.rn(D, mlp, std=0.5 * mlp ** -0.5), b + "mlp.lin2.bias": G.rn(D),
        })
    return sd


def dinov2_state_dict(cfg: DinoCfg, seed: int, device, prefix: str = "model.visual_model_dinov2.") -> Dict[str, Tensor]:
    """hub `dinov2_vitl14` parameter names; LayerScale gammas O(1) (trained checkpoints are far from the 1e-5 init)."""
    G = _Gen(seed, device)
    D, p = cfg.embed_dim, cfg.patch_size
    mlp = int(D * cfg.mlp_ratio)
    sd = {
        prefix + "cls_token": G.rn(1, 1, D, std=0.5),
        prefix + "pos_embed": G.rn(1, 1 + cfg.train_grid ** 2, D, std=0.3),
        prefix + "mask_token": G.rn(1, D),
        prefix + "patch_embed.proj.weight": G.rn(D, 3, p, p, std=(3 * p * p) ** -0.5),
        prefix + "patch_embed.proj.bias": G.rn(D, std=0.1),
        prefix + "norm.weight": G.rn(D, std=0.1, mean=1.0), prefix + "norm.bias": G.rn(D, std=0.1),
    }
    for i in range(cfg.depth):
        b = f"{prefix}blocks.{i}."
        sd.update({
            b + "norm1.weight": G.rn(D, std=0.1, mean=1.0), b + "norm1.bias": G.rn(D, std=0.1),
            b + "norm2.weight": G.rn(D, std=0.1, mean=1.0), b + "norm2.bias": G.rn(D, std=0.1),
            b + "attn.qkv.weight": G.rn(3 * D, D, std=D ** -0.5), b + "attn.qkv.bias": G.rn(3 * D, std=0.1),
            b + "attn.proj.weight": G.rn(D, D, std=D ** -0.5), b + "attn.proj.bias": G.rn(D),
            b + "ls1.gamma": G.rn(D, std=0.1, mean=0.5), b + "ls2.gamma": G.rn(D, std=0.1, mean=0.5),
            b + "mlp.fc1.weight": G.rn(mlp, D, std=D ** -0.5), b + "mlp.fc1.bias": G.rn(mlp),
            b + "mlp.fc2.weight": G.rn(D, mlp, std=mlp ** -0.5), b + "mlp.fc2.bias": G.rn(D),
        })
    return sd


def clip_state_dict(cfg: ClipCfg, seed: int, device, prefix: str = "model.vision_tower.vision_tower.vision_model.") -> Dict[str, Tensor]:
    G = _Gen(seed, device)
    D, p = cfg.hidden, cfg.patch_size
    sd = {
        prefix + "embeddings.class_embedding": G.rn(D, std=0.5),
        prefix + "embeddings.patch_embedding.weight": G.rn(D, 3, p, p, std=(3 * p * p) ** -0.5),
        prefix + "embeddings.position_embedding.weight": G.rn(cfg.tokens, D, std=0.3),
        prefix + "pre_layrnorm.weight": G.rn(D, std=0.1, mean=1.0), prefix + "pre_layrnorm.bias": G.rn(D, std=0.1),
    }
    for i in range(cfg.layers):
        l = f"{prefix}encoder.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            sd[l + n + ".weight"] = G.rn(D, std=0.1, mean=1.0)
            sd[l + n + ".bias"] = G.rn(D, std=0.1)
        for n in ("q_proj", "k_proj", "v_proj"):
            sd[l + f"self_attn.{n}.weight"] = G.rn(D, D, std=D ** -0.5)
            sd[l + f"self_attn.{n}.bias"] = G.rn(D, std=0.1)
        sd[l + "self_attn.out_proj.weight"] = G.rn(D, D, std=0.5 * D ** -0.5)
        sd[l + "self_attn.out_proj.bias"] = G.rn(D)
        sd[l + "mlp.fc1.weight"], sd[l + "mlp.fc1.bias"] = G.rn(cfg.mlp, D, std=D ** -0.5), G.rn(cfg.mlp)
        sd[l + "mlp.fc2.weight"], sd[l + "mlp.fc2.bias"] = G.rn(D, cfg.mlp, std=0.5 * cfg.mlp ** -0.5), G.rn(D)
    return sd


def llama_state_dict(cfg: LlamaCfg, seed: int, device, prefix: str = "model.") -> Dict[str, Tensor]:
    G = _Gen(seed, device)
    D = cfg.hidden
    sd = {prefix + "embed_tokens.weight": G.rn(cfg.vocab, D, std=1.0), prefix + "norm.weight": G.rn(D, std=0.1, mean=1.0)}
    for i in range(cfg.layers):
        l = f"{prefix}layers.{i}."
        sd[l + "input_layernorm.weight"] = G.rn(D, std=0.1, mean=1.0)
        sd[l + "post_attention_layernorm.weight"] = G.rn(D, std=0.1, mean=1.0)
        for n in ("q_proj", "k_proj", "v_proj"):
            sd[l + f"self_attn.{n}.weight"] = G.rn(D, D, std=D ** -0.5)
        sd[l + "self_attn.o_proj.weight"] = G.rn(D, D, std=0.5 * D ** -0.5)
        sd[l + "mlp.gate_proj.weight"] = G.rn(cfg.mlp, D, std=D ** -0.5)
        sd[l + "mlp.up_proj.weight"] = G.rn(cfg.mlp, D, std=D ** -0.5)
        sd[l + "mlp.down_proj.weight"] = G.rn(D, cfg.mlp, std=0.5 * cfg.mlp ** -0.5)
    return sd


def selector_state_dict(hidden: int, seed: int, device, prefix: str = "model.") -> Dict[str, Tensor]:
    G = _Gen(seed, device)
    E, MLP = 256, 2048
    sd: Dict[str, Tensor] = {}

    def lin(name, out_f, in_f):
        sd[prefix + name + ".weight"] = G.rn(out_f, in_f, std=in_f ** -0.5)
        sd[prefix + name + ".bias"] = G.rn(out_f, std=0.05)

    def ln(name):
        sd[prefix + name + ".weight"] = G.rn(E, std=0.1, mean=1.0)
        sd[prefix + name + ".bias"] = G.rn(E, std=0.1)

    lin("text_hidden_fcs.0.0", hidden, hidden)
    lin("text_hidden_fcs.0.2", E, hidden)
    for i in range(2):
        p = f"lisa_attention_layers.{i}."
        for att in ("self_attn", "cross_attn_token_to_image", "cross_attn_image_to_token"):
            for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
                lin(p + att + "." + proj, E, E)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            ln(p + n)
        lin(p + "mlp.lin1", MLP, E)
        lin(p + "mlp.lin2", E, MLP)
    for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
        lin("lisa_final_attn." + proj, E, E)
    ln("lisa_norm_final_attn")
    lin("lisa_iou_head.0", 128, E)
    lin("lisa_iou_head.2", 1, 128)
    lin("lisa_embedding_head.0", 2048, E)
    lin("lisa_embedding_head.2", E, 2048)
    return sd


def sam_decoder_state_dict(seed: int, device, prefix: str = "model.visual_model.") -> Dict[str, Tensor]:
    """SAM prompt encoder + mask decoder (reference model/segment_anything/modeling/{prompt_encoder,mask_decoder,
    transformer}.py shapes and names).  The hyper-networks' output layers are scaled so that mask logits span +-20 like a
    trained decoder's (the stability score of the mask generator needs logits well beyond its +-1 offsets)."""
    G = _Gen(seed, device)
    E = 256
    sd: Dict[str, Tensor] = {}

    def lin(name, out_f, in_f, gain=1.0):
        sd[prefix + name + ".weight"] = G.rn(out_f, in_f, std=gain * in_f ** -0.5)
        sd[prefix + name + ".bias"] = G.rn(out_f, std=0.05)

    def ln(name, dim):
        sd[prefix + name + ".weight"] = G.rn(dim, std=0.1, mean=1.0)
        sd[prefix + name + ".bias"] = G.rn(dim, std=0.1)

    g = torch.Generator(device=device).manual_seed(seed + 99)
    sd[prefix + "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"] = torch.randn(2, E // 2, generator=g, device=device)
    for i in range(4):
        sd[prefix + f"prompt_encoder.point_embeddings.{i}.weight"] = G.rn(1, E, std=0.5)
    sd[prefix + "prompt_encoder.not_a_point_embed.weight"] = G.rn(1, E, std=0.5)
    sd[prefix + "prompt_encoder.no_mask_embed.weight"] = G.rn(1, E, std=0.5)
    t = "mask_decoder.transformer."
    for i in range(2):
        p = f"{t}layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(p + "self_attn." + proj, E, E)
        for att in ("cross_attn_token_to_image", "cross_attn_image_to_token"):
            for proj in ("q_proj", "k_proj", "v_proj"):
                lin(p + att + "." + proj, E // 2, E)
            lin(p + att + ".out_proj", E, E // 2)
        for n in ("norm1", "norm2", "norm3", "norm4"):
            ln(p + n, E)
        lin(p + "mlp.lin1", 2048, E)
        lin(p + "mlp.lin2", E, 2048)
    for proj in ("q_proj", "k_proj", "v_proj"):
        lin(t + "final_attn_token_to_image." + proj, E // 2, E)
    lin(t + "final_attn_token_to_image.out_proj", E, E // 2)
    ln(t + "norm_final_attn", E)
    d = "mask_decoder."
    sd[prefix + d + "iou_token.weight"] = G.rn(1, E, std=0.5)
    sd[prefix + d + "mask_tokens.weight"] = G.rn(4, E, std=0.5)
    sd[prefix + d + "output_upscaling.0.weight"] = G.rn(E, 64, 2, 2, std=E ** -0.5)
    sd[prefix + d + "output_upscaling.0.bias"] = G.rn(64, std=0.05)
    ln(d + "output_upscaling.1", 64)
    sd[prefix + d + "output_upscaling.3.weight"] = G.rn(64, 32, 2, 2, std=64 ** -0.5)
    sd[prefix + d + "output_upscaling.3.bias"] = G.rn(32, std=0.05)
    for i in range(4):
        for j, (o, k) in enumerate(((E, E), (E, E), (32, E))):
            lin(f"{d}output_hypernetworks_mlps.{i}.layers.{j}", o, k, gain=8.0 if j == 2 else 1.0)
    for j, (o, k) in enumerate(((E, E), (E, E), (4, E))):
        lin(f"{d}iou_prediction_head.layers.{j}", o, k)
    return sd


def lisa_state_dict(cfg: LisaCfg, seed: int = 0, device="cuda", with_lm_head: bool = False,
                    with_sam_decoder: bool = False) -> Dict[str, Tensor]:
    """Full random-init state dict with the reference's key names, bf16 on `device`.
    with_lm_head adds `lm_head.weight` (only the training forward reads it); with_sam_decoder adds SAM's prompt
    encoder + mask decoder (`model.visual_model.{prompt_encoder,mask_decoder}.*`: proposal generation)."""
    sd: Dict[str, Tensor] = {}
    if with_sam_decoder:
        sd.update(sam_decoder_state_dict(seed + 8, device))
    if with_lm_head:
        sd["lm_head.weight"] = _Gen(seed + 7, device).rn(cfg.llama.vocab, cfg.llama.hidden, std=cfg.llama.hidden ** -0.5)
    if cfg.image_encoder == "dinov2":
        sd.update(dinov2_state_dict(cfg.dino, seed + 1, device))
        G = _Gen(seed + 6, device)
        sd["model.lisa_dino_conv.weight"] = G.rn(cfg.dino.out_chans, cfg.dino.embed_dim, 1, 1, std=cfg.dino.embed_dim ** -0.5)
        sd["model.lisa_dino_conv.bias"] = G.rn(cfg.dino.out_chans, std=0.1)
    else:
        sd.update(sam_state_dict(cfg.sam, seed + 1, device))
    sd.update(clip_state_dict(cfg.clip, seed + 2, device))
    sd.update(llama_state_dict(cfg.llama, seed + 3, device))
    sd.update(selector_state_dict(cfg.llama.hidden, seed + 4, device))
    G = _Gen(seed + 5, device)
    sd["model.mm_projector.weight"] = G.rn(cfg.llama.hidden, cfg.clip.hidden, std=cfg.clip.hidden ** -0.5)
    sd["model.mm_projector.bias"] = G.rn(cfg.llama.hidden)
    return sd


def make_proposals(K: int, gen: torch.Generator, device, area_range: Tuple[float, float] = (0.01, 0.4)) -> Tensor:
    """K structured soft masks [K,256,256] in [0,1]: axis-aligned rectangles with log-uniform area in
    `area_range` (fraction of the image; default [1%, 40%]), blurred by a 3x3 box (mimics the antialiased resize of
    reference utils/dataset.py:620-622).  Small proposals (a few cells of the 64x64 feature grid) pool very
    different features each, which spreads the similarities: the margin-qualified index tests use them."""
    import math
    lo, hi = math.log(area_range[0]), math.log(area_range[1])
    area = torch.exp(torch.empty(K, device=device).uniform_(lo, hi, generator=gen))
    aspect = torch.exp(torch.empty(K, device=device).uniform_(-0.7, 0.7, generator=gen))
    h = (area * aspect).sqrt().clamp(max=1.0) * 256
    w = (area / aspect).sqrt().clamp(max=1.0) * 256
    cy = torch.rand(K, device=device, generator=gen) * (256 - h) + h / 2
    cx = torch.rand(K, device=device, generator=gen) * (256 - w) + w / 2
    yy = torch.arange(256, device=device).view(1, 256, 1).float() + 0.5
    xx = torch.arange(256, device=device).view(1, 1, 256).float() + 0.5
    m = ((yy - cy.view(K, 1, 1)).abs() <= h.view(K, 1, 1) / 2) & ((xx - cx.view(K, 1, 1)).abs() <= w.view(K, 1, 1) / 2)
    m = torch.nn.functional.avg_pool2d(m.float().unsqueeze(1), 3, 1, 1).squeeze(1)
    return m.to(BF16)


def make_inputs(cfg: LisaCfg, batch: int, n_props: int, t_text: int, seed: int = 1234, device="cuda",
                area_range: Tuple[float, float] = (0.01, 0.4)) -> dict:
    """ReasonSeg-shaped synthetic `input_dict` (keys of reference utils/dataset.py:150-170 that the
    forward reads).  Token layout per SURVEY §8d: [bos, .., <im_start>, IMAGE, <im_end>, text.., [SEG], '.', eos]."""
    assert t_text >= 8
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    S = cfg.dino.img_size if cfg.image_encoder == "dinov2" else cfg.sam.img_size
    Sc = cfg.clip.image_size
    images = torch.randn(batch, 3, S, S, generator=g, device=dev).to(BF16)
    images_clip = torch.randn(batch, 3, Sc, Sc, generator=g, device=dev).to(BF16)
    ids = torch.randint(3, 31999, (batch, t_text), generator=g, device=dev, dtype=torch.int64)
    ids[:, 0] = 1
    ids[:, 1], ids[:, 2], ids[:, 3] = 32001, IMAGE_TOKEN_INDEX, 32002
    ids[:, t_text - 3] = cfg.seg_token_idx
    ids[:, t_text - 2] = 29889
    ids[:, t_text - 1] = 2
    return {
        "images": images, "images_clip": images_clip, "input_ids": ids, "labels": ids.clone(),
        "attention_masks": torch.ones(batch, t_text, dtype=torch.bool, device=dev),
        "offset": torch.arange(batch + 1), "masks_list": [None] * batch, "label_list": [None] * batch,
        "resize_list": [(S, S)] * batch,
        "sam_segs_list": [make_proposals(n_props, g, dev, area_range) for _ in range(batch)],
        "sam_ious_list": None, "sam_iops_list": None, "inference": True,
    }


def make_train_inputs(cfg: LisaCfg, convs_per_image, n_props, t_text: int, seed: int = 4321, device="cuda") -> dict:
    """Training-shaped `input_dict` (reference utils/dataset.py:150-170): image i carries convs_per_image[i]
    conversations (rounds) of t_text tokens, each ending `.. [SEG] . </s>` (right-padded variants get shorter
    answers), labels = IGNORE over the prompt and the ids over the answer, n_props[i] proposals with
    ground-truth IoU / IoP rows per round (`sam_ious_list[i]`: [R_i, K_i], bf16-representable values)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    B = len(convs_per_image)
    N = sum(convs_per_image)
    inp = make_inputs(cfg, B, 8, t_text, seed=seed, device=device)
    ids = torch.randint(3, 31999, (N, t_text), generator=g, device=dev, dtype=torch.int64)
    ids[:, 0] = 1
    ids[:, 1], ids[:, 2], ids[:, 3] = 32001, IMAGE_TOKEN_INDEX, 32002
    labels = torch.full_like(ids, -100)
    mask = torch.ones(N, t_text, dtype=torch.bool, device=dev)
    for n in range(N):
        end = t_text - (n % 3) * 2           # conversations of different lengths, right padded
        ids[n, end - 3], ids[n, end - 2], ids[n, end - 1] = cfg.seg_token_idx, 29889, 2
        ids[n, end:] = 0
        mask[n, end:] = False
        labels[n, end - 6:end] = ids[n, end - 6:end]      # the assistant's answer is the supervised span
    off = [0]
    for c in convs_per_image:
        off.append(off[-1] + c)
    q = lambda t: t.to(BF16).float()
    inp.update({
        "input_ids": ids, "labels": labels, "attention_masks": mask, "offset": torch.tensor(off),
        "sam_segs_list": [make_proposals(k, g, dev) for k in n_props],
        "sam_ious_list": [q(torch.rand(c, k, generator=g, device=dev)) for c, k in zip(convs_per_image, n_props)],
        "sam_iops_list": [q(torch.rand(c, k, generator=g, device=dev)) for c, k in zip(convs_per_image, n_props)],
        "inference": False,
    })
    return inp
