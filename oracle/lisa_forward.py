"""ORACLE (test infrastructure, never on the product path).

Restatement of `LISAForCausalLM.model_forward(inference=True)` (reference model/LISA.py:225-414),
variant A of SURVEY §0/T1: image features from the SAM ViT-H encoder (`get_visual_embs`,
LISA.py:173-184), text state from LLaVA (CLIP tower → mm_projector → splice → LLaMA,
reference llava_arch.py:93-347, llava_llama.py:55-135), selector per image.

The reference class itself cannot be imported in the build container (missing skimage, hard .cuda()
calls, hub download, transformers-5 incompatibilities — SURVEY §8c), so this glue is a line-by-line
restatement; the heavy sub-modules it calls (oracle/sam_encoder.py, oracle/selector.py) ARE pinned
against the reference's importable modules by oracle/make_golden.py.

State-dict layout (reference names):
  model.visual_model.image_encoder.*      SAM ViT-H
  model.vision_tower.vision_tower.*       CLIP (vision_model.*)
  model.mm_projector.{weight,bias}
  model.embed_tokens.weight, model.layers.*, model.norm.weight
  model.text_hidden_fcs.0.{0,2}.*, model.lisa_*   selector
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import clip_llama, dinov2, sam_encoder, selector

Tensor = torch.Tensor

IMAGE_TOKEN_INDEX = -200  # reference utils/utils.py:12
SEG_TOKEN_IDX = 32000     # `[SEG]` is the first added token (reference training.py:121-137)


@dataclass
class LisaConfig:
    sam: sam_encoder.SamConfig = field(default_factory=sam_encoder.SamConfig)
    clip: clip_llama.ClipConfig = field(default_factory=clip_llama.ClipConfig)
    llama: clip_llama.LlamaConfig = field(default_factory=clip_llama.LlamaConfig)
    dino: dinov2.Dinov2Config = field(default_factory=dinov2.Dinov2Config)
    image_encoder: str = "sam"   # "sam" = variant A (LISA.py:173-184); "dinov2" = variant B (LISA.py:186-199,244-245)
    seg_token_idx: int = SEG_TOKEN_IDX
    out_dim: int = 256

    @property
    def n_image_tokens(self) -> int:
        return self.clip.tokens - 1


def sub_dict(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def encode_images(images_clip: Tensor, sd: Dict[str, Tensor], cfg: LisaConfig) -> Tensor:
    """CLIP patch features → mm_projector (llava_arch.py:93-96)."""
    feats = clip_llama.clip_patch_features(images_clip, sub_dict(sd, "model.vision_tower.vision_tower."), cfg.clip)
    return F.linear(feats, sd["model.mm_projector.weight"], sd["model.mm_projector.bias"])


def splice_inputs(input_ids: Tensor, attention_mask: Tensor, image_feats: Tensor, embed: Tensor):
    """`prepare_inputs_labels_for_multimodal` for the only layout LLM-Seg produces: one IMAGE token per
    row, `mm_use_im_start_end=True` (llava_arch.py:185-208,230-245,332-345).  Returns
    (embeds [N,T,D], mask [N,T]) with T = T_text + n_img - 1 and (n_img-1) True values PREPENDED."""
    rows = []
    for n in range(input_ids.shape[0]):
        ids = input_ids[n]
        pos = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten()
        assert pos.numel() == 1, "exactly one <image> token per conversation on this path"
        i = int(pos[0])
        parts = [F.embedding(ids[:i], embed), image_feats[n], F.embedding(ids[i + 1:i + 2], embed),
                 F.embedding(ids[i + 2:], embed)]
        rows.append(torch.cat(parts, dim=0))
    embeds = torch.stack(rows, dim=0)
    extra = embeds.shape[1] - input_ids.shape[1]
    left = torch.ones((attention_mask.shape[0], extra), dtype=attention_mask.dtype, device=attention_mask.device)
    return embeds, torch.cat([left, attention_mask], dim=1)


def seg_token_mask(input_ids: Tensor, cfg: LisaConfig) -> Tensor:
    """Shift-by-one [SEG] mask with the (n_img-1)-token image offset (LISA.py:254-266)."""
    m = input_ids[:, 1:] == cfg.seg_token_idx
    m = torch.cat([m, torch.zeros((m.shape[0], 1), dtype=torch.bool, device=m.device)], dim=1)
    return torch.cat([torch.zeros((m.shape[0], cfg.n_image_tokens - 1), dtype=torch.bool, device=m.device), m], dim=1)


def image_features(sd: Dict[str, Tensor], cfg: LisaConfig, images: Tensor) -> Tensor:
    """[B,3,S,S] -> [B,256,64,64]: variant A `get_visual_embs` (LISA.py:173-184, commented out at :242) or
    variant B `lisa_dino_conv(get_dinov2_visual_embs(images))` (LISA.py:244-245, the checked-in branch)."""
    if cfg.image_encoder == "dinov2":
        return dinov2.image_embeddings(images, sub_dict(sd, "model.visual_model_dinov2."),
                                       sd["model.lisa_dino_conv.weight"], sd["model.lisa_dino_conv.bias"], cfg.dino)
    return sam_encoder.image_encoder(images, sub_dict(sd, "model.visual_model.image_encoder."), cfg.sam)


def model_forward_inference(sd: Dict[str, Tensor], cfg: LisaConfig, *, images: Tensor, images_clip: Tensor,
                            input_ids: Tensor, attention_masks: Tensor, offset: Tensor,
                            sam_segs_list: List[Tensor], masks_list: Optional[list] = None) -> dict:
    """One reference inference call (batch of ONE image, LISA.py:271).  Returns the reference's dict."""
    assert images_clip.shape[0] == 1, "reference inference is one image per forward (LISA.py:271)"
    image_embeddings = image_features(sd, cfg, images)
    assert image_embeddings.shape[0] == len(offset) - 1
    seg_mask = seg_token_mask(input_ids, cfg)

    n_conv = input_ids.shape[0]
    clip_in = images_clip.expand(n_conv, -1, -1, -1).contiguous()
    feats = encode_images(clip_in, sd, cfg)
    embeds, mask = splice_inputs(input_ids, attention_masks, feats, sd["model.embed_tokens.weight"])
    hidden = clip_llama.llama_last_hidden(embeds, mask, sub_dict(sd, "model."), cfg.llama)

    sel_sd = sub_dict(sd, "model.")
    last = selector.text_hidden_fc(hidden, sel_sd)            # all positions (LISA.py:317-318)
    pred = last[seg_mask]                                      # [n_seg, 256]
    counts = seg_mask.int().sum(-1)
    seg_off = torch.cat([torch.zeros(1, dtype=torch.long, device=counts.device), counts.cumsum(-1)], dim=0)[offset]
    pred_list = [pred[int(seg_off[i]):int(seg_off[i + 1])] for i in range(len(seg_off) - 1)]

    emb_up = selector.upsample_embeddings(image_embeddings)    # LISA.py:350-354
    sims, ious = [], []
    for b in range(len(sam_segs_list)):
        s, u = selector.selector_forward(emb_up[b], sam_segs_list[b], pred_list[b], sel_sd)
        sims.append(s)
        ious.append(u)
    return {"pred_similarity": sims, "gt_masks": masks_list, "pred_iou": ious}


def forward_batched(sd, cfg, *, images, images_clip, input_ids, attention_masks, sam_segs_list) -> dict:
    """Batched inference as defined in SURVEY §0/T6: B independent reference batch-1 calls."""
    sims, ious = [], []
    one = torch.arange(2, device=input_ids.device)
    for b in range(images.shape[0]):
        out = model_forward_inference(sd, cfg, images=images[b:b + 1], images_clip=images_clip[b:b + 1],
                                      input_ids=input_ids[b:b + 1], attention_masks=attention_masks[b:b + 1],
                                      offset=one, sam_segs_list=[sam_segs_list[b]])
        sims += out["pred_similarity"]
        ious += out["pred_iou"]
    return {"pred_similarity": sims, "gt_masks": None, "pred_iou": ious}


IGNORE_INDEX = -100       # reference utils/utils.py:11


def splice_labels(input_ids: Tensor, labels: Tensor, n_img: int) -> Tensor:
    """Labels of `prepare_inputs_labels_for_multimodal` (llava_arch.py:185-208,230-245): the IMAGE position
    is replaced by n_img IGNORE entries, everything else keeps its order."""
    rows = []
    for n in range(input_ids.shape[0]):
        i = int((input_ids[n] == IMAGE_TOKEN_INDEX).nonzero().flatten()[0])
        ign = torch.full((n_img,), IGNORE_INDEX, dtype=labels.dtype, device=labels.device)
        rows.append(torch.cat([labels[n, :i], ign, labels[n, i + 1:i + 2], labels[n, i + 2:]], dim=0))
    return torch.stack(rows, dim=0)


def model_forward_training(sd: Dict[str, Tensor], cfg: LisaConfig, *, images: Tensor, images_clip: Tensor,
                           input_ids: Tensor, labels: Tensor, attention_masks: Tensor, offset: Tensor,
                           sam_segs_list: List[Tensor], sam_ious_list: List[Tensor], sam_iops_list: List[Tensor],
                           ce_loss_weight: float = 1.0, align_loss_weight: float = 1.0,
                           regression_loss_weight: float = 1.0) -> dict:
    """`model_forward(inference=False)` (LISA.py:243-266,292-392,416-474) with the LLaVA forward + CE of
    llava_llama.py:83-118: returns {"loss","ce_loss","align_loss","regression_loss"}."""
    image_embeddings = image_features(sd, cfg, images)
    assert image_embeddings.shape[0] == len(offset) - 1
    seg_mask = seg_token_mask(input_ids, cfg)
    # one CLIP image per conversation (LISA.py:293-303)
    clip_in = torch.cat([images_clip[i:i + 1].expand(int(offset[i + 1] - offset[i]), -1, -1, -1)
                         for i in range(len(offset) - 1)], dim=0).contiguous()
    feats = encode_images(clip_in, sd, cfg)
    embeds, mask = splice_inputs(input_ids, attention_masks, feats, sd["model.embed_tokens.weight"])
    new_labels = splice_labels(input_ids, labels, feats.shape[1])
    hidden = clip_llama.llama_last_hidden(embeds, mask, sub_dict(sd, "model."), cfg.llama)
    logits = F.linear(hidden, sd["lm_head.weight"])
    ce_loss = F.cross_entropy(logits[..., :-1, :].reshape(-1, logits.shape[-1]), new_labels[..., 1:].reshape(-1))

    sel_sd = sub_dict(sd, "model.")
    last = selector.text_hidden_fc(hidden, sel_sd)
    pred = last[seg_mask]
    counts = seg_mask.int().sum(-1)
    seg_off = torch.cat([torch.zeros(1, dtype=torch.long, device=counts.device), counts.cumsum(-1)], dim=0)[offset]
    pred_list = [pred[int(seg_off[i]):int(seg_off[i + 1])] for i in range(len(seg_off) - 1)]

    emb_up = selector.upsample_embeddings(image_embeddings)
    align_loss, regression_loss, valid_batch = 0.0, 0.0, 0
    for b in range(len(sam_segs_list)):
        gt_iou, gt_iop = sam_ious_list[b], sam_iops_list[b]
        rounds = pred_list[b].shape[0]
        if rounds == 0:   # LISA.py:435-437 (raised after the selector there; nothing observable happens in between)
            raise ValueError("number of rounds = 0; gt_iou.shape: {}".format(gt_iou.shape))
        segs_feature, pred_iou = selector.selector_features(emb_up[b], sam_segs_list[b], pred_list[b], sel_sd)
        a_r, r_r = 0.0, 0.0
        for r in range(rounds):
            g_iou = gt_iou[r].unsqueeze(1).to(pred_iou.dtype)
            g_iop = gt_iop[r].unsqueeze(1).to(pred_iou.dtype)
            a_r = a_r + softmax_align_loss(segs_feature[r], pred_list[b][r].unsqueeze(0), g_iou)
            r_r = r_r + iou_regression_loss(pred_iou[r], g_iop)
        valid_batch += 1
        align_loss = align_loss + a_r / (rounds + 1e-8)
        regression_loss = regression_loss + r_r / (rounds + 1e-8)
    if valid_batch > 0:
        align_loss = align_loss / valid_batch
        regression_loss = regression_loss / valid_batch
    ce_loss = ce_loss * ce_loss_weight
    align_loss = align_loss * align_loss_weight
    regression_loss = regression_loss * regression_loss_weight
    return {"loss": ce_loss + align_loss + regression_loss, "ce_loss": ce_loss, "align_loss": align_loss,
            "regression_loss": regression_loss}


# ---- losses (training only; reference model/loss.py) -------------------------------------------
def dice_loss(inputs: Tensor, targets: Tensor, num_masks: float, scale: float = 1000, eps: float = 1e-6) -> Tensor:
    """loss.py:4-30."""
    p = inputs.sigmoid().flatten(1, 2)
    t = targets.flatten(1, 2)
    num = 2 * (p / scale * t).sum(-1)
    den = (p / scale).sum(-1) + (t / scale).sum(-1)
    return (1 - (num + eps) / (den + eps)).sum() / (num_masks + 1e-8)


def sigmoid_ce_loss(inputs: Tensor, targets: Tensor, num_masks: float) -> Tensor:
    """loss.py:33-47."""
    l = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    return l.flatten(1, 2).mean(1).sum() / (num_masks + 1e-8)


def softmax_align_loss(proposal_embeds: Tensor, target_embed: Tensor, gt_ious: Tensor,
                       temperature: float = 0.05) -> Tensor:
    """KL( softmax(iou/τ) ‖ softmax(cos_sim/τ) ) over the K proposals, reduction sum (loss.py:50-80).
    proposal_embeds [K,D], target_embed [1,D], gt_ious [K,1]."""
    pe = proposal_embeds / proposal_embeds.norm(dim=-1, keepdim=True)
    te = target_embed / target_embed.norm(dim=-1, keepdim=True)
    sim = pe @ te.t()
    p_sim = F.softmax(sim / temperature, dim=0)
    p_iou = F.softmax(gt_ious / temperature, dim=0)
    return F.kl_div(p_sim.log(), p_iou, reduction="sum")


def iou_regression_loss(pred: Tensor, gt: Tensor) -> Tensor:
    """mean((p-g)^2 * exp(g-1)) * 50  (loss.py:82-94)."""
    p, g = pred.flatten(), gt.flatten()
    return ((p - g) ** 2 * torch.exp(g - 1.0)).mean() * 50
