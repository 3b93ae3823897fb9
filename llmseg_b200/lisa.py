"""Drop-in host class for the LLM-Seg inference forward:  `LISAForCausalLM.forward(**input_dict)` →
`model_forward(...)` with the reference's argument names and return dict
(reference model/LISA.py:220-241,410-414), executed entirely by the sm_100a kernels.

Variant A of SURVEY §0/T1: image features come from the SAM ViT-H encoder (`get_visual_embs`,
LISA.py:173-184) — the encoder `north_star` names and the one fully in-tree.

Differences from the reference that a caller can observe:
  * batched inference is allowed (reference asserts one image per call, LISA.py:271); a batch of B
    images with one conversation each is defined as B independent reference calls
  * `inference=False` returns the reference's loss dict as forward VALUES (fp32 0-d tensors): the kernels
    build no autograd graph, so this is the evaluation of the training objective, not a trainable step
  * extra keys (`best_index`, padded logits) are added to the returned dict; the reference keys are unchanged
  * the launch sequence of each input shape is captured once into a CUDA graph and replayed
    (`use_cuda_graph=True`): ~2300 launches per forward would otherwise be CPU-bound at batch 1
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops
from .encoders import BF16, ClipTower, Dinov2Encoder, LlamaDecoder, SamEncoder
from .selector import Selector

Tensor = torch.Tensor

IMAGE_TOKEN_INDEX = -200  # reference utils/utils.py:12
DEFAULT_SEG_TOKEN_IDX = 32000


@dataclass
class SamCfg:
    img_size: int = 1024
    patch_size: int = 16
    embed_dim: int = 1280
    depth: int = 32
    num_heads: int = 16
    mlp_ratio: float = 4.0
    out_chans: int = 256
    window_size: int = 14
    global_attn_indexes: tuple = (7, 15, 23, 31)
    ln_eps: float = 1e-6

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size


@dataclass
class DinoCfg:
    """hub `dinov2_vitl14` evaluated on 896x896 images (64x64 patches, reference LISA.py:186-199)."""
    img_size: int = 896
    patch_size: int = 14
    embed_dim: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_ratio: float = 4.0
    train_grid: int = 37
    ln_eps: float = 1e-6
    out_chans: int = 256
    interpolate_offset: float = 0.1

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size


@dataclass
class ClipCfg:
    image_size: int = 224
    patch_size: int = 14
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    eps: float = 1e-5
    select_layer: int = -2

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


@dataclass
class LlamaCfg:
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


@dataclass
class LisaCfg:
    sam: SamCfg = field(default_factory=SamCfg)
    clip: ClipCfg = field(default_factory=ClipCfg)
    llama: LlamaCfg = field(default_factory=LlamaCfg)
    dino: DinoCfg = field(default_factory=DinoCfg)
    # "sam": image features from SAM ViT-H (`get_visual_embs`, LISA.py:173-184 — the encoder north_star names);
    # "dinov2": DINOv2 ViT-L/14 + lisa_dino_conv (`get_dinov2_visual_embs`, LISA.py:186-199,244-245 — the
    # branch the checked-in reference and its released checkpoints use)
    image_encoder: str = "sam"
    seg_token_idx: int = DEFAULT_SEG_TOKEN_IDX
    out_dim: int = 256


def _sub(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def strip_peft_prefix(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Accept checkpoints saved through PEFT (`base_model.model.` prefix, reference training.py:194-237)
    and merge LoRA deltas W + (alpha/r)·B·A into q_proj/v_proj when present (r=8, alpha=16)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("base_model.model."):
            k = k[len("base_model.model."):]
        out[k] = v
    lora_a = {k: v for k, v in out.items() if ".lora_A." in k}
    for ka, a in lora_a.items():
        kb = ka.replace(".lora_A.", ".lora_B.")
        base = ka.split(".lora_A.")[0] + ".weight"
        if kb in out and base in out:
            r = a.shape[0]
            out[base] = out[base].float() + (16.0 / r) * (out[kb].float() @ a.float())
    return {k: v for k, v in out.items() if ".lora_" not in k}


def select_proposals(output_dict: dict, threshold: Optional[float] = None):
    """The two selection rules the reference's validation loops apply to a forward's outputs:
    `validate` keeps the proposal with the highest similarity (training.py:627-629) — here the argmax fused into
    the select kernel (`best_index`, falling back to argmax of `pred_similarity` for a plain reference dict) — and
    `validate_threshold` keeps every proposal whose predicted IoU exceeds `threshold` (training.py:712-718).
    Returns one `(best, kept)` pair per image; `kept` is None without a threshold."""
    sims, ious = output_dict["pred_similarity"], output_dict["pred_iou"]
    best = output_dict.get("best_index")
    best = [int(v) for v in best.tolist()] if best is not None else [int(torch.argmax(s_)) for s_ in sims]
    kept = None
    if threshold is not None:
        kept = [[i for i, v in enumerate(u[0].float().tolist()) if v > threshold] for u in ious]
    return [(best[i], None if kept is None else kept[i]) for i in range(len(sims))]


class LISAForCausalLM:
    """B200 implementation of the reference class of the same name (inference forward only)."""

    def __init__(self, state_dict: Dict[str, Tensor], cfg: Optional[LisaCfg] = None, *,
                 device: str = "cuda:0", seg_token_idx: Optional[int] = None, max_seq: int = 1024,
                 use_cuda_graph: bool = True, ce_loss_weight: float = 1.0, align_loss_weight: float = 1.0,
                 regression_loss_weight: float = 1.0):
        if not torch.cuda.is_available():
            raise RuntimeError("llmseg_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        from . import _lib
        _lib.lib()  # fail loudly now if the CUDA library is missing
        self.cfg = cfg or LisaCfg()
        if seg_token_idx is not None:
            self.cfg.seg_token_idx = seg_token_idx
        self.seg_token_idx = self.cfg.seg_token_idx
        self.device = torch.device(device)
        sd = strip_peft_prefix(state_dict)
        if self.cfg.image_encoder == "sam":
            self.sam = SamEncoder(_sub(sd, "model.visual_model.image_encoder."), self.cfg.sam, self.device)
            self.image_encoder, self.image_size = self.sam, self.cfg.sam.img_size
        elif self.cfg.image_encoder == "dinov2":
            self.dino = Dinov2Encoder(_sub(sd, "model.visual_model_dinov2."), self.cfg.dino, self.device,
                                      sd["model.lisa_dino_conv.weight"], sd["model.lisa_dino_conv.bias"])
            self.image_encoder, self.image_size = self.dino, self.cfg.dino.img_size
        else:
            raise ValueError(f"image_encoder must be 'sam' or 'dinov2', got {self.cfg.image_encoder!r}")
        self.clip = ClipTower(_sub(sd, "model.vision_tower.vision_tower."), self.cfg.clip, self.device,
                              sd["model.mm_projector.weight"], sd["model.mm_projector.bias"])
        # loss weights of the training forward (reference LISA.py:158-160, training.py defaults 1.0)
        self.ce_loss_weight, self.align_loss_weight = ce_loss_weight, align_loss_weight
        self.regression_loss_weight = regression_loss_weight
        self.llama = LlamaDecoder(_sub(sd, "model."), self.cfg.llama, self.device, max_seq=max_seq,
                                  lm_head=sd.get("lm_head.weight"))
        self.selector = Selector(_sub(sd, "model."), self.device)
        self.use_cuda_graph = use_cuda_graph
        self.overlap_branches = os.environ.get("LLMSEG_OVERLAP", "1") != "0"
        self._side_stream = torch.cuda.Stream(device=self.device)    # text branch
        self._copy_stream = torch.cuda.Stream(device=self.device)    # image / proposal staging copies
        self._cap_stream = torch.cuda.Stream(device=self.device)     # capture stream of the image / selector graphs
        self._plans: Dict[tuple, dict] = {}
        self.last_forward_launches = 0

    # ---- reference API -------------------------------------------------------------------------
    def forward(self, **kwargs):
        if "past_key_values" in kwargs:
            raise NotImplementedError("the HF generate() path (reference LISA.py:221-222) is outside the hot path")
        return self.model_forward(**kwargs)

    __call__ = forward

    def get_visual_embs(self, pixel_values: Tensor) -> Tensor:
        """SAM ViT-H features, returned NCHW [B,256,64,64] like reference LISA.py:173-184."""
        if self.cfg.image_encoder != "sam":
            raise RuntimeError("this model was built with image_encoder='dinov2'; use get_dinov2_visual_embs")
        tok = self.sam.forward(pixel_values.to(self.device, BF16))
        g = self.cfg.sam.grid
        return tok.view(tok.shape[0], g, g, -1).permute(0, 3, 1, 2)

    def get_dinov2_visual_embs(self, pixel_values: Tensor) -> Tensor:
        """DINOv2 `x_norm_patchtokens`, returned NCHW [B,1024,64,64] like reference LISA.py:186-199
        (before `lisa_dino_conv`; the forward itself runs norm + conv as one GEMM)."""
        if self.cfg.image_encoder != "dinov2":
            raise RuntimeError("this model was built with image_encoder='sam'; use get_visual_embs")
        tok = self.dino.patch_tokens(pixel_values.to(self.device, BF16))
        g = self.cfg.dino.grid
        return tok.view(tok.shape[0], g, g, -1).permute(0, 3, 1, 2)

    @torch.no_grad()
    def model_forward(self, images: Tensor, images_clip: Tensor, input_ids: Tensor, labels: Optional[Tensor] = None,
                      attention_masks: Optional[Tensor] = None, offset: Optional[Tensor] = None,
                      masks_list: Optional[list] = None, label_list: Optional[list] = None,
                      resize_list: Optional[list] = None, sam_segs_list: Optional[List[Tensor]] = None,
                      sam_ious_list: Optional[list] = None, sam_iops_list: Optional[list] = None,
                      inference: bool = False, **kwargs) -> dict:
        if not inference:
            return self._train_forward(images, images_clip, input_ids, labels, attention_masks, offset,
                                       sam_segs_list, sam_ious_list, sam_iops_list)
        dev = self.device
        B, N, Tt = images.shape[0], input_ids.shape[0], input_ids.shape[1]
        if offset is None:
            offset = torch.arange(B + 1)
        assert B == len(offset) - 1, "batch_size == len(offset) - 1 (reference LISA.py:250)"
        off = tuple(int(v) for v in offset.tolist())
        assert off[-1] == N and sam_segs_list is not None and len(sam_segs_list) == B
        Ks = tuple(int(s.shape[0]) for s in sam_segs_list)
        n_clip = images_clip.shape[0]
        key = (B, N, Tt, Ks, off, n_clip, attention_masks is None)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._make_plan(key)
            self._plans[key] = plan
        # ---- stage the inputs into the plan's static device buffers (H2D or D2D copies) and run.  The small
        # inputs (CLIP image, ids, mask: all the text branch needs) go first on the current stream; the 1024 px
        # images and the proposals follow on a copy stream in the order they are needed, so the text branch starts
        # under the image copy and the proposals (67 MB at batch 8, read only by the selector) arrive under the
        # encoders instead of in front of them.
        st = plan["static"]
        cur = torch.cuda.current_stream()
        st["images_clip"].copy_(images_clip, non_blocking=True)
        st["input_ids"].copy_(input_ids, non_blocking=True)
        if attention_masks is not None:
            st["mask"].copy_(attention_masks, non_blocking=True)
        cp = self._copy_stream if self.overlap_branches else cur
        if cp is not cur:
            cp.wait_stream(cur)      # the previous forward (enqueued on cur) has finished reading the static buffers
        with torch.cuda.stream(cp):
            st["images"].copy_(images, non_blocking=True)
            ev_img = cp.record_event() if cp is not cur else None
            r0 = 0
            for sgs, kk in zip(sam_segs_list, Ks):
                st["segs"][r0:r0 + kk].copy_(sgs, non_blocking=True)
                r0 += kk
            ev_segs = cp.record_event() if cp is not cur else None
        # ---- CUDA-graph replay (three graphs captured on first use of this shape) or eager launches
        if self.use_cuda_graph:
            if plan["graphs"] is None:
                self._capture(plan)
            sim, iou, best = self._pipeline(plan, ev_img, ev_segs, plan["graphs"])
        else:
            from . import _lib
            n0 = _lib.launch_count()
            sim, iou, best = self._pipeline(plan, ev_img, ev_segs, None)
            plan["launches"] = _lib.launch_count() - n0
        self.last_forward_launches = plan["launches"]
        pred_similarity = [sim[i:i + 1, :Ks[i]].to(BF16) for i in range(B)]
        pred_iou = [iou[i:i + 1, :Ks[i]].to(BF16) for i in range(B)]
        return {"pred_similarity": pred_similarity, "gt_masks": masks_list, "pred_iou": pred_iou,
                "best_index": best.clone(), "similarity_padded": sim.clone(), "iou_padded": iou.clone()}

    # ---- training forward (loss values only; no autograd graph is built) ---------------------------
    def _train_forward(self, images, images_clip, input_ids, labels, attention_masks, offset, sam_segs_list,
                       sam_ious_list, sam_iops_list) -> dict:
        """`model_forward(inference=False)` (reference LISA.py:243-266,292-392,416-474 and the LLaVA forward + CE of
        llava_llama.py:83-118): {"loss","ce_loss","align_loss","regression_loss"} as 0-d fp32 tensors.  Eager
        launches (shapes vary per step; the [SEG] positions are read on the host like the reference's boolean
        indexing does).  Exact shortcuts: CLIP runs once per image instead of once per conversation, mask pooling
        once per image instead of once per round, text_hidden_fcs on the [SEG] rows only, the final RMSNorm folded
        into the lm_head GEMM, and the label splice done by index arithmetic inside the CE kernel."""
        dev, cfg = self.device, self.cfg
        if labels is None or sam_segs_list is None or sam_ious_list is None or sam_iops_list is None:
            raise ValueError("the training forward needs labels, sam_segs_list, sam_ious_list and sam_iops_list")
        B, N, Tt = images.shape[0], input_ids.shape[0], input_ids.shape[1]
        if offset is None:
            offset = torch.arange(B + 1)
        assert B == len(offset) - 1, "batch_size == len(offset) - 1 (reference LISA.py:250)"
        off = [int(v) for v in offset.tolist()]
        assert off[-1] == N and len(sam_segs_list) == B and images_clip.shape[0] == B
        # [SEG] rows in row-major (conversation, position) order; the hidden state that predicts [SEG] sits at
        # spliced position j + 255 for text index j with ids[j+1] == [SEG] (LISA.py:254-266)
        ids_host = input_ids.detach().cpu()
        n_img = self.cfg.clip.tokens - 1
        T = Tt + n_img - 1
        hits = (ids_host[:, 1:] == self.seg_token_idx).nonzero()
        seg_rows = (hits[:, 0] * T + hits[:, 1] + (n_img - 1)).to(torch.int32)
        per_conv = torch.bincount(hits[:, 0], minlength=N).tolist()
        rounds = [sum(per_conv[off[i]:off[i + 1]]) for i in range(B)]
        for i, r in enumerate(rounds):
            if r == 0:   # reference LISA.py:435-437
                raise ValueError("number of rounds = 0; gt_iou.shape: {}".format(tuple(sam_ious_list[i].shape)))
        Ks = [int(s.shape[0]) for s in sam_segs_list]
        # ---- encoders
        emb_tokens = self.image_encoder.forward(images.to(dev, BF16))
        feats = self.clip.forward(images_clip.to(dev, BF16))
        conv_image = torch.tensor([i for i in range(B) for _ in range(off[i + 1] - off[i])], device=dev)
        feats = feats.index_select(0, conv_image).contiguous()
        ids = input_ids.to(dev).contiguous()
        mask = None if attention_masks is None else attention_masks.to(dev)
        embeds, kv_len, _ = ops.embed_splice(ids, mask, self.llama.embed, feats, image_token=IMAGE_TOKEN_INDEX,
                                             seg_token=self.seg_token_idx)
        hidden, logits = self.llama.forward(embeds, N, T, kv_len, out_rows=seg_rows.to(dev), with_logits=True)
        ce2, _ = ops.lm_cross_entropy(logits, ids, labels.to(dev), n_img_tokens=n_img, vocab=cfg.llama.vocab,
                                      image_token=IMAGE_TOKEN_INDEX)
        text_embed = self.selector.text_embed(hidden)                       # [G,256], G = total rounds
        # ---- selector over (image, round) groups sharing their image's pooled proposal features
        plan_img = self.selector.make_plan(Ks)
        segs = torch.cat([s.to(dev, BF16) for s in sam_segs_list], dim=0)
        pooled = ops.maskpool(segs.contiguous(), emb_tokens.contiguous(), plan_img["mask_image"])
        group_image = [i for i in range(B) for _ in range(rounds[i])]
        k0 = [0]
        for kk in Ks:
            k0.append(k0[-1] + kk)
        rep = torch.cat([torch.arange(k0[i], k0[i + 1]) for i in group_image]).to(dev)
        plan_g = self.selector.make_plan([Ks[i] for i in group_image])
        sim, iou, _ = self.selector.forward_pooled(pooled.index_select(0, rep).contiguous(), text_embed, plan_g)
        # ---- losses: ground-truth IoU / IoP rows per round, cast to the model dtype like LISA.py:442-444
        G, kmax = len(group_image), plan_g["kmax"]
        gt = torch.zeros((2, G, kmax), dtype=torch.float32)
        g = 0
        for i in range(B):
            for r in range(rounds[i]):
                gt[0, g, :Ks[i]] = sam_ious_list[i][r].detach().to("cpu", BF16).float()
                gt[1, g, :Ks[i]] = sam_iops_list[i][r].detach().to("cpu", BF16).float()
                g += 1
        gt = gt.to(dev)
        valid = sum(1 for r in rounds if r > 0)
        gw = torch.tensor([1.0 / ((rounds[i] + 1e-8) * valid) for i in group_image], dtype=torch.float32, device=dev)
        out4, _ = ops.selector_losses(sim, iou, gt[0], gt[1], plan_g["k_off"], gw, ce=ce2,
                                      weights=(self.ce_loss_weight, self.align_loss_weight, self.regression_loss_weight))
        return {"loss": out4[0], "ce_loss": out4[1], "align_loss": out4[2], "regression_loss": out4[3]}

    # ---- plan / graph machinery ------------------------------------------------------------------
    def _make_plan(self, key) -> dict:
        B, N, Tt, Ks, off, n_clip, no_mask = key
        dev, cfg = self.device, self.cfg
        S, Sc = self.image_size, cfg.clip.image_size
        static = {
            "images": torch.empty((B, 3, S, S), dtype=BF16, device=dev),
            "images_clip": torch.empty((n_clip, 3, Sc, Sc), dtype=BF16, device=dev),
            "input_ids": torch.empty((N, Tt), dtype=torch.int64, device=dev),
            "mask": None if no_mask else torch.empty((N, Tt), dtype=torch.uint8, device=dev),
            "segs": torch.empty((sum(Ks), 256, 256), dtype=BF16, device=dev),
        }
        # CLIP input per conversation: image i for conversations offset[i]..offset[i+1] (LISA.py:272,293-303)
        if n_clip == 1 and B == 1:
            conv_image = [0] * N
        else:
            assert n_clip == B, "images_clip must hold one image per SAM image"
            conv_image = [i for i in range(B) for _ in range(off[i + 1] - off[i])]
        conv_index = None if conv_image == list(range(N)) and n_clip == N else torch.tensor(conv_image, device=dev)
        first_conv = None if N == B else torch.tensor(off[:-1], device=dev, dtype=torch.long)
        return {"key": key, "static": static, "conv_index": conv_index, "first_conv": first_conv,
                "sel": self.selector.make_plan(Ks), "graphs": None, "outputs": None, "launches": 0}

    def _pipeline(self, plan, ev_img, ev_segs, graphs):
        """The forward as three stages — text branch, image branch, selector — each either a captured CUDA graph
        (`graphs`) or eager launches (graph-capturable: no host syncs, every index tensor comes from the plan).
        The image branch (1) and the text branch (2, 3) are independent until the selector; they run on two
        streams so that each one's launch gaps and partial last waves are filled by the other's CTAs.
        ev_img / ev_segs: events after the image / proposal copies on the copy stream (None: copies were issued on
        the current stream).  LLMSEG_OVERLAP=0 serialises everything on the current stream."""
        st = plan["static"]
        cur = torch.cuda.current_stream()

        def text():
            if graphs is None:
                return self._text_branch(plan)
            graphs["text"].replay()
            return plan["text_embed"]

        def image():
            if graphs is None:
                return self.image_encoder.forward(st["images"])
            graphs["image"].replay()
            return plan["emb_tokens"]

        if self.overlap_branches:
            side = self._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                text_embed = text()
            if ev_img is not None:
                cur.wait_event(ev_img)
            emb_tokens = image()
            cur.wait_stream(side)
        else:
            emb_tokens = image()
            text_embed = text()
        if ev_segs is not None:
            cur.wait_event(ev_segs)
        # 4. selector
        if graphs is None:
            return self.selector.forward(emb_tokens, st["segs"], text_embed, plan["sel"])
        graphs["sel"].replay()
        return plan["outputs"]

    def _text_branch(self, plan) -> Tensor:
        """CLIP tower + projector -> splice -> LLaMA -> text_hidden_fcs on the [SEG] rows: [B,256]."""
        st = plan["static"]
        N, Tt = plan["key"][1], plan["key"][2]
        # 2. CLIP tower + projector (once per distinct image; the reference recomputes per conversation)
        feats = self.clip.forward(st["images_clip"])
        if plan["conv_index"] is not None:
            feats = feats.index_select(0, plan["conv_index"]).contiguous()
        # 3. splice + LLaMA; only the hidden state that predicts [SEG] is normalised and returned
        embeds, kv_len, seg_row = ops.embed_splice(st["input_ids"], st["mask"], self.llama.embed, feats,
                                                   image_token=IMAGE_TOKEN_INDEX, seg_token=self.seg_token_idx)
        T = Tt + feats.shape[1] - 1
        rows = seg_row if plan["first_conv"] is None else seg_row.index_select(0, plan["first_conv"]).contiguous()
        hidden = self.llama.forward(embeds, N, T, kv_len, out_rows=rows)      # conversation 0 of each image (LISA.py:400)
        return self.selector.text_embed(hidden)                               # [B,256]

    def _capture(self, plan) -> None:
        """Warm-up (populates scratch buffers, per-stream GEMM workspaces, index maps, function attributes), then
        capture the three stages.  The text graph is captured on the stream it will share SMs from (its GEMM
        workspace is keyed by stream, and it replays concurrently with the image graph: separate capture streams,
        separate memory pools)."""
        from . import _lib
        torch.cuda.synchronize(self.device)          # the staged inputs have landed
        self._pipeline(plan, None, None, None)
        torch.cuda.synchronize(self.device)
        st = plan["static"]
        for stream in (self._side_stream, self._cap_stream):
            with torch.cuda.stream(stream):
                ops.ensure_workspace(self.device)
        torch.cuda.synchronize(self.device)
        graphs = {n: torch.cuda.CUDAGraph() for n in ("text", "image", "sel")}
        n0 = _lib.launch_count()
        with torch.cuda.graph(graphs["text"], stream=self._side_stream):
            plan["text_embed"] = self._text_branch(plan)
        with torch.cuda.graph(graphs["image"], stream=self._cap_stream):
            plan["emb_tokens"] = self.image_encoder.forward(st["images"])
        with torch.cuda.graph(graphs["sel"], stream=self._cap_stream):
            plan["outputs"] = self.selector.forward(plan["emb_tokens"], st["segs"], plan["text_embed"], plan["sel"])
        plan["launches"] = _lib.launch_count() - n0
        torch.cuda.synchronize(self.device)
        plan["graphs"] = graphs
