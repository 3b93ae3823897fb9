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
from collections import OrderedDict
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


def strip_peft_prefix(sd: Dict[str, Tensor], lora_alpha: float = 16.0) -> Dict[str, Tensor]:
    """Accept checkpoints saved through PEFT (`base_model.model.` prefix, reference training.py:194-237)
    and merge LoRA deltas W + (lora_alpha/r)·B·A into their base weights (reference defaults --lora_r 8,
    --lora_alpha 16, training.py:45,90; pass the value the checkpoint was trained with).  Adapter keys are
    `<module>.lora_A[.<adapter>].weight` next to `<module>.weight` (PEFT 0.4 layout); a LoRA pair without its
    base weight is an error, not a silent no-op."""
    out = {}
    for k, v in sd.items():
        if k.startswith("base_model.model."):
            k = k[len("base_model.model."):]
        out[k] = v
    lora_a = {k: v for k, v in out.items() if ".lora_A." in k}
    for ka, a in lora_a.items():
        kb = ka.replace(".lora_A.", ".lora_B.")
        base = ka.split(".lora_A.")[0] + ".weight"
        if kb not in out:
            raise KeyError(f"LoRA checkpoint: {ka} has no matching {kb}")
        if base not in out:
            raise KeyError(f"LoRA checkpoint: adapter {ka} has no base weight {base}")
        r = a.shape[0]
        out[base] = (out[base].float() + (float(lora_alpha) / r) * (out[kb].float() @ a.float())).to(out[base].dtype)
    return {k: v for k, v in out.items() if ".lora_" not in k}


def select_proposals(output_dict: dict, threshold: Optional[float] = None):
    """The two selection rules the reference's validation loops apply to a forward's outputs:
    `validate` keeps the proposal with the highest similarity (training.py:627-629) — here the argmax fused into
    the select kernel (`best_index`, falling back to argmax of `pred_similarity` for a plain reference dict) — and
    `validate_threshold` keeps every proposal whose predicted IoU exceeds `threshold` (training.py:712-718).
    Returns one `(best, kept)` pair per image; `kept` is None without a threshold."""
    sims, ious = output_dict["pred_similarity"], output_dict["pred_iou"]
    best = output_dict.get("best_index")
    best = [int(v) for v in best.tolist()] if best is not None else [int(torch.argmax(s_[0])) for s_ in sims]
    kept = None
    if threshold is not None:
        kept = [[i for i, v in enumerate(u[0].float().tolist()) if v > threshold] for u in ious]
    return [(best[i], None if kept is None else kept[i]) for i in range(len(sims))]


def bucket_tokens(t_text: int, quantum: int = 32) -> int:
    """Prompt lengths are padded up to a multiple of `quantum` tokens (the pad is masked out: right padding is
    inert under the causal + key-padding mask), so one captured graph serves a whole length bucket."""
    return -(-int(t_text) // quantum) * quantum


def bucket_props(k: int) -> int:
    """Proposal-count capacity of a selector plan: 32 / 64 / 128 (the selector kernels' maximum)."""
    for cap in (32, 64, 128):
        if k <= cap:
            return cap
    raise ValueError(f"selector kernels support at most 128 proposals per image, got {k}")


class _LRU:
    """Small LRU of per-shape stage plans (static input buffers + captured CUDA graph + its private memory pool).
    Evicting a plan drops the last references to all three."""

    def __init__(self, cap: int):
        self.cap, self.d = cap, OrderedDict()

    def get(self, key):
        v = self.d.get(key)
        if v is not None:
            self.d.move_to_end(key)
        return v

    def put(self, key, val):
        self.d[key] = val
        self.d.move_to_end(key)
        while len(self.d) > self.cap:
            self.d.popitem(last=False)

    def __len__(self):
        return len(self.d)


class LisaEngine:
    """The inference / training-value forward on the sm_100a kernels, built from a reference-named state dict.
    `LISAForCausalLM` below is the reference-shaped `nn.Module` front end of this class."""

    def __init__(self, state_dict: Dict[str, Tensor], cfg: Optional[LisaCfg] = None, *,
                 device: str = "cuda:0", seg_token_idx: Optional[int] = None, max_seq: int = 1024,
                 use_cuda_graph: bool = True, ce_loss_weight: float = 1.0, align_loss_weight: float = 1.0,
                 regression_loss_weight: float = 1.0, lora_alpha: float = 16.0, max_plans: int = 4,
                 graph_after_uses: int = 1, bucket_shapes: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("llmseg_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        from . import _lib
        _lib.lib()  # fail loudly now if the CUDA library is missing
        self.cfg = cfg or LisaCfg()
        if seg_token_idx is not None:
            self.cfg.seg_token_idx = seg_token_idx
        self.seg_token_idx = self.cfg.seg_token_idx
        self.device = torch.device(device)
        sd = strip_peft_prefix(state_dict, lora_alpha)
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
        # SAM-Everything proposal generation: built when the checkpoint carries SAM's decoder
        self.proposals = None
        if "model.visual_model.mask_decoder.iou_token.weight" in sd and self.cfg.image_encoder == "sam":
            from .proposals import SamProposalGenerator
            self.proposals = SamProposalGenerator(sd, self.device)
        self.use_cuda_graph = use_cuda_graph
        self.overlap_branches = os.environ.get("LLMSEG_OVERLAP", "1") != "0"
        self._side_stream = torch.cuda.Stream(device=self.device)    # text branch
        self._copy_stream = torch.cuda.Stream(device=self.device)    # image / proposal staging copies
        self._cap_stream = torch.cuda.Stream(device=self.device)     # capture stream of the image / selector graphs
        # Stage plans, keyed by BUCKETED shapes and evicted LRU: real validation data varies the prompt length
        # and the proposal count per image (K <= 50, reference utils/sam_mask_reader.py:78-83), and an exact-shape
        # key would capture on almost every call and pin a new set of buffers each time.
        #   image   (B)                                  static images            -> emb_tokens
        #   text    (N, T_text bucket of 32, n_clip, .)  static CLIP images / ids / mask -> text_all, seg_row
        #   sel     (B, N, K capacity 32/64/128, .)      static proposals         -> sim, iou, best
        # A plan runs eagerly on its first `graph_after_uses` uses (one-off shapes never pay for a capture, and the
        # eager pass is the warm-up a capture needs) and is captured into a CUDA graph on the next one.
        # bucket_shapes=False keys the plans on the exact (T_text, max K) instead (tests: the un-padded path).
        self.max_plans, self.graph_after_uses, self.bucket_shapes = max_plans, graph_after_uses, bucket_shapes
        # text branch -> selector interface: fixed buffers per conversation count, so a selector graph captured
        # against one text plan stays valid for every other prompt-length bucket
        self._text_iface: Dict[int, tuple] = {}
        self._plans = {"image": _LRU(max_plans), "text": _LRU(max_plans), "sel": _LRU(max_plans)}
        self.graphs_captured = 0
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
    def generate_proposals(self, images: Tensor, **kwargs) -> List[dict]:
        """SAM-Everything on the kernels: images [B,3,1024,1024] -> one record per image whose "segs"
        ([K,256,256] bf16 soft masks, largest first) is what `model_forward` takes as `sam_segs_list[i]` — the step the
        reference pre-computes offline with SamAutomaticMaskGenerator (prepare_datasets/*.py) and then re-reads and
        resizes per sample (utils/sam_mask_reader.py:69-113, utils/dataset.py:620-622).  Needs the SAM prompt-encoder /
        mask-decoder weights in the state dict (`model.visual_model.{prompt_encoder,mask_decoder}.*`).
        kwargs: see proposals.SamProposalGenerator.generate."""
        if self.cfg.image_encoder != "sam":
            raise RuntimeError("proposal generation runs on the SAM ViT-H features (image_encoder='sam')")
        if self.proposals is None:
            raise RuntimeError("the state dict this model was built from holds no SAM prompt-encoder / mask-decoder weights")
        tok = self.sam.forward(images.to(self.device, BF16))
        return [self.proposals.generate(tok[i].contiguous(), **kwargs) for i in range(tok.shape[0])]

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
        B, N, Tt = images.shape[0], input_ids.shape[0], input_ids.shape[1]
        if offset is None:
            offset = torch.arange(B + 1)
        if B != len(offset) - 1:
            raise ValueError(f"batch_size == len(offset) - 1 (reference LISA.py:250): {B} images, offset of {len(offset)}")
        off = tuple(int(v) for v in offset.tolist())
        if off[0] != 0 or off[-1] != N or any(off[i + 1] <= off[i] for i in range(B)):
            raise ValueError(f"offset {off} must start at 0, end at the number of conversations ({N}) and give every "
                             f"image at least one conversation")
        if sam_segs_list is None or len(sam_segs_list) != B:
            raise ValueError("sam_segs_list must hold one [K_i,256,256] proposal tensor per image")
        Ks = tuple(int(s.shape[0]) for s in sam_segs_list)
        if min(Ks) < 1:
            raise ValueError(f"every image needs at least one mask proposal, got K = {Ks}")
        n_clip = images_clip.shape[0]
        if not (n_clip == B or (n_clip == 1 and B == 1)):
            raise ValueError("images_clip must hold one image per SAM image")
        n_img = self.cfg.clip.tokens - 1
        if Tt + n_img - 1 > self.llama.max_seq:
            raise ValueError(f"sequence length {Tt + n_img - 1} exceeds the RoPE table ({self.llama.max_seq})")
        if not input_ids.is_cuda:
            # Free on host inputs: a conversation without [SEG] has no state to score the proposals with (the
            # reference returns an empty [0,K] similarity or fails inside its attention); more than one is outside
            # the inference contract (LISA.py:394 "during inference, C = 1").  Device-resident ids are not read back
            # (that would be a sync per call): the select kernel marks such rows NaN / best_index -1 instead.
            cnt = (input_ids[:, 1:] == self.seg_token_idx).sum(dim=1)
            if bool((cnt != 1).any()):
                raise ValueError(f"every conversation must contain exactly one [SEG] token after position 0; "
                                 f"counts per conversation: {cnt.tolist()}")
        Tb = bucket_tokens(Tt) if self.bucket_shapes else Tt
        if Tb + n_img - 1 > self.llama.max_seq:
            Tb = Tt
        k_cap = bucket_props(max(Ks))
        if not self.bucket_shapes:
            k_cap = max(Ks)
        identity = N == B
        ip = self._stage_plan("image", (B,))
        tp = self._stage_plan("text", (N, Tb, n_clip, identity))
        sp = self._stage_plan("sel", (B, N, k_cap, identity))
        # ---- stage the inputs into the plans' static device buffers (H2D or D2D copies) and run.  The small
        # inputs (CLIP image, ids, mask: all the text branch needs) go first on the current stream; the 1024 px
        # images and the proposals follow on a copy stream in the order they are needed, so the text branch starts
        # under the image copy and the proposals (67 MB at batch 8, read only by the selector) arrive under the
        # encoders instead of in front of them.
        cur = torch.cuda.current_stream()
        tp["images_clip"].copy_(images_clip, non_blocking=True)
        tp["input_ids"][:, :Tt].copy_(input_ids, non_blocking=True)
        if attention_masks is not None:
            tp["mask"][:, :Tt].copy_(attention_masks, non_blocking=True)
        elif tp["mask_state"] != ("ones", Tt):
            tp["mask"][:, :Tt].fill_(1)
        if Tt < Tb and (tp["mask_state"] is None or tp["mask_state"][1] != Tt):
            tp["mask"][:, Tt:].zero_()            # bucket padding: masked out (right padding is inert), pad token 0
            tp["input_ids"][:, Tt:].zero_()
        tp["mask_state"] = ("ones" if attention_masks is None else "given", Tt)
        if not identity and tp["off"] != off:
            conv_image = [i for i in range(B) for _ in range(off[i + 1] - off[i])]
            tp["conv_index"].copy_(torch.tensor([0] * N if n_clip == 1 else conv_image, dtype=torch.int64))
            tp["off"] = off
        if not identity and sp["off"] != off:
            conv_image = [i for i in range(B) for _ in range(off[i + 1] - off[i])]
            sp["first_conv"].copy_(torch.tensor(off[:-1], dtype=torch.int64))
            sp["conv_group"].copy_(torch.tensor(conv_image, dtype=torch.int32))
            sp["off"] = off
        self.selector.update_plan(sp["sel"], Ks)
        cp = self._copy_stream if self.overlap_branches else cur
        if cp is not cur:
            cp.wait_stream(cur)      # the previous forward (enqueued on cur) has finished reading the static buffers
        with torch.cuda.stream(cp):
            ip["images"].copy_(images, non_blocking=True)
            ev_img = cp.record_event() if cp is not cur else None
            r0 = 0
            for sgs, kk in zip(sam_segs_list, Ks):
                sp["segs"][r0:r0 + kk].copy_(sgs, non_blocking=True)
                r0 += kk
            ev_segs = cp.record_event() if cp is not cur else None
        # ---- CUDA-graph replay (captured on the plan's second use) or eager launches
        plans = (ip, tp, sp)
        use_graph = self.use_cuda_graph and all(p["uses"] >= self.graph_after_uses for p in plans)
        if use_graph and any(p["graph"] is None for p in plans):
            self._capture(ip, tp, sp)
        from . import _lib
        n0 = _lib.launch_count()
        sim, iou, best = self._pipeline(ip, tp, sp, ev_img, ev_segs, use_graph)
        if use_graph:
            self.last_forward_launches = ip["launches"] + tp["launches"] + sp["launches"]
        else:
            self.last_forward_launches = _lib.launch_count() - n0
        for p in plans:
            p["uses"] += 1
        # pred_similarity[i]: one row per conversation of image i ([C_i, K_i], reference LISA.py:397-403);
        # pred_iou[i]: [1, K_i] from the image's first conversation (LISA.py:405-408)
        pred_similarity = [sim[off[i]:off[i + 1], :Ks[i]].to(BF16) for i in range(B)]
        pred_iou = [iou[i:i + 1, :Ks[i]].to(BF16) for i in range(B)]
        sim_first = sim.clone() if identity else sim.index_select(0, sp["first_conv"])
        return {"pred_similarity": pred_similarity, "gt_masks": masks_list, "pred_iou": pred_iou,
                "best_index": best.clone(), "similarity_padded": sim_first, "iou_padded": iou.clone(),
                "similarity_all": sim.clone() if not identity else sim_first}

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
        if B != len(offset) - 1:
            raise ValueError(f"batch_size == len(offset) - 1 (reference LISA.py:250): {B} images, offset of {len(offset)}")
        off = [int(v) for v in offset.tolist()]
        if off[-1] != N or len(sam_segs_list) != B or images_clip.shape[0] != B:
            raise ValueError("offset must end at the number of conversations; sam_segs_list / images_clip hold one entry per image")
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
    def _stage_plan(self, stage: str, key: tuple) -> dict:
        plan = self._plans[stage].get(key)
        if plan is not None:
            return plan
        dev, cfg = self.device, self.cfg
        plan = {"key": key, "graph": None, "out": None, "uses": 0, "launches": 0, "off": None}
        if stage == "image":
            (B,) = key
            S = self.image_size
            plan["images"] = torch.empty((B, 3, S, S), dtype=BF16, device=dev)
            plan["private"] = {}     # per-plan encoder buffers (SAM k / vT with pre-written padding rows)
        elif stage == "text":
            N, Tb, n_clip, identity = key
            Sc = cfg.clip.image_size
            plan["images_clip"] = torch.empty((n_clip, 3, Sc, Sc), dtype=BF16, device=dev)
            plan["input_ids"] = torch.zeros((N, Tb), dtype=torch.int64, device=dev)
            plan["mask"] = torch.zeros((N, Tb), dtype=torch.uint8, device=dev)
            plan["mask_state"] = None
            plan["conv_index"] = None if identity else torch.zeros(N, dtype=torch.int64, device=dev)
        else:
            B, N, k_cap, identity = key
            plan["segs"] = torch.zeros((B * k_cap, 256, 256), dtype=BF16, device=dev)
            plan["sel"] = self.selector.make_plan([1] * B, k_cap=k_cap)
            plan["first_conv"] = None if identity else torch.zeros(B, dtype=torch.int64, device=dev)
            plan["conv_group"] = None if identity else torch.zeros(N, dtype=torch.int32, device=dev)
        self._plans[stage].put(key, plan)
        return plan

    def _pipeline(self, ip, tp, sp, ev_img, ev_segs, use_graph: bool):
        """The forward as three stages — text branch, image branch, selector — each either a captured CUDA graph
        or eager launches (graph-capturable: no host syncs, every index tensor lives in a plan).
        The image branch and the text branch are independent until the selector; they run on two
        streams so that each one's launch gaps and partial last waves are filled by the other's CTAs.
        ev_img / ev_segs: events after the image / proposal copies on the copy stream (None: copies were issued on
        the current stream).  LLMSEG_OVERLAP=0 serialises everything on the current stream."""
        cur = torch.cuda.current_stream()

        def text():
            if use_graph:
                tp["graph"].replay()
                return tp["out"]
            return self._text_branch(tp)

        def image():
            if use_graph:
                ip["graph"].replay()
                return ip["out"]
            return self.image_encoder.forward(ip["images"], private=ip["private"])

        if self.overlap_branches:
            side = self._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                text_out = text()
            if ev_img is not None:
                cur.wait_event(ev_img)
            emb_tokens = image()
            cur.wait_stream(side)
        else:
            emb_tokens = image()
            text_out = text()
        if ev_segs is not None:
            cur.wait_event(ev_segs)
        if use_graph:
            sp["graph"].replay()
            return sp["out"]
        return self._selector_stage(sp, emb_tokens, text_out)

    def _text_branch(self, tp):
        """CLIP tower + projector -> splice -> LLaMA -> text_hidden_fcs on the [SEG] rows:
        (text_all [N,256], seg_row int32 [N])."""
        N, Tb = tp["key"][0], tp["key"][1]
        # CLIP tower + projector (once per distinct image; the reference recomputes per conversation)
        feats = self.clip.forward(tp["images_clip"])
        if tp["conv_index"] is not None:
            feats = feats.index_select(0, tp["conv_index"]).contiguous()
        # splice + LLaMA; only the hidden state that predicts [SEG] is normalised and returned
        text_buf, seg_buf = self._iface(N)
        embeds, kv_len, seg_row = ops.embed_splice(tp["input_ids"], tp["mask"], self.llama.embed, feats,
                                                   image_token=IMAGE_TOKEN_INDEX, seg_token=self.seg_token_idx,
                                                   seg_row_out=seg_buf)
        T = Tb + feats.shape[1] - 1
        hidden = self.llama.forward(embeds, N, T, kv_len, out_rows=seg_row)
        return self.selector.text_embed(hidden, out=text_buf), seg_row

    def _iface(self, N: int):
        buf = self._text_iface.get(N)
        if buf is None:
            buf = (torch.zeros((N, 256), dtype=BF16, device=self.device),
                   torch.zeros(N, dtype=torch.int32, device=self.device))
            self._text_iface[N] = buf
        return buf

    def _selector_stage(self, sp, emb_tokens, text_out):
        text_all, seg_row = text_out
        if sp["first_conv"] is None:
            return self.selector.forward(emb_tokens, sp["segs"], text_all, sp["sel"], conv_valid=seg_row)
        text_first = text_all.index_select(0, sp["first_conv"]).contiguous()   # LISA.py:364,376: keys = round 0's text
        return self.selector.forward(emb_tokens, sp["segs"], text_first, sp["sel"], text_all=text_all,
                                     conv_group=sp["conv_group"], conv_valid=seg_row)

    def _capture(self, ip, tp, sp) -> None:
        """Capture the stages that have no graph yet.  Every plan has run eagerly at least once by now (scratch
        buffers, index maps and function attributes exist).  The text graph is captured on the stream it will share
        SMs from (its GEMM workspace is keyed by stream, and it replays concurrently with the image graph: separate
        capture streams, separate memory pools)."""
        from . import _lib
        torch.cuda.synchronize(self.device)          # the staged inputs have landed
        for stream in (self._side_stream, self._cap_stream):
            with torch.cuda.stream(stream):
                ops.ensure_workspace(self.device)
        torch.cuda.synchronize(self.device)

        def cap(plan, stream, fn):
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, stream=stream):
                plan["out"] = fn()
            plan["launches"] = _lib.launch_count() - n0
            plan["graph"] = g
            self.graphs_captured += 1

        if tp["graph"] is None:
            cap(tp, self._side_stream, lambda: self._text_branch(tp))
        if ip["graph"] is None:
            cap(ip, self._cap_stream, lambda: self.image_encoder.forward(ip["images"], private=ip["private"]))
        if sp["graph"] is None or sp.get("graph_inputs") != ip["out"].data_ptr():
            # the selector graph bakes in the address of the image features it was captured against (one image
            # plan per batch size: this only changes when that plan was evicted and rebuilt) and reads the text
            # branch through the fixed per-N interface buffers
            cap(sp, self._cap_stream, lambda: self._selector_stage(sp, ip["out"], tp["out"]))
            sp["graph_inputs"] = ip["out"].data_ptr()
        torch.cuda.synchronize(self.device)


# ==============================================================================================
# Reference-shaped front end
# ==============================================================================================
def cfg_from_hf_config(config, **kwargs) -> LisaCfg:
    """LisaCfg from the LLaVA/LLaMA `config` object (or dict) the reference constructs its model with
    (reference model/LISA.py:144-170; HF LlamaConfig field names), plus the reference's kwargs."""
    get = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
    cfg = LisaCfg()
    ll = cfg.llama
    ll.hidden = int(get("hidden_size", ll.hidden))
    ll.layers = int(get("num_hidden_layers", ll.layers))
    ll.heads = int(get("num_attention_heads", ll.heads))
    ll.mlp = int(get("intermediate_size", ll.mlp))
    ll.vocab = int(get("vocab_size", ll.vocab))
    ll.eps = float(get("rms_norm_eps", ll.eps))
    ll.rope_theta = float(get("rope_theta", ll.rope_theta) or ll.rope_theta)
    if get("num_key_value_heads") not in (None, ll.heads):
        raise ValueError("grouped-query attention is not on the LLM-Seg path (LLaVA-7B: 32 query = 32 kv heads)")
    cfg.clip.select_layer = int(get("mm_vision_select_layer", cfg.clip.select_layer))
    cfg.out_dim = int(kwargs.get("out_dim", get("out_dim", cfg.out_dim)))
    if cfg.out_dim != 256:
        raise ValueError("the selector kernels are laid out for out_dim = 256 (reference training.py:52 default)")
    if "seg_token_idx" in kwargs:
        cfg.seg_token_idx = int(kwargs["seg_token_idx"])
    cfg.image_encoder = kwargs.get("image_encoder", cfg.image_encoder)
    return cfg


class _Keys(tuple):
    """`load_state_dict` result, shaped like torch's _IncompatibleKeys."""
    missing_keys = property(lambda s: s[0])
    unexpected_keys = property(lambda s: s[1])


class LISAForCausalLM(torch.nn.Module):
    """Drop-in for the reference class of the same name (model/LISA.py:144-170):

        model = LISAForCausalLM(config, seg_token_idx=..., train_mask_decoder=..., out_dim=256,
                                vision_pretrained=..., vision_tower=..., use_mm_start_end=True,
                                ce_loss_weight=..., align_loss_weight=..., regression_loss_weight=...)
        model.load_state_dict(reference_named_state_dict)       # or LISAForCausalLM.from_pretrained(dir, ...)
        model.eval(); out = model(**input_dict)                 # forward(**kwargs) -> model_forward(...)

    It is an `nn.Module` (`.eval()`, `.train()`, `.state_dict()`, `.to()`, hooks) whose compute runs in `LisaEngine`:
    the weights are re-laid-out for the kernels when they are loaded, so the module owns no nn.Parameters
    (`parameters()` is empty — there is no autograd path here); `state_dict()` returns the reference-named tensors
    that were loaded.  Extra kwargs: `device`, `image_encoder` ("sam" | "dinov2"), `lora_alpha`, `max_seq`,
    `use_cuda_graph`, `max_plans`.

    The round-1 signature `LISAForCausalLM(state_dict, cfg=None, device=..., ...)` is kept: a mapping as the first
    argument builds the engine immediately."""

    _ENGINE_KW = ("device", "max_seq", "use_cuda_graph", "lora_alpha", "max_plans", "graph_after_uses", "bucket_shapes",
                  "ce_loss_weight", "align_loss_weight", "regression_loss_weight")

    def __init__(self, config=None, cfg: Optional[LisaCfg] = None, **kwargs):
        super().__init__()
        self._engine: Optional[LisaEngine] = None
        self._ref_sd: Optional[Dict[str, Tensor]] = None
        self._ekw = {k: kwargs[k] for k in self._ENGINE_KW if k in kwargs}
        if isinstance(config, dict) and config and all(torch.is_tensor(v) for v in config.values()):
            # legacy: (state_dict, LisaCfg)
            self.cfg = cfg or LisaCfg()
            if kwargs.get("seg_token_idx") is not None:
                self.cfg.seg_token_idx = int(kwargs["seg_token_idx"])
            self.config = None
            self._build(config, keep=False)
            return
        # reference signature: (config, **kwargs); kwargs the kernels have no use for are accepted and ignored
        # (train_mask_decoder, vision_pretrained, vision_tower, use_mm_start_end: LISA.py:153-156,18-33)
        self.config = config
        self.cfg = cfg if cfg is not None else cfg_from_hf_config(config if config is not None else {}, **kwargs)
        if cfg is not None and kwargs.get("seg_token_idx") is not None:
            self.cfg.seg_token_idx = int(kwargs["seg_token_idx"])

    # ---- weights ---------------------------------------------------------------------------------
    def _build(self, sd: Dict[str, Tensor], keep: bool) -> None:
        self._engine = LisaEngine(sd, self.cfg, **self._ekw)
        self._ref_sd = dict(sd) if keep else None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Build the kernel-side weights from a reference-named state dict (PEFT prefixes / LoRA pairs accepted,
        see strip_peft_prefix).  Keys the forward path never reads (SAM prompt encoder / mask decoder, lm_head at
        inference, DINOv2 when image_encoder='sam', ...) are reported as unexpected, never an error."""
        before = set(state_dict.keys())
        try:
            self._build(state_dict, keep=True)
        except KeyError as e:
            if strict:
                raise RuntimeError(f"Error(s) in loading state_dict for LISAForCausalLM: missing key {e}") from e
            raise
        return _Keys(([], sorted(before - set(self._engine_keys()))))

    def _engine_keys(self):
        e, p = self._engine, []
        enc = "model.visual_model.image_encoder." if e.cfg.image_encoder == "sam" else "model.visual_model_dinov2."
        keep = ("model.vision_tower.", "model.mm_projector.", "model.embed_tokens.", "model.layers.", "model.norm.",
                "model.text_hidden_fcs.", "model.lisa_", "lm_head.", enc)
        if e.proposals is not None:
            keep += ("model.visual_model.prompt_encoder.", "model.visual_model.mask_decoder.")
        for k in (self._ref_sd or {}):
            kk = k[len("base_model.model."):] if k.startswith("base_model.model.") else k
            if kk.startswith(keep):
                p.append(k)
        return p

    def state_dict(self, *args, **kwargs):
        if self._ref_sd is None:
            raise RuntimeError("this model was built without keeping its reference-named state dict "
                               "(legacy constructor); load it through load_state_dict / from_pretrained")
        return dict(self._ref_sd)

    @classmethod
    def from_pretrained(cls, path, *model_args, config=None, **kwargs):
        """Directory with `config.json` + `pytorch_model*.bin` / `*.pt` / `*.safetensors` shards (the layout the
        reference loads with HF `from_pretrained`, training.py:157-159), or a single checkpoint file.
        `torch_dtype`, `low_cpu_mem_usage` and friends are accepted and ignored (the engine is bf16)."""
        import glob
        import json as _json
        for k in ("torch_dtype", "low_cpu_mem_usage", "device_map", "quantization_config", "load_in_8bit", "load_in_4bit"):
            kwargs.pop(k, None)
        files = [path] if os.path.isfile(path) else sorted(
            glob.glob(os.path.join(path, "pytorch_model*.bin")) + glob.glob(os.path.join(path, "*.pt")) +
            glob.glob(os.path.join(path, "*.safetensors")))
        if not files:
            raise FileNotFoundError(f"no checkpoint shards (*.bin, *.pt, *.safetensors) under {path}")
        if config is None and os.path.isdir(path) and os.path.exists(os.path.join(path, "config.json")):
            with open(os.path.join(path, "config.json")) as f:
                config = _json.load(f)
        sd: Dict[str, Tensor] = {}
        for fn in files:
            if fn.endswith(".safetensors"):
                from safetensors.torch import load_file
                sd.update(load_file(fn))
            else:
                part = torch.load(fn, map_location="cpu", weights_only=True)
                sd.update(part.get("module", part) if isinstance(part, dict) else part)   # DeepSpeed wraps in "module"
        model = cls(config if config is not None else {}, **kwargs)
        model.load_state_dict(sd)
        return model

    # ---- nn.Module surface that makes no sense for pre-laid-out bf16 weights: accepted, no-ops --------
    def to(self, *args, **kwargs):
        return self

    def cuda(self, device=None):
        return self

    def bfloat16(self):
        return self

    def half(self):
        raise TypeError("llmseg_b200 computes in bf16 (the reference's --precision bf16 default, training.py:37-42)")

    def get_model(self):
        return self

    def resize_token_embeddings(self, n: int):
        if self._engine is not None and int(n) != self._engine.llama.embed.shape[0]:
            raise NotImplementedError("load a checkpoint whose embedding table already holds the added tokens "
                                      "([SEG], <im_start>, <im_end>: reference training.py:121-137,229)")
        self.cfg.llama.vocab = int(n)

    # ---- forward -----------------------------------------------------------------------------------
    @property
    def engine(self) -> LisaEngine:
        if self._engine is None:
            raise RuntimeError("no weights loaded: call load_state_dict(...) or use from_pretrained(...)")
        return self._engine

    def forward(self, **kwargs):
        return self.engine.forward(**kwargs)

    def model_forward(self, *args, **kwargs):
        return self.engine.model_forward(*args, **kwargs)

    def __getattr__(self, name):
        # everything else (get_visual_embs, seg_token_idx, clip / llama / selector sub-engines, use_cuda_graph, ...)
        try:
            return super().__getattr__(name)
        except AttributeError:
            eng = self.__dict__.get("_engine")
            if eng is not None and hasattr(eng, name):
                return getattr(eng, name)
            raise

    def __setattr__(self, name, value):
        eng = self.__dict__.get("_engine")
        if eng is not None and name in ("use_cuda_graph", "overlap_branches", "graph_after_uses", "ce_loss_weight",
                                        "align_loss_weight", "regression_loss_weight"):
            setattr(eng, name, value)
            return
        super().__setattr__(name, value)
