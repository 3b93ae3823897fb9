"""SAM-Everything proposal generation on the sm_100a kernels: from the SAM ViT-H features of an image to the
K soft mask proposals `[K,256,256]` that `model_forward` takes as `sam_segs_list` — the offline step the reference runs
with `SamAutomaticMaskGenerator` (reference model/segment_anything/automatic_mask_generator.py:141-322, called from
prepare_datasets/*.py) followed by LLM-Seg's own top-50 / resize (utils/sam_mask_reader.py:69-83, utils/dataset.py:620-622).

Per image, with the class defaults of the reference (32 x 32 point grid, one crop layer, no small-region clean-up):

  1. prompt encoder: one foreground point + the padding point per prompt -> 7 decoder tokens      (llmseg_point_tokens)
  2. mask decoder: two-way transformer (2 blocks + final attention), 4x up-scaling, hyper-network product, IoU head
     — every linear and the first ConvTranspose on llmseg_gemm, the 7-token attentions on llmseg_small_attention /
     llmseg_tok2img_attention / llmseg_img2tok_attention, the rest of the up-scaling in llmseg_upscale_logits
                                                                                                   -> low-res logits [P,3,256,256]
  3. per candidate, on the 4x up-sampled logits evaluated on the fly: area, stability counts, box  (llmseg_mask_stats)
  4. predicted-IoU / stability filters, score sort (host, <= 3072 records), box NMS                (llmseg_box_nms)
  5. the `top_k` largest survivors -> antialiased 256 x 256 soft masks                            (llmseg_mask_soft)

Exact shortcuts (same math, fewer bytes):
  * `k_proj(keys + pe) = k_proj(keys) + k_proj(pe)`: the positional term is a constant of the weights, added in the GEMM
    epilogue as a broadcast residual — `keys + pe` is never materialised (three times per prompt in the reference)
  * in block 0 every prompt sees the same image keys: their K / V / Q projections are computed once per image, not per prompt
  * the three projections that read a block's image keys (token->image K and V, image->token Q) are one GEMM
  * ConvTranspose2d(k=2, s=2) is a GEMM whose output row holds the 2 x 2 sub-pixels; LayerNorm2d / GELU / the second
    ConvTranspose / the hyper-network product work on that layout directly, so no pixel shuffle is ever executed
  * the 1024 x 1024 masks are never written: statistics, boxes and the soft proposals are computed from the 256 x 256 logits
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

import os

from . import ops
from .encoders import BF16, _dev

# LLMSEG_AMG_FUSED_UPSCALE=0: LayerNorm2d+GELU, ConvTranspose #2 and the hyper-network product as three launches
# (ln64_gelu + gemm + mask_logits) instead of `upscale_logits` — A/B runs and the parity test of the fused kernel
FUSED_UPSCALE = os.environ.get("LLMSEG_AMG_FUSED_UPSCALE", "1") != "0"

Tensor = torch.Tensor


def point_grid(n_per_side: int, size: int = 1024) -> np.ndarray:
    """`build_point_grid` scaled to the image (reference utils/amg.py:179-186, automatic_mask_generator.py:240-241)."""
    off = 1 / (2 * n_per_side)
    one = np.linspace(off, 1 - off, n_per_side)
    x = np.tile(one[None, :], (n_per_side, 1))
    y = np.tile(one[:, None], (1, n_per_side))
    return (np.stack([x, y], axis=-1).reshape(-1, 2) * size).astype(np.float32)


class SamProposalGenerator:
    """Prompt encoder + mask decoder + automatic mask generator of SAM for already-square 1024 x 1024 inputs."""

    def __init__(self, sd: Dict[str, Tensor], device, prefix: str = "model.visual_model."):
        self.device = torch.device(device)
        dev = self.device
        d = lambda k: _dev(sd[prefix + k], dev)
        f32 = lambda k: sd[prefix + k].detach().to(device=dev, dtype=torch.float32)
        pe_name = "prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"
        self.gauss = f32(pe_name).contiguous()
        self.out_tokens = torch.cat([d("mask_decoder.iou_token.weight"), d("mask_decoder.mask_tokens.weight")], 0).contiguous()
        self.point_embed = d("prompt_encoder.point_embeddings.1.weight").reshape(-1).contiguous()
        self.not_a_point = d("prompt_encoder.not_a_point_embed.weight").reshape(-1).contiguous()
        self.no_mask = d("prompt_encoder.no_mask_embed.weight").reshape(1, -1).contiguous()
        # dense positional encoding of the 64 x 64 feature grid (prompt_encoder.py:197-210), token-major fp32
        g = 64
        c = (torch.arange(g, device=dev, dtype=torch.float32) + 0.5) / g
        xy = torch.stack([c[None, :].expand(g, g), c[:, None].expand(g, g)], dim=-1).reshape(-1, 2)
        arg = 2 * np.pi * ((2 * xy - 1) @ self.gauss)
        pe = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)                       # [4096, 256]
        t = "mask_decoder.transformer."
        lin = lambda name: (d(name + ".weight"), d(name + ".bias"))
        pe_proj = lambda name: (pe @ f32(name + ".weight").T)                           # fp32 [4096, out]
        zeros = lambda n: torch.zeros(g * g, n, device=dev, dtype=torch.float32)
        self.layers = []
        for i in range(2):
            p = f"{t}layers.{i}."
            L = dict(
                w_sqkv=torch.cat([d(p + f"self_attn.{n}_proj.weight") for n in "qkv"], 0).contiguous(),
                b_sqkv=torch.cat([d(p + f"self_attn.{n}_proj.bias") for n in "qkv"], 0).contiguous(),
                so=lin(p + "self_attn.out_proj"),
                n1=lin(p + "norm1"), n2=lin(p + "norm2"), n3=lin(p + "norm3"), n4=lin(p + "norm4"),
                t2i_q=lin(p + "cross_attn_token_to_image.q_proj"), t2i_o=lin(p + "cross_attn_token_to_image.out_proj"),
                lin1=lin(p + "mlp.lin1"), lin2=lin(p + "mlp.lin2"),
                i2t_k=lin(p + "cross_attn_image_to_token.k_proj"), i2t_v=lin(p + "cross_attn_image_to_token.v_proj"),
                i2t_o=lin(p + "cross_attn_image_to_token.out_proj"),
                # the three projections of the image keys as one [384, 256] GEMM: token->image K, V and image->token Q;
                # their positional terms (K and Q only: V reads the keys without pe) as one broadcast residual
                w_img=torch.cat([d(p + "cross_attn_token_to_image.k_proj.weight"), d(p + "cross_attn_token_to_image.v_proj.weight"),
                                 d(p + "cross_attn_image_to_token.q_proj.weight")], 0).contiguous(),
                b_img=torch.cat([d(p + "cross_attn_token_to_image.k_proj.bias"), d(p + "cross_attn_token_to_image.v_proj.bias"),
                                 d(p + "cross_attn_image_to_token.q_proj.bias")], 0).contiguous(),
                pe_img=torch.cat([pe_proj(p + "cross_attn_token_to_image.k_proj"), zeros(128),
                                  pe_proj(p + "cross_attn_image_to_token.q_proj")], dim=1).to(BF16).contiguous(),
            )
            self.layers.append(L)
        f = t + "final_attn_token_to_image."
        self.fin_q, self.fin_o, self.fin_n = lin(f + "q_proj"), lin(f + "out_proj"), lin(t + "norm_final_attn")
        self.w_fin = torch.cat([d(f + "k_proj.weight"), d(f + "v_proj.weight")], 0).contiguous()
        self.b_fin = torch.cat([d(f + "k_proj.bias"), d(f + "v_proj.bias")], 0).contiguous()
        self.pe_fin = torch.cat([pe_proj(f + "k_proj"), zeros(128)], dim=1).to(BF16).contiguous()
        m = "mask_decoder."
        # ConvTranspose2d weight [in, out, kh, kw] -> GEMM weight [(kh, kw, out), in]; the bias repeats per sub-pixel
        w = f32(m + "output_upscaling.0.weight")
        self.w_up1 = w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0]).to(BF16).contiguous()
        self.b_up1 = f32(m + "output_upscaling.0.bias").repeat(4).to(BF16).contiguous()
        self.ln_up = lin(m + "output_upscaling.1")
        w = f32(m + "output_upscaling.3.weight")
        self.w_up2 = w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0]).to(BF16).contiguous()
        self.b_up2 = f32(m + "output_upscaling.3.bias").repeat(4).to(BF16).contiguous()
        self.hyper = [[lin(f"{m}output_hypernetworks_mlps.{i}.layers.{j}") for j in range(3)] for i in range(4)]
        iou = [lin(f"{m}iou_prediction_head.layers.{j}") for j in range(3)]
        # the 4-wide output layer padded to 8 columns (GEMM N % 8 == 0)
        w3 = torch.zeros(8, iou[2][0].shape[1], dtype=BF16, device=dev)
        w3[:4] = iou[2][0]
        b3 = torch.zeros(8, dtype=BF16, device=dev)
        b3[:4] = iou[2][1]
        self.iou_head = [iou[0], iou[1], (w3, b3)]
        self._off: Dict[int, Tensor] = {}

    # ---- mask decoder -------------------------------------------------------------------------------------
    def _offsets(self, P: int) -> Tensor:
        if P not in self._off:
            self._off[P] = (torch.arange(P + 1, dtype=torch.int32, device=self.device) * 7).contiguous()
        return self._off[P]

    def image_keys(self, emb_tokens: Tensor) -> dict:
        """Per-image part of the decoder: keys of block 0 (features + no-mask embedding, mask_decoder.py:134-135,
        prompt_encoder.py:231-235) and their three projections, shared by every prompt of the image."""
        assert emb_tokens.shape == (4096, 256) and emb_tokens.dtype == BF16
        keys0 = ops.add_rows_bcast(emb_tokens.contiguous(), self.no_mask, group=4096)
        L = self.layers[0]
        return {"keys0": keys0, "img0": ops.gemm(keys0, L["w_img"], L["b_img"], residual=L["pe_img"])}

    def decode(self, img: dict, points: Tensor, low_out: Optional[Tensor] = None):
        """points fp32 [P,2] (x, y) -> (low-res mask logits fp32 [P,3,256,256], IoU predictions fp32 [P,3])
        = `predict_torch(points[:,None], labels=1, multimask_output=True)` (reference predictor.py:166-241)."""
        P = points.shape[0]
        ln = lambda x, n: ops.layernorm(x, n[0], n[1], 1e-5)
        tokens = ops.point_tokens(points.contiguous(), self.gauss, self.out_tokens, self.point_embed, self.not_a_point, 1024.0)
        off = self._offsets(P)
        queries, keys = tokens, img["keys0"]
        for i, L in enumerate(self.layers):
            # (1) self attention of the 7 tokens (block 0: no positional term, no residual — transformer.py:154-160)
            if i == 0:
                qkv = ops.gemm(queries, L["w_sqkv"], L["b_sqkv"])
                a = ops.small_attention(qkv[:, 0:256], qkv[:, 256:512], qkv[:, 512:768], off, off, batch=P, heads=8, max_kv=7)
                queries = ln(ops.gemm(a, L["so"][0], L["so"][1]), L["n1"])
            else:
                qpe = ops.add_rows_bcast(queries, tokens, group=1)
                qk = ops.gemm(qpe, L["w_sqkv"][:512], L["b_sqkv"][:512])
                v = ops.gemm(queries, L["w_sqkv"][512:], L["b_sqkv"][512:])
                a = ops.small_attention(qk[:, 0:256], qk[:, 256:512], v, off, off, batch=P, heads=8, max_kv=7)
                queries = ln(ops.gemm(a, L["so"][0], L["so"][1], residual=queries), L["n1"])
            # image-side projections of this block's keys: once per image in block 0, per prompt afterwards
            if i == 0:
                proj, shared = img["img0"], True
            else:
                proj, shared = ops.gemm(keys, L["w_img"], L["b_img"], residual=L["pe_img"], res_mod=4096), False
            # (2) tokens attend to the image
            tq = ops.gemm(ops.add_rows_bcast(queries, tokens, group=1), L["t2i_q"][0], L["t2i_q"][1])
            a = ops.tok2img_attention(tq, proj[:, 0:128], proj[:, 128:256], P, shared)
            queries = ln(ops.gemm(a, L["t2i_o"][0], L["t2i_o"][1], residual=queries), L["n2"])
            # (3) MLP on the tokens
            h = ops.gemm(queries, L["lin1"][0], L["lin1"][1], act="relu")
            queries = ln(ops.gemm(h, L["lin2"][0], L["lin2"][1], residual=queries), L["n3"])
            # (4) the image attends to the tokens
            tk = ops.gemm(ops.add_rows_bcast(queries, tokens, group=1), L["i2t_k"][0], L["i2t_k"][1])
            tv = ops.gemm(queries, L["i2t_v"][0], L["i2t_v"][1])
            a_img = ops.img2tok_attention(proj[:, 256:384], tk, tv, P, shared)
            keys = ln(ops.gemm(a_img, L["i2t_o"][0], L["i2t_o"][1], residual=keys, res_mod=4096 if shared else 0), L["n4"])
            del proj, a_img
        # final token -> image attention (transformer.py:98-105)
        tq = ops.gemm(ops.add_rows_bcast(queries, tokens, group=1), self.fin_q[0], self.fin_q[1])
        kv = ops.gemm(keys, self.w_fin, self.b_fin, residual=self.pe_fin, res_mod=4096)
        a = ops.tok2img_attention(tq, kv[:, 0:128], kv[:, 128:256], P, False)
        hs = ln(ops.gemm(a, self.fin_o[0], self.fin_o[1], residual=queries), self.fin_n)        # [P*7, 256]
        del kv
        # up-scaling (mask_decoder.py:56-64,141-142): two 2 x 2 transposed convolutions as GEMMs on un-shuffled rows
        u1 = ops.gemm(keys, self.w_up1, self.b_up1)                                             # [P*4096, 4 x 64]
        if not FUSED_UPSCALE:
            ops.ln64_gelu(u1, self.ln_up[0], self.ln_up[1], 1e-6)
            u2 = ops.gemm(u1.view(P * 16384, 64), self.w_up2, self.b_up2, act="gelu")           # [P*16384, 4 x 32]
            del u1
        # hyper-networks on the 4 mask tokens, IoU head on the IoU token (mask_decoder.py:143-162)
        hs7 = hs.view(P, 7 * 256)
        hyper = torch.empty((P, 4, 32), dtype=BF16, device=self.device)

        def mlp3(x, layers, out=None):
            x = ops.gemm(x, layers[0][0], layers[0][1], act="relu")
            x = ops.gemm(x, layers[1][0], layers[1][1], act="relu")
            return ops.gemm(x, layers[2][0], layers[2][1], out=out)
        for i in range(4):
            mlp3(hs7[:, (1 + i) * 256:(2 + i) * 256], self.hyper[i], out=hyper[:, i, :])
        iou = mlp3(hs7[:, 0:256], self.iou_head)
        if FUSED_UPSCALE:   # LayerNorm2d + GELU, ConvTranspose #2 + GELU and the hyper-network product in one kernel
            low = ops.upscale_logits(u1, self.ln_up[0], self.ln_up[1], self.w_up2, self.b_up2, hyper, P, 1e-6, out=low_out)
        else:
            low = ops.mask_logits(u2, hyper, P, out=low_out)
        return low, iou[:, 1:4].float()

    # ---- automatic mask generator ----------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, emb_tokens: Tensor, *, points_per_side: int = 32, points_per_batch: int = 256,
                 pred_iou_thresh: float = 0.88, stability_score_thresh: float = 0.95, stability_score_offset: float = 1.0,
                 box_nms_thresh: float = 0.7, top_k: int = 50, return_masks: bool = False, low_res=None, iou_preds=None) -> dict:
        """emb_tokens bf16 [4096,256]: SAM encoder features of ONE image (token-major, `SamEncoder.forward(...)[i]`).
        `points_per_batch` only bounds memory (prompts are independent; the reference's default is 64).
        low_res / iou_preds: skip the decoder and post-process these instead (tests).
        -> {"segs" bf16 [K,256,256] (largest mask first), "boxes" int64 [K,4] XYXY, "iou_preds", "stability", "areas",
            "points" [K,2], "n_masks" (records after NMS, before the top-k cut), "masks" uint8 [K,1024,1024] on request}"""
        dev = self.device
        pts = point_grid(points_per_side)
        P = pts.shape[0]
        if low_res is None:
            img = self.image_keys(emb_tokens)
            low_res = torch.empty((P, 3, 256, 256), dtype=torch.float32, device=dev)
            iou_parts = []
            pts_dev = torch.from_numpy(pts).to(dev)
            for i in range(0, P, points_per_batch):
                _, iou = self.decode(img, pts_dev[i:i + points_per_batch], low_out=low_res[i:i + points_per_batch])
                iou_parts.append(iou)
            iou_preds = torch.cat(iou_parts, dim=0)
        cand_logits = low_res.reshape(P * 3, 256, 256)
        empty = {"segs": torch.zeros((0, 256, 256), dtype=BF16, device=dev), "boxes": torch.zeros((0, 4), dtype=torch.int64),
                 "iou_preds": torch.zeros(0), "stability": torch.zeros(0), "areas": torch.zeros(0, dtype=torch.int64),
                 "points": torch.zeros((0, 2)), "n_masks": 0, "candidates": torch.zeros(0, dtype=torch.int64)}
        # ---- host: <= 3 * P records (reference automatic_mask_generator.py:288-304).  The predicted-IoU filter comes
        # first, as upstream, so the up-sampled statistics are only evaluated for the candidates that pass it
        iou_np = iou_preds.reshape(-1).float().cpu().numpy()                                           # sync 1
        idx = np.arange(P * 3)
        keep = iou_np > pred_iou_thresh if pred_iou_thresh > 0.0 else np.ones(P * 3, dtype=bool)
        pre = idx[keep]
        if pre.size == 0:
            return empty
        if pre.size == P * 3:
            stats = ops.mask_stats(cand_logits, None, 0.0, stability_score_offset).cpu().numpy()      # sync 2
        else:
            stats = np.zeros((P * 3, 8), dtype=np.int32)
            stats[pre] = ops.mask_stats(cand_logits, torch.from_numpy(pre.astype(np.int32)).to(dev), 0.0,
                                        stability_score_offset).cpu().numpy()
        with np.errstate(divide="ignore", invalid="ignore"):
            stab = stats[:, 1].astype(np.float32) / stats[:, 2].astype(np.float32)
        if stability_score_thresh > 0.0:
            keep &= stab >= np.float32(stability_score_thresh)
        idx = idx[keep]
        if idx.size == 0:
            return empty
        boxes = np.stack([1023 - stats[:, 3], 1023 - stats[:, 4], stats[:, 5], stats[:, 6]], axis=1)
        boxes[stats[:, 0] == 0] = 0
        order = idx[np.argsort(-iou_np[idx], kind="stable")]                                           # score order
        if order.size > 4096:
            raise ValueError(f"{order.size} candidates after filtering; box NMS handles at most 4096")
        keep_nms = ops.box_nms(torch.from_numpy(boxes[order].astype(np.float32)).to(dev), box_nms_thresh).cpu().numpy()  # sync 3
        kept = order[keep_nms.astype(bool)]
        areas = stats[kept, 0].astype(np.int64)
        top = kept[np.argsort(-areas, kind="stable")[:top_k]]                                          # largest first
        cand = torch.from_numpy(top.astype(np.int32)).to(dev)
        out = {"segs": ops.mask_soft(cand_logits, cand, 0.0),
               "boxes": torch.from_numpy(boxes[top].astype(np.int64)), "iou_preds": torch.from_numpy(iou_np[top]),
               "stability": torch.from_numpy(stab[top]), "areas": torch.from_numpy(stats[top, 0].astype(np.int64)),
               "points": torch.from_numpy(pts[top // 3]), "n_masks": int(kept.size), "candidates": torch.from_numpy(top.astype(np.int64))}
        if return_masks:
            out["masks"] = ops.mask_binarize(cand_logits, cand, 0.0)
        return out
