"""Mask-proposal selector on the sm_100a kernels (reference model/LISA.py:350-408 with the blocks
of model/transformer.py:215-341).  All images of a batch are processed together: mask tokens are
concatenated row-wise ([ΣK_i, 256]) with int32 offsets, so every linear is one tcgen05 GEMM and
the ≤128-token attentions run in `llmseg_small_attention`.

Exact simplifications (same math, fewer bytes — SURVEY §A.4):
  * upsample∘pool is evaluated in its adjoint form (ops.maskpool)
  * attention over the single text key has softmax ≡ 1, so mask→text cross attention and the final
    attention are out_proj(v_proj(text)) broadcast over the mask tokens (their q/k projections are dead)
  * text_hidden_fcs runs on the gathered [SEG] rows only
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops
from .encoders import BF16, _dev

Tensor = torch.Tensor


class Selector:
    def __init__(self, sd: Dict[str, Tensor], device, prefix: str = ""):
        self.device = device
        d = lambda k: _dev(sd[prefix + k], device)
        self.fc0 = (d("text_hidden_fcs.0.0.weight"), d("text_hidden_fcs.0.0.bias"))
        self.fc2 = (d("text_hidden_fcs.0.2.weight"), d("text_hidden_fcs.0.2.bias"))
        self.blocks = []
        for i in range(2):
            p = f"lisa_attention_layers.{i}."
            cat = lambda att, names, suffix: torch.cat([d(p + f"{att}.{n}_proj.{suffix}") for n in names], 0).contiguous()
            self.blocks.append(dict(
                w_qkv=cat("self_attn", "qkv", "weight"), b_qkv=cat("self_attn", "qkv", "bias"),
                w_so=d(p + "self_attn.out_proj.weight"), b_so=d(p + "self_attn.out_proj.bias"),
                n1=(d(p + "norm1.weight"), d(p + "norm1.bias")),
                w_tv=d(p + "cross_attn_token_to_image.v_proj.weight"), b_tv=d(p + "cross_attn_token_to_image.v_proj.bias"),
                w_to=d(p + "cross_attn_token_to_image.out_proj.weight"), b_to=d(p + "cross_attn_token_to_image.out_proj.bias"),
                n2=(d(p + "norm2.weight"), d(p + "norm2.bias")),
                w1=d(p + "mlp.lin1.weight"), b1=d(p + "mlp.lin1.bias"),
                w2=d(p + "mlp.lin2.weight"), b2=d(p + "mlp.lin2.bias"),
                n3=(d(p + "norm3.weight"), d(p + "norm3.bias")),
                w_iq=d(p + "cross_attn_image_to_token.q_proj.weight"), b_iq=d(p + "cross_attn_image_to_token.q_proj.bias"),
                w_ikv=cat("cross_attn_image_to_token", "kv", "weight"), b_ikv=cat("cross_attn_image_to_token", "kv", "bias"),
                w_io=d(p + "cross_attn_image_to_token.out_proj.weight"), b_io=d(p + "cross_attn_image_to_token.out_proj.bias"),
                n4=(d(p + "norm4.weight"), d(p + "norm4.bias")),
            ))
        self.w_fv, self.b_fv = d("lisa_final_attn.v_proj.weight"), d("lisa_final_attn.v_proj.bias")
        self.w_fo, self.b_fo = d("lisa_final_attn.out_proj.weight"), d("lisa_final_attn.out_proj.bias")
        self.n_fin = (d("lisa_norm_final_attn.weight"), d("lisa_norm_final_attn.bias"))
        self.w_i1, self.b_i1 = d("lisa_iou_head.0.weight"), d("lisa_iou_head.0.bias")
        self.w_i2 = d("lisa_iou_head.2.weight").reshape(-1).contiguous()
        self.b_i2 = torch.zeros(8, dtype=BF16, device=device)
        self.b_i2[0] = d("lisa_iou_head.2.bias")[0]
        self.w_e1, self.b_e1 = d("lisa_embedding_head.0.weight"), d("lisa_embedding_head.0.bias")
        self.w_e2, self.b_e2 = d("lisa_embedding_head.2.weight"), d("lisa_embedding_head.2.bias")

    def text_embed(self, hidden_rows: Tensor, out: Tensor = None) -> Tensor:
        """text_hidden_fcs on the gathered rows (reference LISA.py:56-65,317-337)."""
        t = ops.gemm(hidden_rows, self.fc0[0], self.fc0[1], act="relu")
        return ops.gemm(t, self.fc2[0], self.fc2[1], out=out)

    def make_plan(self, Ks, k_cap: int = 0) -> dict:
        """Device-side index tensors for a batch with Ks[i] proposals per image (built outside any CUDA-graph
        capture).  With k_cap > 0 the plan is sized for UP TO k_cap proposals per image (rows = B * k_cap, the
        proposals stay concatenated at the front) and `update_plan` re-targets it to other counts in place, so a
        captured graph serves every call of its (B, k_cap) bucket."""
        Ks = [int(k) for k in Ks]
        dev, B = self.device, len(Ks)
        if B == 0:
            raise ValueError("selector plan needs at least one image")
        kmax = k_cap if k_cap > 0 else max(Ks)
        if kmax > 128 or max(Ks) > kmax:
            raise ValueError(f"selector kernels support at most 128 proposals per image, got {max(max(Ks), kmax)}")
        rows = B * k_cap if k_cap > 0 else sum(Ks)
        plan = {
            "Ks": None, "kmax": kmax, "B": B, "rows": rows,
            "k_off": torch.zeros(B + 1, dtype=torch.int32, device=dev),
            "b_off": torch.arange(B + 1, dtype=torch.int32, device=dev),
            "mask_image": torch.zeros(max(rows, 1), dtype=torch.int32, device=dev),
        }
        return self.update_plan(plan, Ks)

    def update_plan(self, plan: dict, Ks) -> dict:
        Ks = [int(k) for k in Ks]
        if plan["Ks"] == Ks:
            return plan
        if len(Ks) != plan["B"] or min(Ks) < 1 or max(Ks) > plan["kmax"] or sum(Ks) > plan["rows"]:
            raise ValueError(f"proposal counts {Ks} do not fit the selector plan (B={plan['B']}, K<={plan['kmax']})")
        offs = [0]
        for kk in Ks:
            offs.append(offs[-1] + kk)
        img = torch.repeat_interleave(torch.arange(len(Ks), dtype=torch.int32), torch.tensor(Ks))
        # rows beyond the last proposal (capacity padding) point at the last image: they are computed and ignored
        pad = torch.full((plan["rows"] - offs[-1],), len(Ks) - 1, dtype=torch.int32)
        plan["k_off"].copy_(torch.tensor(offs, dtype=torch.int32))
        plan["mask_image"][:plan["rows"]].copy_(torch.cat([img, pad]))
        plan["Ks"] = Ks
        return plan

    def forward(self, emb_tokens: Tensor, seg_cat: Tensor, text_embed: Tensor, plan: dict, **conv):
        """emb_tokens [B,4096,256]; seg_cat [rows,256,256] bf16 (proposals of all images, concatenated);
        text_embed [B,256] (conversation 0 of each image); plan from make_plan.
        -> (sim fp32 [n_conv,Kmax], iou fp32 [B,Kmax], best int32 [B])."""
        feat = ops.maskpool(seg_cat.contiguous(), emb_tokens.contiguous(), plan["mask_image"])
        return self.forward_pooled(feat, text_embed, plan, **conv)

    def forward_pooled(self, feat: Tensor, text_embed: Tensor, plan: dict, *, text_all: Tensor = None,
                       conv_group: Tensor = None, conv_valid: Tensor = None):
        """The selector after mask pooling, over G groups of mask tokens (plan = make_plan(K per group)):
        feat [rows,256] pooled proposal features, text_embed [G,256].  At inference a group is an image; in the
        training forward it is an (image, round) pair sharing its image's pooled features (the reference expands
        them per conversation, LISA.py:372-375).
        text_all [n_conv,256] + conv_group int32 [n_conv]: every conversation of a group is scored against the
        group's mask embeddings (LISA.py:397-403 — they were updated with the group's first conversation,
        `sam_segs_feature_list[b][0]`); default: one conversation per group."""
        B, kmax = plan["B"], plan["kmax"]
        k_off, b_off, mask_image = plan["k_off"], plan["b_off"], plan["mask_image"]
        text = text_embed
        ln = lambda x, n: ops.layernorm(x, n[0], n[1], 1e-5)
        for blk in self.blocks:
            qkv = ops.gemm(feat, blk["w_qkv"], blk["b_qkv"])
            a = ops.small_attention(qkv[:, 0:256], qkv[:, 256:512], qkv[:, 512:768], k_off, k_off,
                                    batch=B, heads=8, max_kv=kmax)
            feat = ln(ops.gemm(a, blk["w_so"], blk["b_so"], residual=feat), blk["n1"])
            to = ops.gemm(ops.gemm(text, blk["w_tv"], blk["b_tv"]), blk["w_to"], blk["b_to"])
            feat = ln(ops.add_rows_bcast(feat, to, row_group=mask_image), blk["n2"])
            h = ops.gemm(feat, blk["w1"], blk["b1"], act="relu")
            feat = ln(ops.gemm(h, blk["w2"], blk["b2"], residual=feat), blk["n3"])
            tq = ops.gemm(text, blk["w_iq"], blk["b_iq"])
            kv = ops.gemm(feat, blk["w_ikv"], blk["b_ikv"])
            a = ops.small_attention(tq, kv[:, 0:256], kv[:, 256:512], b_off, k_off, batch=B, heads=8, max_kv=kmax)
            text = ln(ops.gemm(a, blk["w_io"], blk["b_io"], residual=text), blk["n4"])
        to = ops.gemm(ops.gemm(text, self.w_fv, self.b_fv), self.w_fo, self.b_fo)
        feat = ln(ops.add_rows_bcast(feat, to, row_group=mask_image), self.n_fin)
        h_iou = ops.gemm(feat, self.w_i1, self.b_i1, act="relu")
        e = ops.gemm(ops.gemm(feat, self.w_e1, self.b_e1, act="relu"), self.w_e2, self.b_e2)
        return ops.select(e, text_embed if text_all is None else text_all, h_iou, self.w_i2, self.b_i2, k_off,
                          batch=B, k_stride=kmax, conv_group=conv_group, conv_valid=conv_valid)
