"""Generate tests/golden/*.pt by running the REFERENCE's own modules (imported from /root/reference)
and the installed transformers eager modules, and check the oracle restatements against them.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference ships no tests or golden vectors, so these fixtures are the pin.

Each fixture stores inputs, the reference output, and either the (tiny) weights or the seed that
regenerates them through the oracle's `random_state_dict` plus a checksum of the regenerated weights.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

sys.path.insert(0, str(ROOT))
from oracle import clip_llama, dinov2, lisa_forward, sam_amg, sam_encoder, selector  # noqa: E402


def checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def _ref_imports():
    sys.path.insert(0, str(REF))
    from model import loss as ref_loss  # type: ignore
    from model import transformer as ref_tr  # type: ignore
    from model.segment_anything.modeling import image_encoder as ref_ie  # type: ignore
    return ref_ie, ref_tr, ref_loss


def gold_sam(ref_ie, name, cfg: sam_encoder.SamConfig, seed: int, store_weights: bool):
    torch.manual_seed(seed)
    m = ref_ie.ImageEncoderViT(
        img_size=cfg.img_size, patch_size=cfg.patch_size, embed_dim=cfg.embed_dim, depth=cfg.depth,
        num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio, out_chans=cfg.out_chans, qkv_bias=True,
        norm_layer=lambda d: torch.nn.LayerNorm(d, eps=cfg.ln_eps), use_rel_pos=True,
        window_size=cfg.window_size, global_attn_indexes=cfg.global_attn_indexes).eval()
    sd = sam_encoder.random_state_dict(cfg, seed)
    missing = m.load_state_dict(sd, strict=True)
    x = torch.randn(1, 3, cfg.img_size, cfg.img_size, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        y_ref = m(x)
        y = sam_encoder.image_encoder(x, sd, cfg)
    err = (y - y_ref).abs().max().item()
    print(f"[sam:{name}] oracle vs reference ImageEncoderViT max|d| = {err:.3e}  (out {tuple(y_ref.shape)})")
    assert err < 2e-4, err
    fx = {"cfg": cfg.__dict__, "seed": seed, "x_seed": seed + 1, "out": y_ref, "weights_checksum": checksum(sd)}
    if store_weights:
        fx["sd"] = sd
        fx["x"] = x
    torch.save(fx, GOLD / f"sam_{name}.pt")


def gold_relpos(ref_ie):
    g = torch.Generator().manual_seed(7)
    out = {}
    for S, hd in ((14, 80), (5, 16)):
        q = torch.randn(3, S * S, hd, generator=g)
        rh = torch.randn(2 * S - 1, hd, generator=g) * 0.2
        rw = torch.randn(2 * S - 1, hd, generator=g) * 0.2
        attn0 = torch.zeros(3, S * S, S * S)
        ref = ref_ie.add_decomposed_rel_pos(attn0, q, rh, rw, (S, S), (S, S))
        mine = sam_encoder.decomposed_rel_pos_bias(q, rh, rw, (S, S))
        err = (ref - mine).abs().max().item()
        print(f"[relpos S={S}] max|d| = {err:.3e}")
        assert err < 1e-5
        out[f"S{S}"] = {"q": q, "rel_h": rh, "rel_w": rw, "bias": ref}
    # window partition / unpartition round trip with padding
    x = torch.randn(2, 7, 7, 4, generator=g)
    w_ref, pad = ref_ie.window_partition(x, 3)
    w_mine, pad2 = sam_encoder.partition_windows(x, 3)
    assert pad == pad2 and torch.equal(w_ref, w_mine)
    back = ref_ie.window_unpartition(w_ref, 3, pad, (7, 7))
    assert torch.equal(back, sam_encoder.unpartition_windows(w_mine, 3, pad2, (7, 7)))
    out["partition"] = {"x": x, "windows": w_ref}
    torch.save(out, GOLD / "sam_relpos.pt")


def _load_selector_ref(ref_tr, sd):
    blocks = []
    for i in range(2):
        b = ref_tr.LISA_TwoWayAttentionBlock(embedding_dim=256, num_heads=8, mlp_dim=2048,
                                             attention_downsample_rate=1).eval()
        b.load_state_dict(lisa_forward.sub_dict(sd, f"lisa_attention_layers.{i}."), strict=True)
        blocks.append(b)
    fin = ref_tr.Attention(embedding_dim=256, num_heads=8, downsample_rate=1).eval()
    fin.load_state_dict(lisa_forward.sub_dict(sd, "lisa_final_attn."), strict=True)
    return blocks, fin


def gold_selector(ref_tr):
    """Mirrors reference model/LISA.py:350-408 using the reference's own transformer.py modules."""
    for K, seed in ((32, 11), (64, 12), (7, 13)):
        sd = selector.random_state_dict(seed, hidden=64)
        blocks, fin = _load_selector_ref(ref_tr, sd)
        emb, segs, hidden = selector.synthetic_case(seed, K, 64)
        with torch.no_grad():
            text = selector.text_hidden_fc(hidden, sd)
            # --- reference-side computation (LISA.py:350-408) with reference modules
            up = F.interpolate(emb.float(), size=(256, 256), mode="bilinear", align_corners=False)
            e = up[0].flatten(1, 2)
            w = segs.flatten(1, 2)
            feat = (w @ e.T) / (w.sum(-1, keepdim=True) + 1e-8)
            q, t = feat.unsqueeze(0), text.unsqueeze(1)
            for b in blocks:
                q, t = b(queries=q, keys=t)
            q = F.layer_norm(q + fin(q=q, k=t, v=t), (256,), sd["lisa_norm_final_attn.weight"], sd["lisa_norm_final_attn.bias"])
            iou = torch.sigmoid(F.linear(F.relu(F.linear(q, sd["lisa_iou_head.0.weight"], sd["lisa_iou_head.0.bias"])),
                                         sd["lisa_iou_head.2.weight"], sd["lisa_iou_head.2.bias"]))
            em = F.linear(F.relu(F.linear(q, sd["lisa_embedding_head.0.weight"], sd["lisa_embedding_head.0.bias"])),
                          sd["lisa_embedding_head.2.weight"], sd["lisa_embedding_head.2.bias"])
            tn = text / text.norm(dim=-1, keepdim=True)
            fn = em[0] / em[0].norm(dim=-1, keepdim=True)
            sim_ref, iou_ref = tn @ fn.T, iou[0].T
            # --- oracle restatement
            sim, io = selector.selector_forward(selector.upsample_embeddings(emb)[0], segs, text, sd)
        e1, e2 = (sim - sim_ref).abs().max().item(), (io - iou_ref).abs().max().item()
        print(f"[selector K={K}] sim max|d|={e1:.3e} iou max|d|={e2:.3e}")
        assert e1 < 1e-5 and e2 < 1e-5
        torch.save({"seed": seed, "K": K, "hidden_dim": 64, "weights_checksum": checksum(sd),
                    "inputs_checksum": float(emb.double().sum() + segs.double().sum() + hidden.double().sum()),
                    "feat": feat,
                    "pred_similarity": sim_ref, "pred_iou": iou_ref}, GOLD / f"selector_K{K}.pt")


def gold_losses(ref_loss):
    g = torch.Generator().manual_seed(3)
    K = 50
    pe, te = torch.randn(K, 256, generator=g), torch.randn(1, 256, generator=g)
    gi, pi = torch.rand(K, 1, generator=g), torch.rand(K, 1, generator=g)
    logits = torch.randn(3, 32, 32, generator=g)
    tgt = (torch.rand(3, 32, 32, generator=g) > 0.5).float()
    ref = {
        "softmax_align": ref_loss.softmax_align_loss(pe, te, gi),
        "iou_regression": ref_loss.iou_regression_loss(pi, gi),
        "dice": ref_loss.dice_loss(logits, tgt, 3.0),
        "sigmoid_ce": ref_loss.sigmoid_ce_loss(logits, tgt, 3.0),
    }
    mine = {
        "softmax_align": lisa_forward.softmax_align_loss(pe, te, gi),
        "iou_regression": lisa_forward.iou_regression_loss(pi, gi),
        "dice": lisa_forward.dice_loss(logits, tgt, 3.0),
        "sigmoid_ce": lisa_forward.sigmoid_ce_loss(logits, tgt, 3.0),
    }
    for k in ref:
        d = abs(float(ref[k]) - float(mine[k]))
        print(f"[loss {k}] ref={float(ref[k]):.6f} |d|={d:.2e}")
        assert d < 1e-5
    torch.save({"pe": pe, "te": te, "gt_ious": gi, "pred_ious": pi, "logits": logits, "targets": tgt,
                "expected": {k: float(v) for k, v in ref.items()}}, GOLD / "losses.pt")


def gold_clip():
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = clip_llama.ClipConfig(image_size=56, patch_size=14, hidden=64, layers=4, heads=4, mlp=128)
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4,
                                          num_attention_heads=4, image_size=56, patch_size=14,
                                          hidden_act="quick_gelu", layer_norm_eps=1e-5,
                                          attn_implementation="eager")).eval()
    sd = clip_llama.clip_random_state_dict(cfg, seed=21)
    full = dict(sd)
    full["vision_model.post_layernorm.weight"] = torch.ones(64)
    full["vision_model.post_layernorm.bias"] = torch.zeros(64)
    hf_sd = hf.state_dict()
    for k in hf_sd:
        if k not in full:  # e.g. position_ids buffers
            full[k] = hf_sd[k]
    hf.load_state_dict(full, strict=True)
    x = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(22))
    with torch.no_grad():
        ref = hf(x, output_hidden_states=True).hidden_states[-2][:, 1:]
        mine = clip_llama.clip_patch_features(x, sd, cfg)
    err = (ref - mine).abs().max().item()
    print(f"[clip tiny] oracle vs transformers eager hidden_states[-2] max|d| = {err:.3e}")
    assert err < 1e-4
    torch.save({"cfg": cfg.__dict__, "sd": sd, "x": x, "out": ref}, GOLD / "clip_tiny.pt")


def gold_llama():
    from transformers import LlamaConfig, LlamaModel
    cfg = clip_llama.LlamaConfig(hidden=64, layers=3, heads=4, mlp=176, vocab=100)
    hf = LlamaModel(LlamaConfig(hidden_size=64, intermediate_size=176, num_hidden_layers=3,
                                num_attention_heads=4, num_key_value_heads=4, vocab_size=100,
                                rms_norm_eps=1e-6, rope_theta=10000.0, attn_implementation="eager")).eval()
    sd = clip_llama.llama_random_state_dict(cfg, seed=31)
    hf_sd = hf.state_dict()
    full = dict(sd)
    for k in hf_sd:
        if k not in full:
            full[k] = hf_sd[k]
    hf.load_state_dict(full, strict=True)
    g = torch.Generator().manual_seed(32)
    emb = torch.randn(2, 23, 64, generator=g)
    mask = torch.ones(2, 23, dtype=torch.bool)
    mask[1, 17:] = False  # right padding
    with torch.no_grad():
        ref = hf(inputs_embeds=emb, attention_mask=mask.long()).last_hidden_state
        mine = clip_llama.llama_last_hidden(emb, mask, sd, cfg)
    valid = mask[:, :, None]
    err = ((ref - mine) * valid).abs().max().item()
    print(f"[llama tiny] oracle vs transformers eager last_hidden_state max|d| (valid rows) = {err:.3e}")
    assert err < 1e-4
    torch.save({"cfg": cfg.__dict__, "sd": sd, "embeds": emb, "mask": mask, "out": ref}, GOLD / "llama_tiny.pt")


def gold_dinov2():
    """Variant-B image encoder (hub DINOv2, absent offline) pinned against transformers' Dinov2Model."""
    from transformers import Dinov2Config as HFConfig, Dinov2Model
    # stored table 5x5 (image 70), evaluated on a 9x9 grid (image 126): the position table IS resampled
    cfg = dinov2.Dinov2Config(img_size=126, patch_size=14, embed_dim=64, depth=3, num_heads=4, train_grid=5,
                              out_chans=16, interpolate_offset=0.0)
    hf = Dinov2Model(HFConfig(hidden_size=64, num_hidden_layers=3, num_attention_heads=4, mlp_ratio=4,
                              image_size=70, patch_size=14, layer_norm_eps=1e-6, hidden_act="gelu",
                              layerscale_value=1.0, attn_implementation="eager")).eval()
    sd = dinov2.random_state_dict(cfg, seed=51)
    full = dinov2.to_hf_names(sd, cfg)
    hf_sd = hf.state_dict()
    missing = [k for k in hf_sd if k not in full]
    assert not missing, missing
    hf.load_state_dict(full, strict=True)
    g = torch.Generator().manual_seed(52)
    x = torch.randn(2, 3, 126, 126, generator=g)
    conv_w, conv_b = torch.randn(16, 64, 1, 1, generator=g) * 0.125, torch.randn(16, generator=g) * 0.1
    with torch.no_grad():
        ref = hf(x).last_hidden_state[:, 1:]                  # == x_norm_patchtokens
        mine = dinov2.forward_features(x, sd, cfg)
        emb = dinov2.image_embeddings(x, sd, conv_w, conv_b, cfg)
    err = (ref - mine).abs().max().item()
    print(f"[dinov2 tiny] oracle vs transformers eager x_norm_patchtokens max|d| = {err:.3e}")
    assert err < 1e-4
    # the hub default (offset 0.1) gives a slightly different table: record both so the test pins the switch
    pos01 = dinov2.interpolate_pos_embed(sd["pos_embed"], cfg.grid, 0.1)
    pos00 = dinov2.interpolate_pos_embed(sd["pos_embed"], cfg.grid, 0.0)
    assert pos01.shape == pos00.shape and (pos01 - pos00).abs().max() > 0
    torch.save({"cfg": cfg.__dict__, "sd": sd, "x": x, "tokens": ref, "conv_w": conv_w, "conv_b": conv_b,
                "embeddings": emb, "pos_offset01": pos01}, GOLD / "dinov2_tiny.pt")


def gold_splice():
    """Index arithmetic of SURVEY §A.6 on a worked example (no reference module is importable for it)."""
    cfg = lisa_forward.LisaConfig()
    ids = torch.tensor([[1, 32001, -200, 32002, 5, 6, 7, 32000, 9, 2]])
    m = lisa_forward.seg_token_mask(ids, cfg)
    s = 7
    assert m.shape == (1, 10 + 255) and m[0].nonzero().flatten().tolist() == [s + 254]
    emb_table = torch.arange(40000, dtype=torch.float32)[:, None].repeat(1, 2)
    feats = -torch.arange(1, 257, dtype=torch.float32)[None, :, None].repeat(1, 1, 2)
    e, am = lisa_forward.splice_inputs(ids, torch.ones(1, 10, dtype=torch.bool), feats, emb_table)
    assert e.shape == (1, 265, 2) and am.shape == (1, 265)
    assert e[0, :2, 0].tolist() == [1, 32001] and e[0, 2, 0] == -1 and e[0, 257, 0] == -256
    assert e[0, 258, 0] == 32002 and e[0, s + 255, 0] == 32000 and e[0, s + 254, 0] == 7
    print("[splice] index arithmetic ok")


def gold_sam_amg():
    """SAM-Everything: the reference's own PromptEncoder + MaskDecoder on seeded point prompts, and its
    own SamAutomaticMaskGenerator.generate() end to end (image encoder replaced by a stub that returns a seeded
    embedding: the ViT-H encoder is pinned separately by sam_*.pt) — against oracle/sam_amg.py."""
    import numpy as np
    sys.path.insert(0, str(REF))
    from model.segment_anything import SamAutomaticMaskGenerator  # type: ignore
    from model.segment_anything.modeling import MaskDecoder, PromptEncoder, Sam, TwoWayTransformer  # type: ignore

    seed = 11
    sd = sam_amg.random_state_dict(seed)

    class StubEncoder(torch.nn.Module):
        img_size = 1024

        def __init__(self, emb):
            super().__init__()
            self.emb = emb

        def forward(self, x):
            return self.emb

    g = torch.Generator().manual_seed(seed + 1)
    # a smooth random field + noise: mask logits with coherent regions, so boxes / NMS / areas are exercised
    low = torch.randn(1, 256, 8, 8, generator=g)
    emb = F.interpolate(low, size=(64, 64), mode="bilinear", align_corners=False) + 0.3 * torch.randn(1, 256, 64, 64, generator=g)
    sam = Sam(image_encoder=StubEncoder(emb),
              prompt_encoder=PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024),
                                           mask_in_chans=16),
              mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                       transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                                       transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256),
              pixel_mean=[123.675, 116.28, 103.53], pixel_std=[58.395, 57.12, 57.375]).eval()
    # The vendored predictor calls `prompt_encoder(points=, boxes=, masks=)` (predictor.py:233-237) while LISA's fork of
    # the encoder added a required `text_embeds` argument (prompt_encoder.py:128-134) — the reference's own proposal
    # scripts therefore run the pip `segment_anything` package, whose encoder has no such argument.  Default it to None.
    _pe_forward = sam.prompt_encoder.forward
    sam.prompt_encoder.forward = lambda points, boxes, masks, text_embeds=None: _pe_forward(points, boxes, masks, text_embeds)
    res = sam.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith(("image_encoder", "prompt_encoder.mask_downscaling", "pixel_")) for k in res.missing_keys), res.missing_keys
    # (1) point prompts -> low-res mask logits + IoU predictions
    pts = torch.tensor([[100.0, 200.0], [512.5, 512.5], [900.0, 40.0], [16.0, 1000.0], [700.25, 333.0]])
    with torch.no_grad():
        sparse, dense = sam.prompt_encoder(points=(pts[:, None, :], torch.ones(5, 1, dtype=torch.int)), boxes=None, masks=None,
                                           text_embeds=None)
        low_ref, iou_ref = sam.mask_decoder(image_embeddings=emb, image_pe=sam.prompt_encoder.get_dense_pe(),
                                            sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                                            multimask_output=True)
        low, iou = sam_amg.predict_points(emb, pts, sd)
        e_sparse = (sam_amg.embed_points(pts, sd) - sparse).abs().max().item()
    e_low, e_iou = (low - low_ref).abs().max().item(), (iou - iou_ref).abs().max().item()
    print(f"[sam_amg] oracle vs reference PromptEncoder/MaskDecoder: sparse {e_sparse:.2e} masks {e_low:.2e} (|masks| max "
          f"{low_ref.abs().max().item():.2f}) iou {e_iou:.2e}")
    assert e_sparse < 1e-5 and e_low < 2e-3 and e_iou < 1e-4
    # (2) the reference's generator end to end: 8 x 8 point grid.  Random decoder weights give masks that span the
    # image, so at the default box-NMS threshold a single mask survives; the second configuration (threshold 1.0:
    # nothing is suppressed) keeps every mask that passes the IoU / stability filters, which pins the per-mask
    # records; the suppression rule itself is pinned against torchvision on random boxes in (4).
    image = np.zeros((1024, 1024, 3), dtype=np.uint8)
    runs = {}
    for tag, nms_thr in (("nms07", 0.7), ("nms10", 1.0)):
        kw = dict(points_per_side=8, points_per_batch=16, pred_iou_thresh=-0.6, stability_score_thresh=0.5,
                  stability_score_offset=1.0, box_nms_thresh=nms_thr)
        amg = SamAutomaticMaskGenerator(sam, crop_n_layers=0, min_mask_region_area=0, output_mode="binary_mask", **kw)
        with torch.no_grad():
            anns = amg.generate(image)
            data = sam_amg.generate(emb, sd, **kw)
        print(f"[sam_amg:{tag}] reference generate(): {len(anns)} masks; oracle: {data['masks'].shape[0]}")
        assert len(anns) == data["masks"].shape[0] and len(anns) >= 1
        ref_masks = torch.from_numpy(np.stack([a["segmentation"] for a in anns]))
        ref_iou = torch.tensor([a["predicted_iou"] for a in anns])
        ref_stab = torch.tensor([a["stability_score"] for a in anns])
        ref_area = torch.tensor([a["area"] for a in anns])
        ref_box = torch.tensor([a["bbox"] for a in anns])           # XYWH
        ref_pts = torch.tensor([a["point_coords"][0] for a in anns])
        assert torch.equal(ref_masks, data["masks"]), "oracle masks differ from the reference generator's"
        assert (ref_iou - data["iou_preds"]).abs().max().item() < 1e-4 and (ref_stab - data["stability"]).abs().max().item() < 1e-5
        assert torch.equal(ref_area, data["areas"]) and (ref_pts - data["points"]).abs().max().item() < 1e-3
        b = data["boxes"].float()
        assert torch.equal(ref_box, torch.stack([b[:, 0], b[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]], dim=-1))
        runs[tag] = {"kw": kw, "n_masks": len(anns), "ref_iou": ref_iou, "ref_stability": ref_stab, "ref_area": ref_area,
                     "ref_box_xywh": ref_box, "ref_points": ref_pts}
    assert runs["nms10"]["n_masks"] >= 8
    # (3) LLM-Seg's consumer side: largest masks first, antialiased bilinear resize to 256 x 256
    soft, order = sam_amg.llmseg_proposals(data, top_k=50)
    lo, w = sam_amg.aa_downsample_weights(1024, 256)
    m0 = data["masks"][order[0]].double()
    rows = torch.stack([(m0[int(lo[i]):int(lo[i]) + w.shape[1]] * w[i, :m0[int(lo[i]):int(lo[i]) + w.shape[1]].shape[0], None]).sum(0)
                        for i in range(256)])
    sep = torch.stack([(rows[:, int(lo[i]):int(lo[i]) + w.shape[1]] * w[i, :rows[:, int(lo[i]):int(lo[i]) + w.shape[1]].shape[1]]).sum(1)
                       for i in range(256)], dim=1)
    e_aa = (sep.float() - soft[0]).abs().max().item()
    print(f"[sam_amg] explicit antialias filter vs F.interpolate(antialias=True): max|d| = {e_aa:.2e}")
    assert e_aa < 1e-5
    # (4) the suppression rule against torchvision's batched_nms (what automatic_mask_generator.py:256-262 calls)
    from torchvision.ops.boxes import batched_nms
    gb = torch.Generator().manual_seed(seed + 2)
    xy = torch.rand(300, 2, generator=gb) * 900
    wh = torch.rand(300, 2, generator=gb) * 300 + 4
    boxes = torch.cat([xy, xy + wh], dim=1).round()
    boxes[100:140] = boxes[:40] + torch.randint(-6, 7, (40, 4), generator=gb).float()      # near-duplicates
    scores = torch.rand(300, generator=gb)
    scores[200:210] = scores[0]                                                          # ties
    keep_tv = batched_nms(boxes, scores, torch.zeros(300), 0.7)
    keep_or = sam_amg.nms(boxes, scores, 0.7)
    assert torch.equal(keep_tv, keep_or), "oracle NMS differs from torchvision batched_nms"
    print(f"[sam_amg] NMS vs torchvision: {len(keep_or)} of 300 boxes kept, identical order")
    keep = min(6, data["masks"].shape[0])
    torch.save({"seed": seed, "emb_seed": seed + 1, "weights_checksum": checksum(sd), "emb": emb.to(torch.bfloat16),
                "points": pts, "low_res": low_ref[:2].to(torch.float16), "iou": iou_ref, "runs": runs,
                "mask_rows_sum": ref_masks.sum(-1).to(torch.int16)[:keep], "soft_top": soft[:keep].to(torch.float16),
                "soft_order": order, "nms_boxes": boxes, "nms_scores": scores, "nms_keep": keep_tv}, GOLD / "sam_amg.pt")


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    ref_ie, ref_tr, ref_loss = _ref_imports()
    gold_relpos(ref_ie)
    gold_sam(ref_ie, "tiny", sam_encoder.SamConfig(img_size=112, embed_dim=64, depth=3, num_heads=2, out_chans=32,
                                                    window_size=3, global_attn_indexes=(1,)), seed=41, store_weights=True)
    gold_sam(ref_ie, "geom", sam_encoder.SamConfig(img_size=1024, embed_dim=32, depth=2, num_heads=2, out_chans=16,
                                                    window_size=14, global_attn_indexes=(1,)), seed=42, store_weights=False)
    gold_selector(ref_tr)
    gold_losses(ref_loss)
    gold_clip()
    gold_llama()
    gold_dinov2()
    gold_splice()
    gold_sam_amg()
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
