"""Developer check (GPU; imports oracle/, hence under tests/ — not collected by pytest):
WHERE does the full-depth forward lose accuracy against the fp32 oracle?

For several input seeds, at full depth (SAM ViT-H 32 blocks, CLIP 23 layers, LLaMA-7B 32 layers), three
implementations are run on identical bf16 weights and inputs:

    ours   the sm_100a kernels of this repo
    ref16  the oracle restatement executed as eager bf16 PyTorch on the same GPU (the reference's own path)
    fp32   the oracle in fp32 (the truth)

and the stages are swapped between them: every (image branch, text branch) pair of {ours, ref16, fp32} is fed to
BOTH selectors (ours on bf16 inputs, the fp32 oracle selector), so the error of `pred_similarity` / `pred_iou`
splits into   image-encoder error  +  text-branch error  +  selector error.

    python tests/parity_bisect.py [--seeds 4] [--k 64] [--t-text 64] [--fold-norm image|all|0] > profiles/<name>.txt
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=4)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--t-text", type=int, default=64)
    ap.add_argument("--fold-norm", default="image", choices=["image", "all", "0"],
                    help="where norms are folded into the consuming GEMM (LLMSEG_FOLD_NORM)")
    ap.add_argument("--depth", default="32,24,32", help="SAM blocks, CLIP layers (config), LLaMA layers")
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from llmseg_b200 import encoders, lisa, ops, synthetic
    from oracle import clip_llama as o_cl, lisa_forward as o_lf, sam_encoder as o_sam, selector as o_sel
    encoders.FOLD_NORM_SAM = encoders.FOLD_NORM_IMAGE = args.fold_norm in ("image", "all")
    encoders.FOLD_NORM_TEXT = args.fold_norm == "all"
    dev = "cuda"
    sam_d, clip_l, llama_l = (int(v) for v in args.depth.split(","))
    glob = tuple(i for i in (7, 15, 23, 31) if i < sam_d) or (sam_d - 1,)
    cfg = lisa.LisaCfg()
    cfg.sam.depth, cfg.sam.global_attn_indexes = sam_d, glob
    cfg.clip.layers, cfg.llama.layers = clip_l, llama_l
    ocfg = o_lf.LisaConfig(sam=o_sam.SamConfig(depth=sam_d, global_attn_indexes=glob),
                           clip=o_cl.ClipConfig(layers=clip_l), llama=o_cl.LlamaConfig(layers=llama_l))
    sd = synthetic.lisa_state_dict(cfg, seed=0, device=dev)
    model = lisa.LISAForCausalLM(sd, cfg, device=dev, use_cuda_graph=False)
    fsd = {k: v.float() for k, v in sd.items()}
    print(f"# parity bisect: depth {args.depth}, K={args.k}, T_text={args.t_text}, FOLD_NORM={args.fold_norm}, "
          f"{args.seeds} input seeds, weights seed 0")

    def oracle_parts(s, inp, cast):
        """-> (image embedding [1,256,64,64], text embedding [1,256]) of the oracle in the dtype of `s`."""
        with torch.no_grad():
            img = o_lf.image_features(s, ocfg, cast(inp["images"]))
            feats = o_lf.encode_images(cast(inp["images_clip"]), s, ocfg)
            embeds, mask = o_lf.splice_inputs(inp["input_ids"], inp["attention_masks"], feats, s["model.embed_tokens.weight"])
            hidden = o_cl.llama_last_hidden(embeds, mask, o_lf.sub_dict(s, "model."), ocfg.llama)
            last = o_sel.text_hidden_fc(hidden, o_lf.sub_dict(s, "model."))
            text = last[o_lf.seg_token_mask(inp["input_ids"], ocfg)]
        return img, text

    def ours_parts(inp):
        with torch.no_grad():
            emb = model.image_encoder.forward(inp["images"])                       # [1,4096,256] token-major
            feats = model.clip.forward(inp["images_clip"])
            embeds, kv_len, seg_row = ops.embed_splice(inp["input_ids"], inp["attention_masks"], model.llama.embed, feats,
                                                       image_token=lisa.IMAGE_TOKEN_INDEX, seg_token=model.seg_token_idx)
            T = inp["input_ids"].shape[1] + feats.shape[1] - 1
            hidden = model.llama.forward(embeds, 1, T, kv_len, out_rows=seg_row)
            text = model.selector.text_embed(hidden)
        return emb, text

    def sel_ours(emb_tok, text, segs):
        plan = model.selector.make_plan([segs.shape[0]])
        with torch.no_grad():
            sim, iou, _ = model.selector.forward(emb_tok.to(torch.bfloat16).contiguous(), segs,
                                                 text.to(torch.bfloat16).contiguous(), plan)
        return sim[0].float(), iou[0].float()

    def sel_fp32(img_nchw, text, segs):
        with torch.no_grad():
            up = o_sel.upsample_embeddings(img_nchw.float())
            s, u = o_sel.selector_forward(up[0], segs.float(), text.float(), o_lf.sub_dict(fsd, "model."))
        return s[0], u[0]

    tok = lambda nchw: nchw.permute(0, 2, 3, 1).reshape(1, 4096, 256)
    nchw = lambda t: t.float().reshape(1, 64, 64, 256).permute(0, 3, 1, 2)
    rows = {}

    def rec(name, sim, iou, ref):
        e = rows.setdefault(name, [])
        e.append(((sim - ref[0]).abs().max().item(), (sim - ref[0]).abs().mean().item(),
                  (iou - ref[1]).abs().max().item(), (iou - ref[1]).abs().mean().item()))

    stage = {}

    def rec_stage(name, a, b):
        d = (a.float() - b.float())
        stage.setdefault(name, []).append((d.abs().max().item(), (d.pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt()).item()))

    margins = []
    for si in range(args.seeds):
        inp = synthetic.make_inputs(cfg, 1, args.k, args.t_text, seed=1234 + 17 * si, device=dev)
        segs = inp["sam_segs_list"][0]
        img32, txt32 = oracle_parts(fsd, inp, lambda t: t.float())
        img16, txt16 = oracle_parts(sd, inp, lambda t: t)
        emb_o, txt_o = ours_parts(inp)
        ref = sel_fp32(img32, txt32, segs)
        top2 = ref[0].topk(2).values
        margins.append(float(top2[0] - top2[1]))
        rec_stage("image embedding  ours  vs fp32", nchw(emb_o), img32)
        rec_stage("image embedding  ref16 vs fp32", img16, img32)
        rec_stage("text embedding   ours  vs fp32", txt_o, txt32)
        rec_stage("text embedding   ref16 vs fp32", txt16, txt32)
        src_img = {"ours": nchw(emb_o), "ref16": img16.float(), "fp32": img32}
        src_txt = {"ours": txt_o.float(), "ref16": txt16.float(), "fp32": txt32}
        for ni, im in src_img.items():
            for nt, tx in src_txt.items():
                s, u = sel_ours(tok(im), tx, segs)
                rec(f"image={ni:5s} text={nt:5s} selector=ours", s, u, ref)
                s, u = sel_fp32(im, tx, segs)
                rec(f"image={ni:5s} text={nt:5s} selector=fp32", s, u, ref)
        # the reference's own bf16 selector on its own bf16 stages (== ref16 end to end)
        with torch.no_grad():
            up = o_sel.upsample_embeddings(img16)
            s, u = o_sel.selector_forward(up[0], segs, txt16, o_lf.sub_dict(sd, "model."))
        rec("image=ref16 text=ref16 selector=ref16 (the bf16 reference path)", s[0].float(), u[0].float(), ref)
        with torch.no_grad():
            up = o_sel.upsample_embeddings(img32.to(torch.bfloat16))
            s, u = o_sel.selector_forward(up[0], segs, txt32.to(torch.bfloat16), o_lf.sub_dict(sd, "model."))
        rec("image=fp32  text=fp32  selector=ref16", s[0].float(), u[0].float(), ref)
        with torch.no_grad():
            out = model.forward(**inp)
        rec("model.forward (ours end to end)", out["similarity_padded"][0, :args.k].float(), out["iou_padded"][0, :args.k].float(), ref)
        del img32, img16, emb_o
    print(f"# oracle top-1/top-2 similarity margins per seed: {[round(m, 4) for m in margins]}")
    print("\n## stage outputs (max |d|, relative rms error), mean over seeds")
    for name, v in stage.items():
        t = torch.tensor(v)
        print(f"{name:36s} max|d| {t[:, 0].mean():.4f}   rel rms {t[:, 1].mean():.5f}")
    print("\n## pred_similarity / pred_iou error against the fp32 oracle: max|d| (max over seeds / mean over seeds), mean|d|")
    print(f"{'configuration':66s} {'sim max':>8s} {'sim max~':>8s} {'sim mean':>8s}   {'iou max':>8s} {'iou max~':>8s} {'iou mean':>8s}")
    for name, v in rows.items():
        t = torch.tensor(v)
        print(f"{name:66s} {t[:, 0].max():8.5f} {t[:, 0].mean():8.5f} {t[:, 1].mean():8.5f}   "
              f"{t[:, 2].max():8.5f} {t[:, 2].mean():8.5f} {t[:, 3].mean():8.5f}")


if __name__ == "__main__":
    main()
