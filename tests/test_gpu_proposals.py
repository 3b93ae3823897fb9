"""GPU (-m gpu): SAM-Everything proposal generation through the C ABI against oracle/sam_amg.py, which is
pinned to the reference's own PromptEncoder / MaskDecoder / SamAutomaticMaskGenerator by tests/golden/sam_amg.pt.

The mask decoder is bf16 tensor-core arithmetic (tolerances below); everything after it — up-sampling, thresholds, counts,
boxes, NMS, top-k, the antialiased resize — is compare-and-count work on whatever logits it is given and is checked for
EXACT agreement with the oracle on identical logits."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _blobs(n, seed):
    """n low-res logit maps [256,256]: a few Gaussian bumps each on a negative floor, amplitudes +-20 like a trained
    decoder's, so that masks have holes, several components, boxes of every size and edge contacts."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    yy = torch.arange(256, device=DEV).view(1, 256, 1).float()
    xx = torch.arange(256, device=DEV).view(1, 1, 256).float()
    out = torch.full((n, 256, 256), -6.0, device=DEV)
    for _ in range(4):
        cy = torch.rand(n, 1, 1, generator=g, device=DEV) * 300 - 22
        cx = torch.rand(n, 1, 1, generator=g, device=DEV) * 300 - 22
        sg = torch.rand(n, 1, 1, generator=g, device=DEV) * 40 + 3
        amp = torch.rand(n, 1, 1, generator=g, device=DEV) * 30 - 6
        out = out + amp * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg ** 2))
    out[0] = -5.0                                    # an empty mask
    out[1] = 5.0                                     # a full one
    return (out + 0.3 * torch.randn(n, 256, 256, generator=g, device=DEV)).contiguous()


def _mha_ref(q, k, v):
    """fp32 softmax attention over 8 heads of 16 (reference/model/segment_anything/modeling/transformer.py:222-236)."""
    P = q.shape[0]
    qh, kh, vh = (t.float().view(P, -1, 8, 16).transpose(1, 2) for t in (q, k, v))
    a = torch.softmax(qh @ kh.transpose(-1, -2) * 0.25, -1)
    return (a @ vh).transpose(1, 2).reshape(P, -1, 128)


@pytest.mark.parametrize("kernel", ["mma", "v1"])
@pytest.mark.parametrize("shared_kv", [True, False])
def test_tok2img_attention_vs_torch(cuda_lib, kernel, shared_kv, monkeypatch):
    """7 prompt tokens x 4096 image keys, 8 heads of 16: the tensor-core kernel (mma.sync, warp per head, cp.async ring)
    and the CUDA-core one it replaced (LLMSEG_T2I_V1=1), K / V shared by all prompts (layer 0) or per prompt, read as column
    views of a wider buffer like the decoder does.  Scores get a few large entries so the online max moves."""
    from llmseg_b200 import ops
    P = 5
    g = torch.Generator(device=DEV).manual_seed(3 + shared_kv)
    q = (torch.randn(P * 7, 128, generator=g, device=DEV) * 2).bfloat16()
    nk = 4096 if shared_kv else P * 4096
    kv = torch.randn(nk, 384, generator=g, device=DEV).bfloat16()
    kv[torch.randint(0, nk, (64,), generator=g, device=DEV), :128] *= 4
    k, v = kv[:, :128], kv[:, 256:]
    if kernel == "v1":
        monkeypatch.setenv("LLMSEG_T2I_V1", "1")
    out = ops.tok2img_attention(q, k, v, P, shared_kv)
    kk = (k if not shared_kv else k.unsqueeze(0).expand(P, -1, -1)).reshape(P, 4096, 128)
    vv = (v if not shared_kv else v.unsqueeze(0).expand(P, -1, -1)).reshape(P, 4096, 128)
    ref = _mha_ref(q.view(P, 7, 128), kk, vv).reshape(P * 7, 128)
    d = (out.float() - ref).abs()
    # P is rounded to bf16 for the second MMA (2^-9 relative on each of ~tens of effective terms) and the output is bf16
    assert d.max().item() <= 2 ** -7 * float(ref.abs().max()) and d.mean().item() <= 2e-3 * float(ref.abs().mean()) + 1e-4


@pytest.mark.parametrize("shared_q", [True, False])
def test_img2tok_attention_vs_torch(cuda_lib, shared_q):
    """4096 image queries x 7 token keys per prompt (warp per head, K / V broadcast from shared memory)."""
    from llmseg_b200 import ops
    P = 3
    g = torch.Generator(device=DEV).manual_seed(11 + shared_q)
    nq = 4096 if shared_q else P * 4096
    qb = (torch.randn(nq, 256, generator=g, device=DEV) * 2).bfloat16()
    q = qb[:, 128:]
    k = torch.randn(P * 7, 128, generator=g, device=DEV).bfloat16()
    v = torch.randn(P * 7, 128, generator=g, device=DEV).bfloat16()
    out = ops.img2tok_attention(q, k, v, P, shared_q)
    qq = (q if not shared_q else q.unsqueeze(0).expand(P, -1, -1)).reshape(P, 4096, 128)
    ref = _mha_ref(qq, k.view(P, 7, 128), v.view(P, 7, 128)).reshape(P * 4096, 128)
    assert (out.float() - ref).abs().max().item() <= 2 ** -8 * float(ref.abs().max()) + 1e-3


def test_upscale_logits_fused_vs_unfused_and_fp32(cuda_lib):
    """`upscale_logits` (LayerNorm2d(64)+GELU -> ConvTranspose #2 + GELU -> hyper-network product, one kernel, ConvTranspose
    #2 activations kept in fp32) against the three launches it replaces (ln64_gelu + gemm + mask_logits, which round them
    to bf16) and against the same arithmetic in fp32 torch (mask_decoder.py:56-64,139-157 on un-shuffled rows): the fused
    kernel must not be further from fp32 than the un-fused path."""
    from llmseg_b200 import ops
    P = 3
    g = torch.Generator(device=DEV).manual_seed(5)
    rnd = lambda *sh: torch.randn(*sh, generator=g, device=DEV)
    u1 = (rnd(P * 4096, 256) * 1.5 + 0.2).bfloat16()
    gamma, beta = (1 + 0.2 * rnd(64)).bfloat16(), (0.1 * rnd(64)).bfloat16()
    w2, b2 = (rnd(128, 64) / 8).bfloat16(), (0.1 * rnd(32)).repeat(4).bfloat16().contiguous()
    hyper = rnd(P, 4, 32).bfloat16()
    fused = ops.upscale_logits(u1, gamma, beta, w2, b2, hyper, P, 1e-6)
    x = u1.clone()
    ops.ln64_gelu(x, gamma, beta, 1e-6)
    unfused = ops.mask_logits(ops.gemm(x.view(P * 16384, 64), w2, b2, act="gelu"), hyper, P)
    xf = torch.nn.functional.layer_norm(u1.float().view(-1, 64), (64,), gamma.float(), beta.float(), 1e-6)
    u2 = torch.nn.functional.gelu(torch.nn.functional.gelu(xf) @ w2.float().T + b2.float())          # [P*16384, 128]
    # rows (p, ty, tx, dy, dx), cols (dy2, dx2, c)  ->  pixel (4 ty + 2 dy + dy2, 4 tx + 2 dx + dx2)
    u2 = u2.view(P, 64, 64, 2, 2, 2, 2, 32)
    ref = torch.einsum("pyxabcdk,pmk->pmyacxbd", u2, hyper.float()[:, 1:4]).reshape(P, 3, 256, 256)
    scale = float(ref.abs().max())
    e_f, e_u = (fused - ref).abs(), (unfused - ref).abs()
    print(f"upscale_logits: |ref| max {scale:.2f}; fused-fp32 max {e_f.max():.2e} mean {e_f.mean():.2e}; "
          f"unfused-fp32 max {e_u.max():.2e} mean {e_u.mean():.2e}")
    assert e_f.mean().item() <= 1.05 * e_u.mean().item() + 1e-6
    assert e_f.max().item() <= 2 ** -7 * scale


def test_mask_stats_boxes_soft_binarize_vs_oracle(cuda_lib):
    from llmseg_b200 import ops
    from oracle import sam_amg
    low = _blobs(24, 3)
    up = sam_amg.upsample_logits(low[:, None])[:, 0]                      # [24,1024,1024] fp32 (ATen bilinear)
    stats = ops.mask_stats(low, None, 0.0, 1.0)
    ref_area = (up > 0).flatten(1).sum(-1)
    ref_hi = (up > 1.0).flatten(1).sum(-1)
    ref_lo = (up > -1.0).flatten(1).sum(-1)
    ref_box = sam_amg.mask_to_box(up > 0)
    # the up-sampling is evaluated with ATen's formula; a pixel whose logit lands within an fp32 ulp of a threshold may
    # still fall on the other side (fused multiply-adds): allow 2 pixels of 1 M per count, boxes must agree exactly
    for got, ref in ((stats[:, 0], ref_area), (stats[:, 1], ref_hi), (stats[:, 2], ref_lo)):
        assert (got.long() - ref).abs().max().item() <= 2, (got.tolist(), ref.tolist())
    box = torch.stack([1023 - stats[:, 3], 1023 - stats[:, 4], stats[:, 5], stats[:, 6]], dim=1).long()
    box[stats[:, 0] == 0] = 0
    assert torch.equal(box, ref_box), (box.tolist(), ref_box.tolist())
    assert int(stats[0, 0]) == 0 and int(stats[1, 0]) == 1024 * 1024
    # candidate indirection + binary masks
    cand = torch.tensor([5, 1, 17, 0, 9], dtype=torch.int32, device=DEV)
    assert torch.equal(ops.mask_stats(low, cand, 0.0, 1.0), stats[cand.long()])
    masks = ops.mask_binarize(low, cand, 0.0)
    diff = (masks.bool() != (up[cand.long()] > 0)).flatten(1).sum(-1)
    assert diff.max().item() <= 2
    # soft proposals: antialiased bilinear 1024 -> 256 of the binary mask (reference utils/dataset.py:620-622)
    soft = ops.mask_soft(low, cand, 0.0)
    ref_soft = torch.nn.functional.interpolate(masks.float()[None], size=(256, 256), mode="bilinear", align_corners=False,
                                               antialias=True)[0]
    assert soft.dtype == torch.bfloat16 and soft.shape == (5, 256, 256)
    assert (soft.float() - ref_soft).abs().max().item() <= 2 ** -8          # one bf16 rounding of values in [0, 1]
    assert (soft.float() - ref_soft.to(torch.bfloat16).float()).abs().max().item() <= 2 ** -7
    assert float(soft[3].float().abs().max()) == 0.0 and float(soft[1].float().min()) == 1.0


def test_box_nms_vs_torchvision_golden(cuda_lib, golden_dir):
    from llmseg_b200 import ops
    from oracle import sam_amg
    fx = torch.load(golden_dir / "sam_amg.pt", weights_only=False)
    boxes, scores, keep_ref = fx["nms_boxes"], fx["nms_scores"], fx["nms_keep"]
    order = torch.argsort(scores, descending=True, stable=True)
    keep = ops.box_nms(boxes[order].to(DEV).contiguous(), 0.7).cpu().bool()
    assert torch.equal(order[keep], keep_ref)                              # torchvision.ops.batched_nms, same order
    assert torch.equal(order[keep], sam_amg.nms(boxes, scores, 0.7))
    assert int(ops.box_nms(boxes[order].to(DEV).contiguous(), 1.0).sum()) == 300      # IoU > 1 never: nothing suppressed


def _generator(seed=11):
    from llmseg_b200 import proposals
    from oracle import sam_amg
    sd = sam_amg.random_state_dict(seed)
    pre = {"model.visual_model." + k: v for k, v in sd.items()}
    gen = proposals.SamProposalGenerator(pre, DEV)
    # the oracle runs on the bf16-rounded weights the kernels use (the Gaussian matrix stays fp32 on both sides)
    osd = {k: (v if "gaussian" in k else v.to(torch.bfloat16).float()).to(DEV) for k, v in sd.items()}
    return gen, osd


def test_mask_decoder_vs_oracle(cuda_lib, golden_dir):
    """Prompt encoder + mask decoder on the golden image embedding: low-res mask logits and IoU predictions for point
    prompts against the fp32 oracle (itself equal to the reference classes on these inputs, tests/golden/sam_amg.pt).
    Tolerance: bf16 activations through 2 two-way blocks + 2 up-scaling steps; logits span +-35."""
    from oracle import sam_amg
    gen, osd = _generator()
    fx = torch.load(golden_dir / "sam_amg.pt", weights_only=False)
    emb = fx["emb"].to(DEV)                                                # bf16 [1,256,64,64]
    tok = emb[0].permute(1, 2, 0).reshape(4096, 256).contiguous()
    pts = torch.cat([fx["points"], torch.from_numpy(__import__("llmseg_b200.proposals", fromlist=["x"]).point_grid(4))]).to(DEV)
    with torch.no_grad():
        low, iou = gen.decode(gen.image_keys(tok), pts)
        ref_low, ref_iou = sam_amg.predict_points(emb.float(), pts, osd)
    scale = ref_low.abs().max().item()
    d = (low - ref_low).abs()
    print(f"mask decoder: |logit| max {scale:.1f}; max|d| {d.max().item():.3f} mean|d| {d.mean().item():.4f}; "
          f"iou max|d| {(iou - ref_iou).abs().max().item():.4f}; sign agreement {((low > 0) == (ref_low > 0)).float().mean().item():.5f}")
    assert low.shape == (21, 3, 256, 256) and iou.shape == (21, 3)
    assert d.max().item() <= 0.04 * scale and d.mean().item() <= 0.004 * scale
    assert (iou - ref_iou).abs().max().item() <= 3e-2
    assert ((low > 0) == (ref_low > 0)).float().mean().item() > 0.995
    # the golden's own low-res logits (reference MaskDecoder, fp32 weights) for the first two prompts
    assert (low[:2] - fx["low_res"].to(DEV).float()).abs().max().item() <= 0.05 * scale


def test_generate_proposals_vs_oracle_on_identical_logits(cuda_lib, golden_dir):
    """The generator end to end (8 x 8 and 32 x 32 point grids) — and, on the logits it produced, every step after the
    decoder against the oracle: same surviving candidates in the same order, same boxes / areas / stability, soft masks
    equal up to the bf16 rounding, for the default NMS threshold and with suppression off."""
    from oracle import sam_amg
    gen, osd = _generator()
    fx = torch.load(golden_dir / "sam_amg.pt", weights_only=False)
    tok = fx["emb"].to(DEV)[0].permute(1, 2, 0).reshape(4096, 256).contiguous()
    from llmseg_b200 import proposals
    for pps, nms_thr, ppb in ((8, 1.0, 16), (8, 0.7, 64), (32, 0.9, 256)):
        kw = dict(points_per_side=pps, pred_iou_thresh=-0.6 if pps == 8 else 0.2, stability_score_thresh=0.5,
                  stability_score_offset=1.0, box_nms_thresh=nms_thr)
        P = pps * pps
        pts = torch.from_numpy(proposals.point_grid(pps)).to(DEV)
        with torch.no_grad():
            img = gen.image_keys(tok)
            low = torch.empty((P, 3, 256, 256), device=DEV)
            ious = []
            for i in range(0, P, ppb):
                _, iou = gen.decode(img, pts[i:i + ppb], low_out=low[i:i + ppb])
                ious.append(iou)
            iou = torch.cat(ious)
            out = gen.generate(tok, points_per_batch=ppb, top_k=50, return_masks=True, **kw)
            again = gen.generate(None, low_res=low, iou_preds=iou, top_k=50, **kw)
            ref = sam_amg.generate(None, osd, points_per_batch=64, decoded=(low, iou), **kw)
            soft_ref, order = sam_amg.llmseg_proposals(ref, top_k=50)
        K = order.numel()
        print(f"grid {pps}x{pps}, nms {nms_thr}: {ref['masks'].shape[0]} masks after NMS, {K} proposals")
        assert out["n_masks"] == ref["masks"].shape[0] and out["segs"].shape == (K, 256, 256) and K >= 1
        assert torch.equal(out["candidates"], again["candidates"])          # batching of the decoder does not matter
        assert torch.equal(out["candidates"].to(DEV), ref["candidates"][order])
        assert torch.equal(out["boxes"].to(DEV), ref["boxes"][order])
        assert (out["areas"].to(DEV) - ref["areas"][order]).abs().max().item() <= 2
        assert (out["stability"].to(DEV) - ref["stability"][order]).abs().max().item() <= 1e-5
        assert (out["points"].to(DEV) - ref["points"][order]).abs().max().item() <= 1e-3
        assert (out["segs"].float() - soft_ref).abs().max().item() <= 2 ** -8
        assert ((out["masks"].bool() != ref["masks"][order]).flatten(1).sum(-1)).max().item() <= 2


def test_proposals_feed_the_forward(cuda_lib):
    """Pixels in, selection out: SAM ViT-H features -> SAM-Everything proposals -> `model_forward(sam_segs_list=...)`
    on the same encoder features (reduced depth; the decoder weights ride in the same reference-named state dict)."""
    from llmseg_b200 import lisa, synthetic
    from oracle import sam_amg
    cfg = lisa.LisaCfg()
    cfg.sam.depth, cfg.sam.global_attn_indexes, cfg.clip.layers, cfg.llama.layers = 2, (1,), 2, 1
    sd = synthetic.lisa_state_dict(cfg, seed=0, device=DEV)
    sd.update({"model.visual_model." + k: v.to(DEV) for k, v in sam_amg.random_state_dict(5).items()})
    model = lisa.LISAForCausalLM(sd, cfg, device=DEV)
    inp = synthetic.make_inputs(cfg, 2, 8, 16, device=DEV)
    props = model.generate_proposals(inp["images"], points_per_side=8, pred_iou_thresh=0.0, stability_score_thresh=0.3,
                                     box_nms_thresh=0.9, top_k=50)
    assert len(props) == 2 and all(p["segs"].dtype == torch.bfloat16 and p["segs"].shape[0] >= 1 for p in props)
    with torch.no_grad():
        out = model.forward(**dict(inp, sam_segs_list=[p["segs"] for p in props]))
    for b in range(2):
        K = props[b]["segs"].shape[0]
        assert out["pred_similarity"][b].shape == (1, K) and 0 <= int(out["best_index"][b]) < K
        assert torch.isfinite(out["pred_similarity"][b].float()).all()
