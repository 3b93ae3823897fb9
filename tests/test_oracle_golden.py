"""CPU: the oracle restatements reproduce the golden vectors produced by the reference's own
modules (oracle/make_golden.py).  This is what pins the oracle (the reference has no tests)."""
import torch

from oracle import clip_llama, dinov2, lisa_forward, sam_encoder, selector


def _load(golden_dir, name):
    return torch.load(golden_dir / name, weights_only=False)


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def test_relpos_and_partition(golden_dir):
    fx = _load(golden_dir, "sam_relpos.pt")
    for S in (14, 5):
        c = fx[f"S{S}"]
        mine = sam_encoder.decomposed_rel_pos_bias(c["q"], c["rel_h"], c["rel_w"], (S, S))
        assert torch.allclose(mine, c["bias"], atol=1e-5)
    x = fx["partition"]["x"]
    w, pad = sam_encoder.partition_windows(x, 3)
    assert torch.equal(w, fx["partition"]["windows"])
    assert torch.equal(sam_encoder.unpartition_windows(w, 3, pad, (7, 7)), x)
    # padding tokens are zeros (they are real keys in attention)
    assert float(w[2, :, 1:].abs().sum() + w[-1, 1:].abs().sum()) == 0.0


def test_sam_tiny(golden_dir):
    fx = _load(golden_dir, "sam_tiny.pt")
    cfg = sam_encoder.SamConfig(**fx["cfg"])
    out = sam_encoder.image_encoder(fx["x"], fx["sd"], cfg)
    assert torch.allclose(out, fx["out"], atol=2e-4)


def test_sam_production_geometry(golden_dir):
    """1024 px / 64x64 tokens / 14x14 windows with the 64->70 padding, small width; weights by seed."""
    fx = _load(golden_dir, "sam_geom.pt")
    cfg = sam_encoder.SamConfig(**fx["cfg"])
    sd = sam_encoder.random_state_dict(cfg, fx["seed"])
    assert abs(_checksum(sd) - fx["weights_checksum"]) < 1e-6 * fx["weights_checksum"], "RNG drift: regenerate goldens"
    x = torch.randn(1, 3, cfg.img_size, cfg.img_size, generator=torch.Generator().manual_seed(fx["x_seed"]))
    out = sam_encoder.image_encoder(x, sd, cfg)
    assert torch.allclose(out, fx["out"], atol=2e-4)


def test_selector(golden_dir):
    for K in (32, 64, 7):
        fx = _load(golden_dir, f"selector_K{K}.pt")
        sd = selector.random_state_dict(fx["seed"], hidden=fx["hidden_dim"])
        assert abs(_checksum(sd) - fx["weights_checksum"]) < 1e-6 * fx["weights_checksum"]
        emb, segs, hidden = selector.synthetic_case(fx["seed"], K, fx["hidden_dim"])
        text = selector.text_hidden_fc(hidden, sd)
        up = selector.upsample_embeddings(emb)[0]
        assert torch.allclose(selector.mask_pooling(up, segs), fx["feat"], atol=1e-5)
        sim, iou = selector.selector_forward(up, segs, text, sd)
        assert sim.shape == (1, K) and iou.shape == (1, K)
        assert torch.allclose(sim, fx["pred_similarity"], atol=1e-5)
        assert torch.allclose(iou, fx["pred_iou"], atol=1e-5)
        idx, keep = selector.select(sim, iou)
        assert idx == int(fx["pred_similarity"].argmax())


def test_losses(golden_dir):
    fx = _load(golden_dir, "losses.pt")
    e = fx["expected"]
    assert abs(float(lisa_forward.softmax_align_loss(fx["pe"], fx["te"], fx["gt_ious"])) - e["softmax_align"]) < 1e-5
    assert abs(float(lisa_forward.iou_regression_loss(fx["pred_ious"], fx["gt_ious"])) - e["iou_regression"]) < 1e-5
    assert abs(float(lisa_forward.dice_loss(fx["logits"], fx["targets"], 3.0)) - e["dice"]) < 1e-5
    assert abs(float(lisa_forward.sigmoid_ce_loss(fx["logits"], fx["targets"], 3.0)) - e["sigmoid_ce"]) < 1e-5


def test_clip_tiny(golden_dir):
    fx = _load(golden_dir, "clip_tiny.pt")
    cfg = clip_llama.ClipConfig(**fx["cfg"])
    out = clip_llama.clip_patch_features(fx["x"], fx["sd"], cfg)
    assert out.shape == (2, cfg.tokens - 1, cfg.hidden)
    assert torch.allclose(out, fx["out"], atol=1e-4)


def test_llama_tiny(golden_dir):
    fx = _load(golden_dir, "llama_tiny.pt")
    cfg = clip_llama.LlamaConfig(**fx["cfg"])
    out = clip_llama.llama_last_hidden(fx["embeds"], fx["mask"], fx["sd"], cfg)
    valid = fx["mask"][:, :, None]
    assert torch.allclose(out * valid, fx["out"] * valid, atol=1e-4)


def test_splice_and_seg_mask():
    """SURVEY §A.6 worked example."""
    cfg = lisa_forward.LisaConfig()
    ids = torch.tensor([[1, 32001, -200, 32002, 5, 6, 7, 32000, 9, 2]])
    m = lisa_forward.seg_token_mask(ids, cfg)
    assert m.shape == (1, 265) and m[0].nonzero().flatten().tolist() == [7 + 254]
    table = torch.arange(40000, dtype=torch.float32)[:, None].repeat(1, 2)
    feats = -torch.arange(1, 257, dtype=torch.float32)[None, :, None].repeat(1, 1, 2)
    e, am = lisa_forward.splice_inputs(ids, torch.ones(1, 10, dtype=torch.bool), feats, table)
    assert e.shape == (1, 265, 2) and bool(am.all())
    assert e[0, 2, 0] == -1 and e[0, 257, 0] == -256 and e[0, 258, 0] == 32002 and e[0, 7 + 254, 0] == 7


def test_end_to_end_tiny_oracle_runs():
    """The glue restatement (LISA.py:225-414) executes on a tiny configuration and is deterministic."""
    cfg = lisa_forward.LisaConfig(
        sam=sam_encoder.SamConfig(img_size=1024, embed_dim=32, depth=2, num_heads=2, out_chans=256,
                                  window_size=14, global_attn_indexes=(1,)),
        clip=clip_llama.ClipConfig(image_size=224, patch_size=14, hidden=32, layers=3, heads=2, mlp=64),
        llama=clip_llama.LlamaConfig(hidden=64, layers=2, heads=2, mlp=96, vocab=32003))
    sd = {}
    sd.update({"model.visual_model.image_encoder." + k: v for k, v in sam_encoder.random_state_dict(cfg.sam, 1).items()})
    sd.update({"model.vision_tower.vision_tower." + k: v for k, v in clip_llama.clip_random_state_dict(cfg.clip, 2).items()})
    sd.update({"model." + k: v for k, v in clip_llama.llama_random_state_dict(cfg.llama, 3).items()})
    sd.update({"model." + k: v for k, v in selector.random_state_dict(4, hidden=64).items()})
    g = torch.Generator().manual_seed(5)
    sd["model.mm_projector.weight"] = torch.randn(64, 32, generator=g) * 0.1
    sd["model.mm_projector.bias"] = torch.zeros(64)
    K, T = 5, 12
    ids = torch.randint(3, 31999, (1, T), generator=g)
    ids[0, 0], ids[0, 1], ids[0, 2], ids[0, 3], ids[0, T - 3] = 1, 32001, -200, 32002, 32000
    inp = dict(images=torch.randn(1, 3, 1024, 1024, generator=g), images_clip=torch.randn(1, 3, 224, 224, generator=g),
               input_ids=ids, attention_masks=torch.ones(1, T, dtype=torch.bool), offset=torch.tensor([0, 1]),
               sam_segs_list=[torch.rand(K, 256, 256, generator=g)])
    out = lisa_forward.model_forward_inference(sd, cfg, **inp)
    assert out["pred_similarity"][0].shape == (1, K) and out["pred_iou"][0].shape == (1, K)
    assert float(out["pred_similarity"][0].abs().max()) <= 1.0 + 1e-5
    out2 = lisa_forward.forward_batched(sd, cfg, images=inp["images"], images_clip=inp["images_clip"], input_ids=ids,
                                        attention_masks=inp["attention_masks"], sam_segs_list=inp["sam_segs_list"])
    assert torch.equal(out2["pred_similarity"][0], out["pred_similarity"][0])


def test_dinov2_tiny(golden_dir):
    """Variant-B image encoder: hub-named weights, position table resampled 5x5 -> 9x9; golden tokens come from
    transformers' Dinov2Model, golden embeddings add the reference's reshape + 1x1 lisa_dino_conv."""
    fx = _load(golden_dir, "dinov2_tiny.pt")
    cfg = dinov2.Dinov2Config(**fx["cfg"])
    tok = dinov2.forward_features(fx["x"], fx["sd"], cfg)
    assert tok.shape == (2, 81, 64) and torch.allclose(tok, fx["tokens"], atol=1e-5)
    emb = dinov2.image_embeddings(fx["x"], fx["sd"], fx["conv_w"], fx["conv_b"], cfg)
    assert emb.shape == (2, 16, 9, 9) and torch.allclose(emb, fx["embeddings"], atol=1e-5)
    # hub default resampling (interpolate_offset=0.1) is a different table from the size= one
    pos = dinov2.interpolate_pos_embed(fx["sd"]["pos_embed"], cfg.grid, 0.1)
    assert torch.allclose(pos, fx["pos_offset01"], atol=1e-6)
    assert (pos - dinov2.interpolate_pos_embed(fx["sd"]["pos_embed"], cfg.grid, 0.0)).abs().max() > 1e-4
    # a table already stored at the evaluation grid is used as is
    same = dinov2.interpolate_pos_embed(fx["sd"]["pos_embed"], cfg.train_grid, 0.1)
    assert same is fx["sd"]["pos_embed"]


def test_host_pos_embed_resampling_matches_oracle(golden_dir):
    """The product's load-time resampling (llmseg_b200.encoders) against the oracle's, both offsets."""
    from llmseg_b200.encoders import _resample_pos_embed
    fx = _load(golden_dir, "dinov2_tiny.pt")
    pe = fx["sd"]["pos_embed"]
    for off in (0.1, 0.0):
        mine = _resample_pos_embed(pe, 9, off)
        assert torch.allclose(mine, dinov2.interpolate_pos_embed(pe, 9, off)[0], atol=1e-6)
    assert torch.equal(_resample_pos_embed(pe, 5, 0.1), pe[0])


def test_sam_amg_oracle_matches_reference_golden(golden_dir):
    """oracle/sam_amg.py against the outputs of the reference's own PromptEncoder / MaskDecoder /
    SamAutomaticMaskGenerator (tests/golden/sam_amg.pt, written by oracle/make_golden.py): decoder logits and IoU
    predictions for 5 point prompts, the generator's records for an 8 x 8 point grid at two NMS thresholds, the soft
    256 x 256 proposals of the largest masks, and the suppression rule against torchvision's batched_nms."""
    from oracle import sam_amg
    fx = torch.load(golden_dir / "sam_amg.pt", weights_only=False)
    sd = sam_amg.random_state_dict(fx["seed"])
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values()) - fx["weights_checksum"]) < 1e-6 * fx["weights_checksum"]
    emb = fx["emb"].float()
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    with torch.no_grad():
        low, iou = sam_amg.predict_points(emb, fx["points"], sd)
    # the fixture stores the embedding in bf16 and the logits in fp16: compare at that resolution
    assert (iou - fx["iou"]).abs().max().item() < 2e-2
    assert (low[:2] - fx["low_res"].float()).abs().max().item() < 0.02 * float(fx["low_res"].float().abs().max())
    assert torch.equal(sam_amg.nms(fx["nms_boxes"], fx["nms_scores"], 0.7), fx["nms_keep"])
    run = fx["runs"]["nms07"]
    with torch.no_grad():
        data = sam_amg.generate(emb, sd, **run["kw"])
    # (bf16 embedding: a borderline candidate may flip; the default-threshold run keeps the single dominant mask)
    assert data["masks"].shape[0] == run["n_masks"]
    assert (data["areas"] - run["ref_area"]).abs().max().item() <= 0.01 * float(run["ref_area"].max())
    soft, order = sam_amg.llmseg_proposals(data, top_k=50)
    assert soft.shape == (run["n_masks"], 256, 256) and float(soft.min()) >= 0.0 and float(soft.max()) <= 1.0 + 1e-6
    lo, w = sam_amg.aa_downsample_weights(1024, 256)
    assert int(lo[0]) == 0 and int(lo[1]) == 2 and int(lo[255]) == 1018 and abs(float(w[7].sum()) - 1.0) < 1e-12
