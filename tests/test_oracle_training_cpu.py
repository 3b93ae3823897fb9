"""CPU: the oracle's training forward (oracle/lisa_forward.model_forward_training, restating reference
LISA.py:292-313,416-474 + llava_llama.py:83-118) on a tiny configuration: index arithmetic of the label
splice and of the per-image [SEG] grouping, checked against independent straight-line computations."""
import pytest
import torch
import torch.nn.functional as F

from oracle import clip_llama, dinov2, lisa_forward as lf, sam_encoder, selector


def _tiny():
    cfg = lf.LisaConfig(
        sam=sam_encoder.SamConfig(img_size=112, embed_dim=64, depth=2, num_heads=2, out_chans=256, window_size=3,
                                  global_attn_indexes=(1,)),
        clip=clip_llama.ClipConfig(image_size=56, patch_size=14, hidden=64, layers=3, heads=4, mlp=128),
        llama=clip_llama.LlamaConfig(hidden=64, layers=2, heads=4, mlp=176, vocab=100), seg_token_idx=90)
    sd = {}
    sd.update(sam_encoder.random_state_dict(cfg.sam, 1, prefix="model.visual_model.image_encoder."))
    sd.update(clip_llama.clip_random_state_dict(cfg.clip, 2, prefix="model.vision_tower.vision_tower.vision_model."))
    sd.update(clip_llama.llama_random_state_dict(cfg.llama, 3, prefix="model."))
    sd.update({"model." + k: v for k, v in selector.random_state_dict(4, hidden=64).items()})
    g = torch.Generator().manual_seed(5)
    sd["model.mm_projector.weight"] = torch.randn(64, 64, generator=g) * 0.125
    sd["model.mm_projector.bias"] = torch.randn(64, generator=g) * 0.02
    sd["lm_head.weight"] = torch.randn(100, 64, generator=g) * 0.125
    return cfg, sd, g


def _inputs(cfg, g, convs=(2, 1), Ks=(6, 4), Tt=12):
    N, B = sum(convs), len(convs)
    ids = torch.randint(3, 80, (N, Tt), generator=g)
    ids[:, 0], ids[:, 1], ids[:, 2], ids[:, 3] = 1, 91, lf.IMAGE_TOKEN_INDEX, 92
    labels = torch.full_like(ids, lf.IGNORE_INDEX)
    mask = torch.ones(N, Tt, dtype=torch.bool)
    for n in range(N):
        end = Tt - n
        ids[n, end - 3] = cfg.seg_token_idx
        ids[n, end:] = 0
        mask[n, end:] = False
        labels[n, end - 5:end] = ids[n, end - 5:end]
    off = torch.tensor([0] + list(torch.tensor(convs).cumsum(0)))
    return dict(
        images=torch.randn(B, 3, 112, 112, generator=g), images_clip=torch.randn(B, 3, 56, 56, generator=g),
        input_ids=ids, labels=labels, attention_masks=mask, offset=off,
        sam_segs_list=[torch.rand(k, 256, 256, generator=g) for k in Ks],
        sam_ious_list=[torch.rand(c, k, generator=g) for c, k in zip(convs, Ks)],
        sam_iops_list=[torch.rand(c, k, generator=g) for c, k in zip(convs, Ks)])


def test_splice_labels_layout():
    ids = torch.tensor([[1, 91, -200, 92, 5, 6, 90, 9]])
    labels = torch.tensor([[-100, -100, -100, -100, 5, 6, 90, 9]])
    out = lf.splice_labels(ids, labels, 4)
    assert out.tolist() == [[-100, -100, -100, -100, -100, -100, -100, 5, 6, 90, 9]]


def test_training_forward_tiny():
    cfg, sd, g = _tiny()
    inp = _inputs(cfg, g)
    with torch.no_grad():
        out = lf.model_forward_training(sd, cfg, **inp, ce_loss_weight=1.0, align_loss_weight=2.0,
                                        regression_loss_weight=0.5)
        assert all(torch.isfinite(torch.as_tensor(out[k])) for k in out)
        assert abs(float(out["loss"]) - float(out["ce_loss"] + out["align_loss"] + out["regression_loss"])) < 1e-6
        # independent route: per conversation single-image inference pieces, then the documented reductions
        n_img = cfg.n_image_tokens
        feats_all, hid = [], []
        conv_img = [0, 0, 1]
        ce_num, ce_den = 0.0, 0
        sel_sd = lf.sub_dict(sd, "model.")
        for n in range(3):
            b = conv_img[n]
            feats = lf.encode_images(inp["images_clip"][b:b + 1], sd, cfg)
            e, m = lf.splice_inputs(inp["input_ids"][n:n + 1], inp["attention_masks"][n:n + 1], feats,
                                    sd["model.embed_tokens.weight"])
            h = clip_llama.llama_last_hidden(e, m, sel_sd, cfg.llama)[0]
            logp = F.log_softmax(F.linear(h, sd["lm_head.weight"]), dim=-1)
            ids, lab = inp["input_ids"][n], inp["labels"][n]
            for j in range(4, ids.shape[0]):            # text index j (> image index 2) sits at spliced j + n_img - 1
                if lab[j] != lf.IGNORE_INDEX:
                    ce_num -= float(logp[j + n_img - 2, lab[j]])
                    ce_den += 1
            s = int((ids == cfg.seg_token_idx).nonzero()[0])
            hid.append(selector.text_hidden_fc(h[s + n_img - 2][None], sel_sd))
        assert abs(ce_num / ce_den - float(out["ce_loss"])) < 1e-4
        emb_up = selector.upsample_embeddings(lf.image_features(sd, cfg, inp["images"]))
        al, rl = [], []
        for n in range(3):
            b, r = conv_img[n], (n if n < 2 else 0)
            f, u = selector.selector_features(emb_up[b], inp["sam_segs_list"][b], hid[n], sel_sd)
            al.append(lf.softmax_align_loss(f[0], hid[n], inp["sam_ious_list"][b][r][:, None]))
            rl.append(lf.iou_regression_loss(u[0], inp["sam_iops_list"][b][r][:, None]))
        a_ref = 2.0 * (0.5 * (al[0] + al[1]) + al[2]) / 2
        r_ref = 0.5 * (0.5 * (rl[0] + rl[1]) + rl[2]) / 2
        assert abs(float(a_ref) - float(out["align_loss"])) < 1e-4 * (1 + abs(float(a_ref)))
        assert abs(float(r_ref) - float(out["regression_loss"])) < 1e-4 * (1 + abs(float(r_ref)))


def test_training_forward_zero_rounds_raises():
    cfg, sd, g = _tiny()
    inp = _inputs(cfg, g)
    inp["input_ids"][2][inp["input_ids"][2] == cfg.seg_token_idx] = 5
    with pytest.raises(ValueError), torch.no_grad():
        lf.model_forward_training(sd, cfg, **inp)
