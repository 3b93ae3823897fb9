"""ORACLE-side CPU baseline (test/bench infrastructure, never on the product path).

Times the reference algorithm (the oracle restatement: fp32 eager PyTorch, as the reference runs
on a CPU) on the host cores, on a BOUNDED sample of the forward, and extrapolates to images/s:

    one image = patch-embed + 28 windowed + 4 global SAM blocks + neck
              + CLIP embed + 23 CLIP layers + projector
              + 32 LLaMA layers at T = T_text + 255
              + selector (upsample, pooling, 2 two-way blocks, heads, cosine)

Every distinct layer type is executed for real at full width (one instance each, random weights),
timed after one warm-up, and multiplied by its count; the full 32+23+32-layer stack would take
~1 minute per image on 8 cores, which is why the sample is bounded.
"""
from __future__ import annotations

import os
import time
from typing import Dict

import torch
import torch.nn.functional as F

from . import clip_llama, sam_encoder, selector


def _timeit(fn, reps: int = 1) -> float:
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def cpu_forward_sample(t_text: int = 64, k_props: int = 64, threads: int | None = None, reps: int = 1) -> Dict:
    g = torch.Generator().manual_seed(0)
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    if threads is None:
        # "all the host threads it can use": eager PyTorch stops scaling (and regresses) well before
        # 100+ threads, so pick the fastest of {all cores, 64, 32, 16} on one SAM-sized GEMM
        ncpu = os.cpu_count() or 1
        a, b = rn(4096, 1280, std=1.0), rn(5120, 1280, std=1.0)
        best_t, threads = None, ncpu
        for cand in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
            torch.set_num_threads(cand)
            t = _timeit(lambda: a @ b.T, 2)
            if best_t is None or t < best_t:
                best_t, threads = t, cand
    torch.set_num_threads(threads)
    out: Dict = {"cores": threads, "kind": "port"}
    parts = {}
    with torch.no_grad():
        # ---- SAM ViT-H: one windowed block, one global block, patch embed, neck
        scfg = sam_encoder.SamConfig(depth=2, global_attn_indexes=(1,))
        sd = sam_encoder.random_state_dict(scfg, seed=0)
        x = rn(1, 64, 64, 1280, std=1.0)
        img = rn(1, 3, 1024, 1024, std=1.0)
        parts["sam_window_block"] = _timeit(lambda: sam_encoder.block(x, sd, "blocks.0.", scfg, 14), reps)
        parts["sam_global_block"] = _timeit(lambda: sam_encoder.block(x, sd, "blocks.1.", scfg, 0), reps)
        parts["sam_patch_embed"] = _timeit(lambda: F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16), reps)

        def neck():
            y = F.conv2d(x.permute(0, 3, 1, 2), sd["neck.0.weight"])
            y = sam_encoder.layer_norm_2d(y, sd["neck.1.weight"], sd["neck.1.bias"])
            y = F.conv2d(y, sd["neck.2.weight"], padding=1)
            return sam_encoder.layer_norm_2d(y, sd["neck.3.weight"], sd["neck.3.bias"])
        parts["sam_neck"] = _timeit(neck, reps)
        sam_s = parts["sam_patch_embed"] + 28 * parts["sam_window_block"] + 4 * parts["sam_global_block"] + parts["sam_neck"]
        del sd

        # ---- CLIP ViT-L/14: embeddings + one layer (x23) ; projector
        ccfg1 = clip_llama.ClipConfig(layers=1, select_layer=1)
        csd = clip_llama.clip_random_state_dict(ccfg1, seed=1)
        imc = rn(1, 3, 224, 224, std=1.0)
        t_one = _timeit(lambda: clip_llama.clip_patch_features(imc, csd, ccfg1), reps)
        ccfg0 = clip_llama.ClipConfig(layers=1, select_layer=0)
        t_zero = _timeit(lambda: clip_llama.clip_patch_features(imc, csd, ccfg0), reps)
        wproj = rn(4096, 1024)
        feats = rn(1, 256, 1024, std=1.0)
        parts["clip_embed"] = t_zero
        parts["clip_layer"] = max(t_one - t_zero, 0.0)
        parts["mm_projector"] = _timeit(lambda: F.linear(feats, wproj), reps)
        clip_s = parts["clip_embed"] + 23 * parts["clip_layer"] + parts["mm_projector"]
        del csd

        # ---- LLaMA-7B: one decoder layer at T = t_text + 255 (x32)
        lcfg = clip_llama.LlamaConfig(layers=1)
        lsd = clip_llama.llama_random_state_dict(clip_llama.LlamaConfig(layers=1, vocab=8), seed=2)
        T = t_text + 255
        emb = rn(1, T, 4096, std=1.0)
        parts["llama_layer"] = _timeit(lambda: clip_llama.llama_last_hidden(emb, None, lsd, lcfg), reps)
        llama_s = 32 * parts["llama_layer"]
        del lsd

        # ---- selector stage (full)
        ssd = selector.random_state_dict(seed=3, hidden=4096)
        e = rn(1, 256, 64, 64, std=1.0)
        segs = torch.rand(k_props, 256, 256, generator=g)
        hid = rn(1, 4096, std=1.0)

        def sel():
            t = selector.text_hidden_fc(hid, ssd)
            return selector.selector_forward(selector.upsample_embeddings(e)[0], segs, t, ssd)
        parts["selector"] = _timeit(sel, reps)

    total = sam_s + clip_s + llama_s + parts["selector"]
    out.update(value=1.0 / total, unit="images/s", seconds_per_image=total,
               parts_s={k: round(v, 4) for k, v in parts.items()},
               sample=(f"fp32 eager oracle, 1 image, {t_text}-tok prompt, {k_props} proposals: one instance of each "
                       f"layer type timed at full width (SAM window+global block, CLIP layer, LLaMA layer @T={T}, "
                       f"patch embeds, neck, projector, full selector) x layer counts 28/4/23/32"))
    return out


def _pick_threads() -> int:
    """"all the host threads it can use": eager PyTorch stops scaling (and regresses) well before 100+ threads,
    so pick the fastest of {all cores, 64, 32, 16} on one SAM-sized GEMM."""
    ncpu = os.cpu_count() or 1
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(4096, 1280, generator=g), torch.randn(5120, 1280, generator=g)
    best_t, threads = None, ncpu
    for cand in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(cand)
        t = _timeit(lambda: a @ b.T, 2)
        if best_t is None or t < best_t:
            best_t, threads = t, cand
    return threads


def _aliased(sd: Dict[str, torch.Tensor], fmt: str, src: int, dst_range) -> None:
    """sd[fmt.format(i) + rest] = sd[fmt.format(src) + rest] for i in dst_range (same tensor objects)."""
    p = fmt.format(src)
    for k in [k for k in sd if k.startswith(p)]:
        for i in dst_range:
            sd[fmt.format(i) + k[len(p):]] = sd[k]


def full_state_dict_aliased():
    """Full-depth reference-named fp32 state dict whose layers of one type SHARE their tensors (one windowed SAM
    block, one global SAM block, one CLIP layer, one LLaMA layer; ~1.9 GB instead of 31 GB of host memory and
    seconds instead of minutes of random-number generation).  The forward executes all 32 + 23 + 32 layers for real —
    the arithmetic, the data dependencies and the bytes streamed per layer are those of distinct weights (a
    layer's weights are far larger than the last-level cache either way)."""
    from . import lisa_forward
    scfg = sam_encoder.SamConfig(depth=2, global_attn_indexes=(1,))
    sam = sam_encoder.random_state_dict(scfg, seed=0, prefix="model.visual_model.image_encoder.")
    full = sam_encoder.SamConfig()
    glob = set(full.global_attn_indexes)
    blk = "model.visual_model.image_encoder.blocks.{}."
    gl = {k: v for k, v in sam.items() if k.startswith(blk.format(1))}
    for k in gl:
        del sam[k]
    for i in glob:                                   # global blocks <- the generated global block
        for k, v in gl.items():
            sam[blk.format(i) + k[len(blk.format(1)):]] = v
    _aliased(sam, blk, 0, [i for i in range(1, full.depth) if i not in glob])
    sd = dict(sam)
    ccfg = clip_llama.ClipConfig(layers=1)
    clip = clip_llama.clip_random_state_dict(ccfg, seed=1, prefix="model.vision_tower.vision_tower.vision_model.")
    _aliased(clip, "model.vision_tower.vision_tower.vision_model.encoder.layers.{}.", 0, range(1, 24))
    sd.update(clip)
    ll = clip_llama.llama_random_state_dict(clip_llama.LlamaConfig(layers=1), seed=2, prefix="model.")
    _aliased(ll, "model.layers.{}.", 0, range(1, 32))
    sd.update(ll)
    sd.update({"model." + k: v for k, v in selector.random_state_dict(seed=3, hidden=4096).items()})
    g = torch.Generator().manual_seed(4)
    sd["model.mm_projector.weight"] = torch.randn(4096, 1024, generator=g) * 1024 ** -0.5
    sd["model.mm_projector.bias"] = torch.randn(4096, generator=g) * 0.02
    return sd, lisa_forward.LisaConfig()


def cpu_forward_full(t_text: int = 64, k_props: int = 64, threads: int | None = None, reps: int = 1,
                     warmup: int = 0) -> Dict:
    """`reps` REAL full-depth single-image forwards of the oracle (the reference algorithm as fp32 eager PyTorch,
    reference model/LISA.py:225-414 semantics: one image per call) on the host threads; value = 1 / median."""
    from . import lisa_forward
    threads = threads or _pick_threads()
    torch.set_num_threads(threads)
    sd, cfg = full_state_dict_aliased()
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(3, 31999, (1, t_text), generator=g)
    ids[0, 0], ids[0, 1], ids[0, 2], ids[0, 3] = 1, 32001, lisa_forward.IMAGE_TOKEN_INDEX, 32002
    ids[0, t_text - 3], ids[0, t_text - 2], ids[0, t_text - 1] = cfg.seg_token_idx, 29889, 2
    inp = dict(images=torch.randn(1, 3, 1024, 1024, generator=g), images_clip=torch.randn(1, 3, 224, 224, generator=g),
               input_ids=ids, attention_masks=torch.ones(1, t_text, dtype=torch.bool), offset=torch.arange(2),
               sam_segs_list=[torch.rand(k_props, 256, 256, generator=g)])
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            out = lisa_forward.model_forward_inference(sd, cfg, **inp)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    assert out["pred_similarity"][0].shape == (1, k_props)
    times.sort()
    med = times[len(times) // 2]
    return {"value": 1.0 / med, "unit": "images/s", "cores": threads, "kind": "port", "seconds": [round(t, 3) for t in times],
            "spread": round((times[-1] - times[0]) / med, 3),
            "sample": (f"{reps} full single-image forward(s) of the fp32 eager oracle (SAM ViT-H 32 blocks + CLIP 23 layers + "
                       f"LLaMA-7B 32 layers @T={t_text + 255} + selector, {k_props} proposals), {warmup} warm-up, median; "
                       f"weights of same-type layers aliased to bound host memory")}


if __name__ == "__main__":
    import json
    print(json.dumps(cpu_forward_sample(), indent=1))
