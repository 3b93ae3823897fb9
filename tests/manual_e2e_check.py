"""Developer GPU check (run by hand: `python tests/manual_e2e_check.py [full]`; lives under tests/ because it
imports the oracle): stage-by-stage parity of the CUDA path against the fp32 oracle (same bf16 weights), reduced
depth by default; `full` runs the real depths and times the forward."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from llmseg_b200 import lisa, synthetic, ops, _lib
from oracle import sam_encoder as o_sam, clip_llama as o_cl, selector as o_sel, lisa_forward as o_lf

full = "full" in sys.argv
B = 2 if not full else 1
K, T_text = 64, 64
cfg = lisa.LisaCfg()
if not full:
    cfg.sam.depth, cfg.sam.global_attn_indexes = 3, (1,)
    cfg.clip.layers = 4
    cfg.llama.layers = 2
dev = "cuda"
t0 = time.time()
sd = synthetic.lisa_state_dict(cfg, seed=0, device=dev)
torch.cuda.synchronize(); print(f"weights: {sum(v.numel() for v in sd.values())/1e9:.2f} B params in {time.time()-t0:.1f}s", flush=True)
model = lisa.LISAForCausalLM(sd, cfg, device=dev)
inp = synthetic.make_inputs(cfg, B, K, T_text, device=dev)
torch.cuda.synchronize(); print(f"model ready {time.time()-t0:.1f}s", flush=True)

ocfg = o_lf.LisaConfig(
    sam=o_sam.SamConfig(depth=cfg.sam.depth, global_attn_indexes=cfg.sam.global_attn_indexes),
    clip=o_cl.ClipConfig(layers=cfg.clip.layers), llama=o_cl.LlamaConfig(layers=cfg.llama.layers))
def rel(a, b):
    a, b = a.float(), b.float()
    return f"max|d|={(a-b).abs().max().item():.4f} mean|d|={(a-b).abs().mean().item():.5f} ref_rms={b.pow(2).mean().sqrt().item():.3f}"

with torch.no_grad():
    # ---- stage 1: SAM
    tok = model.sam.forward(inp["images"])
    torch.cuda.synchronize()
    if not full or "oracle" in sys.argv:
        sam_sd = {k: v.float() for k, v in o_lf.sub_dict(sd, "model.visual_model.image_encoder.").items()}
        ref = torch.cat([o_sam.image_encoder(inp["images"][b:b+1].float(), sam_sd, ocfg.sam) for b in range(B)], 0)
        ref_tok = ref.permute(0, 2, 3, 1).reshape(B, 4096, 256)
        print("[sam] ", rel(tok, ref_tok), flush=True)
        del sam_sd
    # ---- stage 2: CLIP + projector
    feats = model.clip.forward(inp["images_clip"])
    torch.cuda.synchronize()
    if not full or "oracle" in sys.argv:
        fsd = {k: v.float() for k, v in sd.items() if k.startswith("model.vision_tower") or k.startswith("model.mm_projector")}
        ref_feats = o_lf.encode_images(inp["images_clip"].float(), fsd, ocfg)
        print("[clip+proj] ", rel(feats, ref_feats), flush=True)
    # ---- stage 3: full forward
    out = model.model_forward(**inp)
    torch.cuda.synchronize()
    print("best_index", out["best_index"].tolist(), flush=True)
    if not full or "oracle" in sys.argv:
        fsd = {k: v.float() for k, v in sd.items()}
        oinp = dict(images=inp["images"].float(), images_clip=inp["images_clip"].float(), input_ids=inp["input_ids"],
                    attention_masks=inp["attention_masks"], sam_segs_list=[s.float() for s in inp["sam_segs_list"]])
        ref_out = o_lf.forward_batched(fsd, ocfg, **oinp)
        for b in range(B):
            s, r = out["pred_similarity"][b].float(), ref_out["pred_similarity"][b]
            i_, ri = out["pred_iou"][b].float(), ref_out["pred_iou"][b]
            top2 = r[0].topk(2).values
            print(f"[e2e img{b}] sim max|d|={(s-r).abs().max().item():.4f} iou max|d|={(i_-ri).abs().max().item():.4f} "
                  f"argmax mine={int(s.argmax())} ref={int(r.argmax())} margin={float(top2[0]-top2[1]):.4f}", flush=True)
        del fsd
    # ---- timing
    for _ in range(2): model.model_forward(**inp)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 5
    e0.record()
    for _ in range(n): model.model_forward(**inp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"forward B={B}: {ms:.2f} ms  -> {B/ms*1e3:.1f} img/s   launches/forward={(_lib.launch_count()-l0)//n}", flush=True)
    # per-stage timing
    def timeit(fn, n=5):
        fn(); torch.cuda.synchronize(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    print(f"  sam  {timeit(lambda: model.sam.forward(inp['images'])):.2f} ms", flush=True)
    print(f"  clip {timeit(lambda: model.clip.forward(inp['images_clip'])):.2f} ms", flush=True)
