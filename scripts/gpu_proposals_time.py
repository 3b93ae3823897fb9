"""Per-op time breakdown of SAM-Everything proposal generation for one image (CUDA events around every ops.* call of
llmseg_b200/proposals.py, aggregated by (op, shape)) + wall time of the whole generate() call."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import ops, proposals, synthetic

dev = "cuda"
ppb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
gen = proposals.SamProposalGenerator(synthetic.sam_decoder_state_dict(8, dev), dev)
tok = torch.randn(4096, 256, device=dev).bfloat16()
records, recording = [], [False]


def key(name, args, kwargs):
    ts = [a for a in args if torch.is_tensor(a)]
    if name == "gemm":
        a, w = ts[0], ts[1]
        return f"gemm {a.shape[0]}x{w.shape[0]}x{a.shape[1]}" + ("+" + kwargs["act"] if kwargs.get("act") else "") + \
               ("+res" if kwargs.get("residual") is not None else "")
    return name + (" " + str(tuple(ts[0].shape)) if ts else "")


def wrap(name, fn):
    def inner(*a, **kw):
        if not recording[0]:
            return fn(*a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **kw); e1.record()
        records.append((key(name, a, kw), e0, e1))
        return r
    return inner


for name in ("gemm", "layernorm", "add_rows_bcast", "small_attention", "point_tokens", "tok2img_attention", "img2tok_attention",
             "ln64_gelu", "mask_logits", "upscale_logits", "mask_stats", "box_nms", "mask_soft", "mask_binarize"):
    setattr(ops, name, wrap(name, getattr(ops, name)))
kw = dict(points_per_batch=ppb, pred_iou_thresh=-10.0, stability_score_thresh=0.5, box_nms_thresh=0.7)
with torch.no_grad():
    for it in range(3):
        recording[0] = it == 2
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = gen.generate(tok, **kw)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
agg = collections.OrderedDict()
for k, e0, e1 in records:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"points_per_batch {ppb}: wall {wall:.2f} ms (with event overhead), sum of op times {tot:.2f} ms, {len(records)} ops, "
      f"{out['n_masks']} masks after NMS")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {v[1]:8.3f} ms  {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  avg {v[1] / v[0] * 1e3:8.1f} us  {k}")
