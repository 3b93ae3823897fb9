"""Per-op time breakdown of one eager forward (batch 8 by default): CUDA events around every ops.* call,
aggregated by (op, shape).  Hot-L2 live timings (unlike the ncu launch list, which is cold/serialised)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llmseg_b200 import lisa, synthetic, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = lisa.LisaCfg()
dev = "cuda"
sd = synthetic.lisa_state_dict(cfg, seed=0, device=dev)
model = lisa.LISAForCausalLM(sd, cfg, device=dev, use_cuda_graph=False)
inp = synthetic.make_inputs(cfg, B, 64, 64, device=dev)

records = []
recording = [False]


def shape_key(name, args, kwargs):
    ts = [a for a in args if torch.is_tensor(a)]
    if name == "gemm":
        a, w = ts[0], ts[1]
        extra = ("+bias" if len(args) > 2 and args[2] is not None or kwargs.get("bias") is not None else "") + \
                ("+" + kwargs["act"] if kwargs.get("act") else "") + ("+res" if kwargs.get("residual") is not None else "") + \
                ("+map" if kwargs.get("out_row_map") is not None else "") + ("+swiglu" if kwargs.get("swiglu") else "") + ("+norm" if kwargs.get("row_stats") is not None else "")
        return f"gemm {a.shape[0]}x{w.shape[0]}x{a.shape[1]}{extra}"
    if name == "gemm_qkv":
        a, w = ts[0], ts[1]
        return f"gemm_qkv {a.shape[0]}x{w.shape[0]}x{a.shape[1]}" + ("+rope" if kwargs.get("rope_cos") is not None else "") + \
               ("+map" if kwargs.get("row_map") is not None else "") + ("+norm" if kwargs.get("row_stats") is not None else "")
    if name == "attention":
        return f"attention b={kwargs['batch']} h={kwargs['heads']} hd={kwargs['head_dim']} s={kwargs['seq']} ext={kwargs.get('ext_cols', 0)}"
    if name in ("layernorm", "rmsnorm", "norm_stats"):
        return f"{name} {tuple(ts[0].shape)}" + ("+map" if kwargs.get("src_row_map") is not None else "")
    if name == "relpos_prep":
        return f"relpos_prep bh={kwargs['bh']} s={kwargs['seq']}"
    return name + (" " + str(tuple(ts[0].shape)) if ts else "")


def wrap(name, fn):
    def inner(*args, **kwargs):
        if not recording[0]:
            return fn(*args, **kwargs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kwargs)
        e1.record()
        records.append((shape_key(name, args, kwargs), e0, e1))
        return r
    return inner


for name in ["gemm", "gemm_qkv", "attention", "layernorm", "rmsnorm", "norm_stats", "relpos_prep", "fill_kv_rows", "patchify",
             "embed_splice", "add_rows_bcast", "im2col3x3", "maskpool", "small_attention", "select"]:
    if hasattr(ops, name):
        setattr(ops, name, wrap(name, getattr(ops, name)))

with torch.no_grad():
    for it in range(3):
        recording[0] = it == 2
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        model.model_forward(**inp)
        t1.record()
        torch.cuda.synchronize()
print(f"eager forward with probes: {t0.elapsed_time(t1):.2f} ms  (B={B})")
agg = collections.OrderedDict()
for key, e0, e1 in records:
    t = e0.elapsed_time(e1) * 1e3
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"sum of op times: {tot / 1e3:.2f} ms over {len(records)} ops")
for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.1f}  {key}")
