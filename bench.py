#!/usr/bin/env python
"""bench.py — images/s of the LLM-Seg inference forward (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N --steps K --warmup W --batch B]          (N>1: launched by torchrun)
  python bench.py --impl reference ...                                (CPU reference arm)

A "step" is one forward over one batch of synthetic ReasonSeg-shaped inputs per GPU (weak scaling:
`--batch` images per GPU).  `value` times steps whose inputs are already resident in HBM; `e2e`
times the public API call `LISAForCausalLM.forward(**input_dict)` with HOST (pinned) inputs, the
host→device copies and the device→host read of the selected indices inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec fwd 1024px+64tok"
T_TEXT, K_PROPS = 64, 64

# algorithmic FLOPs per image: fused-attention cores
GF_SAM_GLOBAL_ATTN = 85.899 + 1.342   # per global block: QK^T + PV + decomposed rel-pos
GF_SAM_WINDOW_ATTN = 4.917 + 0.351    # per windowed block


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.first = 0

    def mark(self):
        """Start of the timed region: samples taken before this call (warm-up) are dropped.  The sampler is
        started before the warm-up because nvidia-smi needs ~0.5 s before its first line."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[min(self.first, max(len(self.rows) - 1, 0)):]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle port: fp32 eager PyTorch on the
    host threads), one REAL full-depth single-image forward per step (oracle/cpu_baseline.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.cpu_baseline import cpu_forward_full
    steps = max(1, min(args.steps, 3))
    warm = max(0, min(args.warmup, 1))
    t0 = time.perf_counter()
    res = cpu_forward_full(args.t_text, K_PROPS, reps=steps, warmup=warm)
    wall = time.perf_counter() - t0
    v = res["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"1 image per step, 1024px, {args.t_text}-tok prompt, {K_PROPS} proposals: full-depth forward "
                               f"(SAM ViT-H + CLIP ViT-L/14 + LLaMA-7B + selector) on the host CPU"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": res["cores"], "kind": "port", "sample": res["sample"],
                         "seconds_per_image": res["seconds"], "spread": res["spread"]},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(wall, 1),
    }
    print(json.dumps(line), flush=True)


def main():
    global T_TEXT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step (weak scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU and eager-GPU baseline legs")
    ap.add_argument("--sync-gather", action="store_true",
                    help="blocking all-gather every step (default: issued async, consumed one step later)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the configs[3] / configs[4] secondary objects")
    # non-default workloads (the default is the BASELINE metric's configuration)
    ap.add_argument("--t-text", type=int, default=T_TEXT, help="prompt tokens (configs[4]: 512)")
    ap.add_argument("--encoder", default="sam", choices=["sam", "dinov2"],
                    help="image branch: SAM ViT-H (north_star) or DINOv2 ViT-L/14 + lisa_dino_conv (reference LISA.py:244-245)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from llmseg_b200 import _lib, lisa, ops, synthetic
    from llmseg_b200 import dist as lsd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    rank, world, local = lsd.init_from_env("nccl")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    warmup = max(args.warmup, 3)
    B = args.batch
    cfg = lisa.LisaCfg()
    cfg.image_encoder = args.encoder
    T_TEXT = args.t_text
    sd = synthetic.lisa_state_dict(cfg, seed=0, device=dev)          # same weights on every rank
    model = lisa.LISAForCausalLM(sd, cfg, device=str(dev))
    del sd
    torch.cuda.empty_cache()
    inp = synthetic.make_inputs(cfg, B, K_PROPS, T_TEXT, seed=1234 + 1000 * rank, device=dev)
    k_max, b_max = K_PROPS, B
    pending = []          # the in-flight all-gather of the previous step: [(gathered tensor, work handle)]
    last = {}

    def consume():
        """Finish the previous step's all-gather (the logits of every rank's images become readable)."""
        if pending:
            out, work = pending.pop()
            if work is not None:
                work.wait()
            last["gathered"] = out

    def make_steps(model_inputs, host_inputs, batch):
        def pack(out):
            return lsd.pack_logits(out["similarity_padded"], out["iou_padded"], out["best_index"], [K_PROPS] * batch,
                                   k_max, batch)

        def resident():
            packed = pack(model.model_forward(**model_inputs))
            if args.sync_gather:
                last["gathered"] = lsd.all_gather_logits(packed)
                return
            # ONE collective per forward, off the critical path: this rank starts its next forward while slower ranks
            # finish this one; the result is consumed one step later (and drained before the clock stops)
            consume()
            pending.append(lsd.all_gather_logits(packed, async_op=True))

        def e2e():
            packed = pack(model.forward(**host_inputs))   # pinned host tensors: forward() stages them with async H2D copies
            res = lsd.all_gather_logits(packed)
            return res.cpu()   # device -> host read of every image's logits + selected index
        return resident, e2e

    def to_host(x):
        h = {k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in x.items()}
        h["sam_segs_list"] = [t.cpu().pin_memory() for t in x["sam_segs_list"]]
        return h

    host = to_host(inp)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("images", "images_clip", "input_ids", "attention_masks"))
    h2d += sum(t.numel() * t.element_size() for t in host["sam_segs_list"])
    step_resident, step_e2e = make_steps(inp, host, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_rank=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        consume()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            if per_rank is not None:
                allms = torch.empty(world, device=dev)
                dist.all_gather_into_tensor(allms, ms)
                per_rank.extend(round(float(v) / steps, 3) for v in allms.tolist())
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        elif per_rank is not None:
            per_rank.append(round(float(ms.item()) / steps, 3))
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        step_resident()
    torch.cuda.synchronize()
    sampler.mark()
    rank_ms = []
    total_ms = timed(step_resident, args.steps, rank_ms)  # CUDA-graph replay of the captured forward
    launches = model.last_forward_launches                # kernels captured in (= launched by) one forward
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    # ---- N > 1: is the all-gather the concatenation of single-GPU results?  Outside any timed region, rank 0
    # rebuilds every other rank's seeded inputs, runs them itself and compares with the rows it gathered — bit-equal,
    # since the replicas are independent and the kernels deterministic (SURVEY §4 tier 4).
    gather_check = per_rank = None
    if world > 1:
        step_resident()
        consume()
        torch.cuda.synchronize()
        gathered = last["gathered"].clone()
        if rank == 0:
            ok, worst = True, 0.0
            for r in range(world):
                inp_r = synthetic.make_inputs(cfg, B, K_PROPS, T_TEXT, seed=1234 + 1000 * r, device=dev)
                out_r = model.model_forward(**inp_r)
                exp = lsd.pack_logits(out_r["similarity_padded"], out_r["iou_padded"], out_r["best_index"],
                                      [K_PROPS] * B, k_max, b_max)
                got = gathered[r * b_max:(r + 1) * b_max]
                ok &= bool(torch.equal(got, exp))
                fin = torch.isfinite(exp)
                worst = max(worst, float((got[fin] - exp[fin]).abs().max()))
            gather_check = {"ranks_verified": world, "rows": int(gathered.shape[0]), "bit_equal": ok, "max_abs_diff": worst,
                            "what": "rank 0 recomputed every rank's seeded inputs locally and compared with the rows it gathered"}
        # per-rank pace: forward time and the time this rank then spends blocked in a SYNCHRONOUS all-gather
        fwd, gat = [], []
        for _ in range(min(args.steps, 5)):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            dist.barrier()
            e[0].record()
            out = model.model_forward(**inp)
            e[1].record()
            lsd.all_gather_logits(lsd.pack_logits(out["similarity_padded"], out["iou_padded"], out["best_index"],
                                                  [K_PROPS] * B, k_max, b_max))
            e[2].record()
            torch.cuda.synchronize()
            fwd.append(e[0].elapsed_time(e[1]))
            gat.append(e[1].elapsed_time(e[2]))
        mine = torch.tensor([statistics.median(fwd), statistics.median(gat)], device=dev)
        allr = torch.empty(2 * world, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, 2).tolist()
        per_rank = {"timed_region_ms_per_step": rank_ms, "forward_ms": [round(v[0], 3) for v in allr],
                    "blocked_in_sync_all_gather_ms": [round(v[1], 3) for v in allr],
                    "gather": "sync" if args.sync_gather else "async, consumed one step later"}

    # ---- secondary workloads on the same replica set (reported like batch1; `value` / `e2e` stay configs[2]):
    # BASELINE configs[3] = batch 32 over 8 GPUs = 4 images / GPU; configs[4] = 512-token prompts, batch 16 over
    # 8 GPUs = 2 images / GPU.  At N < 8 the same per-GPU shapes are timed (weak scaling: global batch = N x per-GPU).
    extra = {}
    if not args.no_extra_configs and args.encoder == "sam" and T_TEXT == 64 and B == 8:
        for name, bb, tt in (("configs3", 4, 64), ("configs4", 2, 512)):
            x = synthetic.make_inputs(cfg, bb, K_PROPS, tt, seed=4321 + 1000 * rank, device=dev)
            res_x, e2e_x = make_steps(x, to_host(x), bb)
            for _ in range(3):
                res_x()
            ms = timed(res_x, args.steps)
            for _ in range(2):
                e2e_x()
            ms2 = timed(e2e_x, args.steps)
            extra[name] = {"workload": (f"configs[{name[-1]}]-shaped: batch={bb}/GPU x {world} GPU(s) = {bb * world}, "
                                        f"{tt}-tok prompt, {K_PROPS} proposals, full fwd"),
                           "value": round(bb * world * args.steps / (ms / 1e3), 3), "unit": "images/s",
                           "ms_per_step": round(ms / args.steps, 3),
                           "e2e": round(bb * world * args.steps / (ms2 / 1e3), 3)}

    # BASELINE configs[1] (batch = 1, the latency configuration) next to the throughput configuration: same
    # forward, one image per step, host inputs (reported as a secondary object; `value` / `e2e` stay configs[2]).
    b1 = None
    if B != 1 and rank == 0 and world == 1:
        inp1 = synthetic.make_inputs(cfg, 1, K_PROPS, T_TEXT, seed=99, device=dev)
        host1 = to_host(inp1)

        def step_b1():
            return model.forward(**host1)["best_index"].cpu()
        for _ in range(3):
            step_b1()
        b1_ms = timed(step_b1, args.steps) / args.steps
        b1 = {"workload": "configs[1]: batch=1 full fwd, host inputs", "ms_per_image": round(b1_ms, 3),
              "value": round(1e3 / b1_ms, 3), "unit": "images/s"}

    # SAM-Everything proposal generation on the same SAM encoder: features -> prompt encoder -> mask
    # decoder over the reference's default 32 x 32 point grid -> statistics / filters / NMS -> top-50 soft proposals.
    # Timed per image with the host-side filter / sort and its device syncs INSIDE the region (rank 0, N = 1).
    prop = None
    if rank == 0 and world == 1 and args.encoder == "sam" and not args.no_extra_configs:
        from llmseg_b200 import proposals as lsp
        gen = lsp.SamProposalGenerator(synthetic.sam_decoder_state_dict(8, dev), dev)
        tok = model.sam.forward(inp["images"][:1])[0].contiguous()
        kw = dict(pred_iou_thresh=-10.0, stability_score_thresh=0.5, box_nms_thresh=0.7)   # random weights: keep the
        gen.generate(tok, **kw)                                                            # filters busy, not empty
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(3):
            rec = gen.generate(tok, **kw)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        prop = {"workload": "SAM-Everything on one image's SAM features: 1024 point prompts x 3 masks (32 x 32 grid, "
                            "batches of 256), filters + box NMS + top-50 soft proposals [K,256,256]",
                "ms_per_image": round(ms, 2), "images_per_s": round(1e3 / ms, 2), "launches_per_image": (_lib.launch_count() - n0) // 3,
                "proposals": int(rec["segs"].shape[0]), "masks_after_nms": rec["n_masks"]}
        del gen, tok
        torch.cuda.empty_cache()

    # dominant-kernel timing: the same steps launched eagerly (graph nodes cannot carry timing events) with
    # CUDA events, on the launching stream, around every GEMM launch (the tcgen05 GEMM kernel is ~75 % of
    # the step) and every SAM global-attention launch.
    probes = []   # (group key, algorithmic flops, start event, end event)
    orig = {n: getattr(ops, n) for n in ("attention", "gemm", "gemm_qkv", "relpos_prep", "fill_kv_rows")}

    def probed(name, key_fn):
        fn = orig[name]

        def inner(*a, **kw):
            key = key_fn(a, kw)
            if key is None:
                return fn(*a, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            probes.append((key[0], key[1], s, e))
            return r
        return inner

    def gemm_key(a, kw):
        M, K = a[0].shape
        N = a[1].shape[0]
        tag = "+".join(t for t, on in (("bias", a[2] is not None if len(a) > 2 else kw.get("bias") is not None),
                                       (str(kw.get("act")), kw.get("act") is not None),
                                       ("residual", kw.get("residual") is not None), ("swiglu", kw.get("swiglu", False)),
                                       ("folded norm", kw.get("row_stats") is not None)) if on)
        return (f"gemm {M}x{N}x{K} {tag}".strip(), 2.0 * M * N * K)

    def qkv_key(a, kw):
        M, K = a[0].shape
        N = a[1].shape[0]
        return (f"gemm_qkv {M}x{N}x{K}" + (" rope" if kw.get("rope_cos") is not None else ""), 2.0 * M * N * K)

    def attn_key(a, kw):
        if kw.get("ext_cols", 0) == 64:
            return ("attn_global", GF_SAM_GLOBAL_ATTN * B * 1e9)
        return (f"attn_other hd={kw.get('head_dim')} seq={kw.get('seq')}", 0.0)

    def aux_key(a, kw):
        return ("attn_aux (rel-pos prep / padding keys)", 0.0)

    model.use_cuda_graph = False
    overlap, model.overlap_branches = model.overlap_branches, False   # per-kernel times: one kernel at a time
    model.model_forward(**inp)      # un-probed eager pass: the first eager launches allocate (cudaMalloc syncs)
    torch.cuda.synchronize()
    ops.attention, ops.gemm, ops.gemm_qkv = probed("attention", attn_key), probed("gemm", gemm_key), probed("gemm_qkv", qkv_key)
    ops.relpos_prep, ops.fill_kv_rows = probed("relpos_prep", aux_key), probed("fill_kv_rows", aux_key)
    try:
        for _ in range(min(args.steps, 3)):
            model.model_forward(**inp)
        torch.cuda.synchronize()
    finally:
        for n, f in orig.items():
            setattr(ops, n, f)
        model.use_cuda_graph = True
        model.overlap_branches = overlap
    groups = {}
    for key, fl, s, e in probes:
        g = groups.setdefault(key, [0, 0.0, 0.0])
        g[0] += 1
        g[1] += fl
        g[2] += s.elapsed_time(e)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peaks, peak_src = load_peaks()
    imgs = B * world * args.steps
    value = imgs / (total_ms / 1e3)
    e2e_value = imgs / (e2e_ms / 1e3)
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    traffic_db = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic_db = json.load(f)

    def roof_of(key, kernel, traffic_key):
        n, fl, ms = groups[key]
        achieved = fl / ms / 1e9                                # FLOP / ms / 1e9 == TFLOP/s
        # peak = the driver-measured SUSTAINED cuBLAS figure (the kernel is timed inside a long power-capped step); a
        # group can come out slightly above it (frac > 1: faster than cuBLAS ran back to back) — frac_of_burst puts the
        # same number against the burst figure of a cold GPU
        return {"kernel": kernel, "bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                "frac": round(achieved / peak, 4), "frac_of_burst": round(achieved / peaks["bf16_tflops"], 4),
                "traffic": traffic_db.get(traffic_key, {}).get("dram_bytes") if B == 8 else None,
                "peak_source": peak_src + " bf16_tflops_sustained", "flops_per_launch": fl / n,
                "avg_launch_ms": round(ms / n, 4), "launches_timed": n}

    roof = roof_attn = gemm_all = None
    gemm_groups = {k: v for k, v in groups.items() if k.startswith("gemm")}
    if gemm_groups:
        # dominant (kernel, shape): the GEMM group with the largest share of the step
        top = max(gemm_groups, key=lambda k: gemm_groups[k][2])
        tkey = "gemm_mlp1_b8" if "5120x1280" in top else ("gemm_swiglu_b8" if "swiglu" in top and "22016x4096" in top else "")
        roof = roof_of(top, f"gemm2_kernel (tcgen05 CTA-pair GEMM): {top}", tkey)
        if traffic_db.get(tkey, {}).get("tensor_pipe_pct") is not None:
            roof["tensor_pipe_pct"] = traffic_db[tkey]["tensor_pipe_pct"]
            roof["traffic_source"] = traffic_db[tkey].get("source")
        fl = sum(v[1] for v in gemm_groups.values())
        ms = sum(v[2] for v in gemm_groups.values())
        steps_probed = max(1, min(args.steps, 3))
        gemm_all = {"achieved": round(fl / ms / 1e9, 1), "unit": "TFLOP/s", "frac": round(fl / ms / 1e9 / peak, 4),
                    "ms_per_step": round(ms / steps_probed, 2), "launches_per_step": sum(v[0] for v in gemm_groups.values()) // steps_probed,
                    # every (shape, epilogue) group that takes >= 2 % of the GEMM time, largest first: the same
                    # kernel family at its other shapes, so that `roofline` above is not the only figure on record
                    "groups": [{"shape": k, "ms_per_step": round(v[2] / steps_probed, 3), "achieved": round(v[1] / v[2] / 1e9, 1),
                                "frac": round(v[1] / v[2] / 1e9 / peak, 4)}
                               for k, v in sorted(gemm_groups.items(), key=lambda kv: -kv[1][2]) if v[2] >= 0.02 * ms]}
    # "Achieved fraction of the attention roofline" (BASELINE north_star): every kernel of the fused attention
    # paths — Q/K/V projection GEMMs, rel-pos prep, padding keys, the attention kernels — against the algorithmic
    # FLOPs of those paths (QKV GEMMs + QK^T + PV + rel-pos, causal pairs counted once).
    fused_attn = None
    fa_keys = [k for k in groups if k.startswith("gemm_qkv") or k.startswith("attn")]
    if fa_keys:
        T = T_TEXT + 255
        if args.encoder == "sam":
            gf_img = 28 * (48.17 + GF_SAM_WINDOW_ATTN) + 4 * (40.27 + GF_SAM_GLOBAL_ATTN)
        else:   # DINOv2 ViT-L/14 @896: 24 x (QKV 2*4097*1024*3072 + core 4*4097^2*1024)
            gf_img = 24 * (2 * 4097 * 1024 * 3072 + 4 * 4097 ** 2 * 1024) / 1e9
        gf_img += 6.2 + 37.2                                                        # CLIP, 23 layers
        gf_img += 32 * (2 * T * 4096 * 12288 + 4 * (T * (T + 1) // 2) * 128 * 32) / 1e9   # LLaMA-7B
        steps_probed = max(1, min(args.steps, 3))
        ms = sum(groups[k][2] for k in fa_keys) / steps_probed
        ach = gf_img * B / ms                                                       # GFLOP / ms == TFLOP/s
        fused_attn = {"what": "Q/K/V GEMMs + rel-pos prep + attention kernels of SAM/DINOv2, CLIP and LLaMA",
                      "algorithmic_gflop_per_image": round(gf_img, 1), "ms_per_step": round(ms, 3),
                      "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
                      "share_of_step": round(ms / (total_ms / args.steps), 3)}
    if "attn_global" in groups:
        roof_attn = roof_of("attn_global", "attn_kernel<80,2> (SAM global attention, 64x64 tokens, rel-pos)", "attn_global_b8")
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": round(total_ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"{'configs[2]' if T_TEXT == 64 else 'configs[4]-shaped'}: batch={B}/GPU full fwd (SAM ViT-H + CLIP ViT-L/14 + LLaMA-7B + selector), "
                                f"1024px, {T_TEXT}-tok prompt, {K_PROPS} proposals, random-init weights") if args.encoder == "sam"
                   else (f"variant B: batch={B}/GPU full fwd (DINOv2 ViT-L/14 @896 + lisa_dino_conv + CLIP ViT-L/14 + "
                         f"LLaMA-7B + selector), {T_TEXT}-tok prompt, {K_PROPS} proposals, random-init weights"),
                   "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": "15.4 GB of weights streamed per step (>> 126 MB L2); no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(world * b_max * lsd.packed_width(k_max) * 4), "ms_per_step": round(e2e_ms / args.steps, 3)},
        "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
        "roofline": roof,
        "roofline_all_gemms": gemm_all,
        "roofline_attn": roof_attn,
        "attention_roofline": fused_attn,
        "batch1": b1,
        "configs3": extra.get("configs3"), "configs4": extra.get("configs4"),
        "gather_check": gather_check, "per_rank": per_rank, "proposals": prop,
    }
    if roof_attn is not None:
        # BASELINE.json's second metric: tensor-pipe utilisation of the fused attention kernels, from the committed
        # `ncu --set full` captures (profiles/; a number taken under the profiler, so it is quoted, not re-measured)
        for k_, name in (("attn_global_b8", "tensor_pipe_pct"), ("attn_win_b8", "tensor_pipe_pct_window_kernel")):
            if traffic_db.get(k_, {}).get("tensor_pipe_pct") is not None:
                roof_attn[name] = traffic_db[k_]["tensor_pipe_pct"]
                roof_attn[name + "_source"] = traffic_db[k_].get("source")
    if not args.no_cpu_baseline:
        # ---- baselines (checker code timed as a baseline, never part of the product path; rank 0, N = 1 only)
        # (1) the reference's algorithm as eager bf16 PyTorch on THIS GPU — cuBLAS GEMMs, materialised attention
        #     scores, one image per call like reference model/LISA.py:271: the number a B200 user of the reference
        #     would see, since no Blackwell kernel exists upstream
        if world == 1 and args.encoder == "sam":
            from oracle import lisa_forward as o_lf
            del model
            torch.cuda.empty_cache()
            sd = synthetic.lisa_state_dict(cfg, seed=0, device=dev)
            ocfg = o_lf.LisaConfig()
            x8 = synthetic.make_inputs(cfg, B, K_PROPS, T_TEXT, seed=1234, device=dev)

            def eager(n_img):
                return o_lf.forward_batched(sd, ocfg, images=x8["images"][:n_img], images_clip=x8["images_clip"][:n_img],
                                            input_ids=x8["input_ids"][:n_img], attention_masks=x8["attention_masks"][:n_img],
                                            sam_segs_list=x8["sam_segs_list"][:n_img])

            def ms_of(fn, reps):
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps
            with torch.no_grad():
                t1 = ms_of(lambda: eager(1), 3)
                tb = ms_of(lambda: eager(B), 2)
            line["eager_gpu_baseline"] = {
                "what": "oracle restatement of reference model/LISA.py:225-414 as eager bf16 PyTorch ops on the same B200 "
                        "(cuBLAS, materialised attention, one image per call as LISA.py:271 asserts)",
                "batch1_ms_per_image": round(t1, 2), "batch1_images_per_s": round(1e3 / t1, 2),
                f"batch{B}_as_{B}_calls_ms": round(tb, 2), f"batch{B}_images_per_s": round(B * 1e3 / tb, 2),
                "ours_over_eager_batch1": round(t1 / b1["ms_per_image"], 2) if b1 else None,
                f"ours_over_eager_batch{B}": round(value / (B * 1e3 / tb), 2)}
            if prop is not None:
                # the mask generator's eager counterpart: the oracle restatement of SamAutomaticMaskGenerator in fp32
                # PyTorch ops on the same GPU, 64 prompts per batch like the reference's default
                from oracle import sam_amg
                osd = {k[len("model.visual_model."):]: v.float() for k, v in synthetic.sam_decoder_state_dict(8, dev).items()}
                emb = torch.randn(1, 256, 64, 64, device=dev)
                with torch.no_grad():
                    tg = ms_of(lambda: sam_amg.generate(emb, osd, pred_iou_thresh=-10.0, stability_score_thresh=0.5), 1)
                line["proposals"]["eager_gpu_ms_per_image"] = round(tg, 1)
                line["proposals"]["ours_over_eager"] = round(tg / line["proposals"]["ms_per_image"], 2)
            del sd
            torch.cuda.empty_cache()
        # (2) the same algorithm on the host CPU: one REAL full-depth single-image forward
        from oracle.cpu_baseline import cpu_forward_full
        cb = cpu_forward_full(T_TEXT, K_PROPS, reps=1, warmup=0)
        line["cpu_baseline"] = {"value": round(cb["value"], 5), "unit": "images/s", "cores": cb["cores"],
                                "kind": cb["kind"], "sample": cb["sample"]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
