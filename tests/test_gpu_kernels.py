"""GPU (-m gpu): each sm_100a kernel, called through the C ABI, against a plain PyTorch fp32
reference of the same op on the same bf16 inputs.  Tolerances are bf16 rounding-level: the kernels
accumulate in fp32 and round once per reference rounding point."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(t):
    return t.to(DEV).bfloat16()


def _ulp_tol(ref, rel=2 ** -7):
    """bf16 has 8 bits of mantissa: allow ~2 ulp of the largest magnitude involved."""
    return float(ref.abs().max()) * rel + 1e-3


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 1032, 1288), (4096, 3840, 1280), (257, 1024, 1024), (1, 256, 4096)])
@pytest.mark.parametrize("act", [None, "gelu", "quick_gelu", "relu"])
def test_gemm_bias_act_residual(cuda_lib, M, N, K, act):
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a, w = _bf(torch.randn(M, K, generator=g)), _bf(torch.randn(N, K, generator=g) / K ** 0.5)
    b, r = _bf(torch.randn(N, generator=g)), _bf(torch.randn(M, N, generator=g))
    out = ops.gemm(a, w, b, act=act, residual=r)
    ref = (a.float() @ w.float().T + b.float()).bfloat16().float()
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref).bfloat16().float()
    elif act == "quick_gelu":
        ref = (ref * torch.sigmoid(1.702 * ref)).bfloat16().float()
    elif act == "relu":
        ref = torch.relu(ref)
    ref = ref + r.float()
    assert (out.float() - ref).abs().max().item() <= _ulp_tol(ref)


def test_gemm_row_scatter_and_broadcast_residual(cuda_lib):
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(0)
    M, N, K, R = 500, 256, 128, 300
    a, w = _bf(torch.randn(M, K, generator=g)), _bf(torch.randn(N, K, generator=g) / K ** 0.5)
    perm = torch.randperm(M, generator=g)[:R]
    m = torch.full((M,), -1, dtype=torch.int32)
    m[perm] = torch.arange(R, dtype=torch.int32)
    res = _bf(torch.randn(50, N, generator=g))
    out = torch.zeros(R, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w, None, residual=res, res_mod=50, out=out, out_row_map=m.to(DEV))
    full = (a.float() @ w.float().T).bfloat16().float()
    ref = full[perm.to(DEV)] + res.float()[torch.arange(R, device=DEV) % 50]
    assert (out.float() - ref).abs().max().item() <= _ulp_tol(ref)


@pytest.mark.parametrize("M,N,K,mode", [(2552, 4096, 4096, "res"), (2552, 4096, 11008, "res"), (2552, 12288, 4096, "qkv"),
                                        (1100, 1024, 4096, "res"), (2552, 2048, 2048, "swiglu")])
def test_gemm_streamk_tail(cuda_lib, M, N, K, mode):
    """Shapes whose last wave of tiles is mostly empty take the stream-K tail (fp32 partials through the
    workspace): same result as the plain schedule up to fp32 summation order, and deterministic."""
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a, w = _bf(torch.randn(M, K, generator=g)), _bf(torch.randn(N, K, generator=g) / K ** 0.5)
    r = _bf(torch.randn(M, N, generator=g))

    def run():
        if mode == "qkv":
            H, hd, S = N // 384, 128, 319
            q = torch.zeros(M // S * H, 320, hd, device=DEV, dtype=torch.bfloat16)
            k, vt = torch.zeros_like(q), torch.zeros(M // S * H, hd, 320, device=DEV, dtype=torch.bfloat16)
            ops.gemm_qkv(a, w, None, q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=320)
            return torch.cat([q.flatten(), k.flatten(), vt.flatten()])
        if mode == "swiglu":
            return ops.gemm(a, w, None, swiglu=True)
        return ops.gemm(a, w, None, residual=r)

    outs = []
    for use in (False, True, True):
        ops.USE_GEMM_WORKSPACE = use
        try:
            outs.append(run().float())
        finally:
            ops.USE_GEMM_WORKSPACE = True
    plain, sk, sk2 = outs
    assert torch.equal(sk, sk2)
    d = (plain - sk).abs()
    assert d.max().item() <= _ulp_tol(plain) and d.mean().item() <= 1e-4
    if mode == "res":
        ref = (a.float() @ w.float().T).bfloat16().float() + r.float()
        assert (sk - ref).abs().max().item() <= _ulp_tol(ref)


@pytest.mark.parametrize("N,mode", [(1280, "res_stats"), (5120, "gelu"), (3840, "qkv")])
def test_gemm_streamk_tail_short_k(cuda_lib, N, mode, monkeypatch):
    """Batch-1 SAM shapes (M = 4096, K = 1280: 20 k-blocks, 1.08 / 4.3 / 3.2 waves of pair tiles) with the stream-K tail
    admitted for short K (LLMSEG_GEMM_SK_MIN_KB=16; the default floor of 32 k-blocks keeps whole tiles — measured no gain
    inside the two-stream step).  Against the default: same result up to fp32 summation order, row statistics included,
    and deterministic."""
    from llmseg_b200 import ops
    M, K = 4096, 1280
    g = torch.Generator(device=DEV).manual_seed(N)
    a = _bf(torch.randn(M, K, generator=g, device=DEV))
    w = _bf(torch.randn(N, K, generator=g, device=DEV) / K ** 0.5)
    b = _bf(torch.randn(N, generator=g, device=DEV))
    x0 = _bf(torch.randn(M, N, generator=g, device=DEV)) if mode == "res_stats" else None

    def run():
        if mode == "qkv":
            H, hd, S = 16, 80, 512
            q = torch.zeros(M // S * H, S, hd, device=DEV, dtype=torch.bfloat16)
            k, vt = torch.zeros_like(q), torch.zeros(M // S * H, hd, S, device=DEV, dtype=torch.bfloat16)
            ops.gemm_qkv(a, w, b, q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=S, row_stats=ops.norm_stats(a, 1e-6))
            return torch.cat([q.flatten(), k.flatten(), vt.flatten()]).float(), None
        if mode == "gelu":
            return ops.gemm(a, w, b, act="gelu", row_stats=ops.norm_stats(a, 1e-6)).float(), None
        x = x0.clone()
        so = ops.gemm_stats_buffer(M, N, M, 1e-6)
        ops.gemm(a, w, b, residual=x, out=x, stats_out=so)
        return x.float(), so.final.clone()
    y0, s0 = run()
    monkeypatch.setenv("LLMSEG_GEMM_SK_MIN_KB", "16")
    y1, s1 = run()
    y2, s2 = run()
    assert torch.equal(y1, y2)
    d = (y0 - y1).abs()
    assert d.max().item() <= _ulp_tol(y0) and d.mean().item() <= 1e-4
    if s0 is not None:
        assert torch.equal(s1, s2)
        assert (s0 - s1).abs().max().item() <= 2e-3 * float(s0.abs().max())


@pytest.mark.parametrize("variant", ["gelu", "res_stats", "res", "bias_res_mod", "bias"])
def test_gemm_epilogue_variants_equal_generic(cuda_lib, variant, monkeypatch):
    """The straight-line epilogue variants of the CTA-pair kernel (csrc/gemm.cu EpiX 1-4: bias + GELU (+ folded norm),
    bias + TMA residual + row statistics, TMA residual alone, bias + TMA residual from a repeating row table) against the generic epilogue on the same problem —
    bit-equal outputs and statistics (same arithmetic, only the control flow is resolved at compile time) — and
    against torch within the usual GEMM bound.  M x N = 2100 x 1280: 17 row tiles, a ragged last one."""
    from llmseg_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(21)
    M, N, K = 2100, 1280, 640
    a = _bf(torch.randn(M, K, generator=g, device=DEV))
    w = _bf(torch.randn(N, K, generator=g, device=DEV) / K ** 0.5)
    b = _bf(torch.randn(N, generator=g, device=DEV))
    x0 = _bf(torch.randn(M, N, generator=g, device=DEV))

    def run():
        if variant == "gelu":
            st = ops.norm_stats(a, 1e-6)
            return ops.gemm(a, w, b, act="gelu", row_stats=st), None
        if variant == "bias":           # variant 5: bias alone
            return ops.gemm(a, w, b), None
        if variant == "bias_res_mod":   # variant 4: bias + a 256-row residual table repeated down the output (TMA landing)
            return ops.gemm(a, w, b, residual=x0[:256], res_mod=256), None
        x = x0.clone()
        if variant == "res_stats":
            so = ops.gemm_stats_buffer(M, N, M, 1e-6)
            ops.gemm(a, w, b, residual=x, out=x, stats_out=so)
            return x, so.final.clone()
        ops.gemm(a, w, None, residual=x, out=x)
        return x, None
    monkeypatch.setenv("LLMSEG_GEMM_EPI", "0")
    y0, s0 = run()
    monkeypatch.setenv("LLMSEG_GEMM_EPI", "1")
    y1, s1 = run()
    assert torch.equal(y0, y1)
    if s0 is not None:
        assert torch.equal(s0, s1)
    af, wf = a.float(), w.float()
    if variant == "gelu":
        mean, var = af.mean(1, keepdim=True), af.var(1, unbiased=False, keepdim=True)
        ref = torch.nn.functional.gelu((af * torch.rsqrt(var + 1e-6)) @ wf.T + b.float())
    elif variant == "res_stats":
        ref = af @ wf.T + b.float() + x0.float()
    elif variant == "bias":
        ref = af @ wf.T + b.float()
    elif variant == "bias_res_mod":
        ref = af @ wf.T + b.float() + x0[:256].float().repeat(9, 1)[:M]
        monkeypatch.setenv("LLMSEG_GEMM_TMA_RES", "0")   # per-lane residual loads: same values, same arithmetic
        y2, _ = run()
        monkeypatch.delenv("LLMSEG_GEMM_TMA_RES")
        assert torch.equal(y1, y2)
    else:
        ref = af @ wf.T + x0.float()
    assert (y1.float() - ref).abs().max().item() <= 2 ** -7 * float(ref.abs().max()) + 2e-2


def test_gemm_swiglu(cuda_lib):
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(1)
    M, F, K = 319, 1024, 512
    a = _bf(torch.randn(M, K, generator=g))
    gate, up = _bf(torch.randn(F, K, generator=g) / K ** 0.5), _bf(torch.randn(F, K, generator=g) / K ** 0.5)
    w = torch.stack([gate, up], 1).reshape(2 * F, K).contiguous()
    out = ops.gemm(a, w, None, swiglu=True)
    gg, uu = (a.float() @ gate.float().T).bfloat16().float(), (a.float() @ up.float().T).bfloat16().float()
    ref = torch.nn.functional.silu(gg).bfloat16().float() * uu
    assert out.shape == (M, F)
    assert (out.float() - ref).abs().max().item() <= _ulp_tol(ref)


@pytest.mark.parametrize("rows,dim,eps", [(4096, 1280, 1e-6), (257, 1024, 1e-5), (319, 4096, 1e-6), (64, 256, 1e-5)])
def test_norms(cuda_lib, rows, dim, eps):
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(rows)
    x = _bf(torch.randn(rows, dim, generator=g) * 2 + 0.5)
    gm, bt = _bf(1 + 0.1 * torch.randn(dim, generator=g)), _bf(0.1 * torch.randn(dim, generator=g))
    y = ops.layernorm(x, gm, bt, eps)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), gm.float(), bt.float(), eps)
    assert (y.float() - ref).abs().max().item() <= _ulp_tol(ref)
    y = ops.rmsnorm(x, gm, eps)
    xf = x.float()
    ref = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).bfloat16().float() * gm.float()
    assert (y.float() - ref).abs().max().item() <= _ulp_tol(ref)
    # gather + zero rows (window padding tokens are zeros AFTER the norm)
    m = torch.tensor([3, -1, 0, rows - 1], dtype=torch.int32, device=DEV)
    y = ops.layernorm(x, gm, bt, eps, src_row_map=m, rows_out=4)
    ref4 = torch.nn.functional.layer_norm(x.float()[[3, 0, 0, rows - 1]], (dim,), gm.float(), bt.float(), eps)
    ref4[1] = 0
    assert (y.float() - ref4).abs().max().item() <= _ulp_tol(ref4)


@pytest.mark.parametrize("rows,dim", [(8192 + 3, 256), (20001, 64), (9000, 200)])
def test_narrow_row_norm(cuda_lib, rows, dim):
    """Rows of <= 256 values, >= 8192 of them (the mask decoder's token stream, the neck): the kernel that takes four
    rows per warp (csrc/norm.cu norm_narrow_kernel) — bit-equal to the one-row kernel (same arithmetic), against torch,
    with a row count that is no multiple of 4, a strided input, and in place."""
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(rows)
    xs = _bf(torch.randn(rows, dim + 8, generator=g) * 2 + 0.5)
    x = xs[:, :dim]
    gm, bt = _bf(1 + 0.1 * torch.randn(dim, generator=g)), _bf(0.1 * torch.randn(dim, generator=g))
    y = ops.layernorm(x, gm, bt, 1e-6)
    ref = torch.nn.functional.layer_norm(x.float(), (dim,), gm.float(), bt.float(), 1e-6)
    assert (y.float() - ref).abs().max().item() <= _ulp_tol(ref)
    one_row = torch.cat([ops.layernorm(x[i:i + 4096], gm, bt, 1e-6) for i in range(0, rows, 4096)])   # < 8192 rows per call
    assert torch.equal(y, one_row)
    yr = ops.rmsnorm(x, gm, 1e-6)
    xf = x.float()
    refr = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16().float() * gm.float()
    assert (yr.float() - refr).abs().max().item() <= _ulp_tol(refr)
    xc = x.contiguous()
    assert torch.equal(ops.layernorm(xc, gm, bt, 1e-6, out=xc), y)


@pytest.mark.parametrize("rows,dim,rms", [(6000, 1280, False), (2100, 4096, True), (3000, 1024, False)])
def test_streaming_norm_with_row_gather(cuda_lib, rows, dim, rms):
    """The bulk-copy-staged kernel (>= 1 MB, rows >= 2 KB) with a gather map that repeats rows and contains
    negative entries (zero rows), more rows than one pass of the persistent grid, and a strided input."""
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(rows + dim)
    xs = _bf(torch.randn(rows // 2, dim + 64, generator=g) * 1.5 - 0.3)
    x = xs[:, :dim]                                                    # row stride dim + 64
    gm, bt = _bf(1 + 0.1 * torch.randn(dim, generator=g)), _bf(0.1 * torch.randn(dim, generator=g))
    m = torch.randint(0, rows // 2, (rows,), generator=g)
    m[::17] = -1
    md = m.to(torch.int32).to(DEV)
    eps = 1e-6
    if rms:
        y = ops.rmsnorm(x, gm, eps, src_row_map=md, rows_out=rows)
        xf = x.float()
        full = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).bfloat16().float() * gm.float()
    else:
        y = ops.layernorm(x, gm, bt, eps, src_row_map=md, rows_out=rows)
        full = torch.nn.functional.layer_norm(x.float(), (dim,), gm.float(), bt.float(), eps)
    ref = full[m.clamp_min(0).to(DEV)]
    ref[(m < 0).to(DEV)] = 0
    assert y.shape == (rows, dim)
    assert (y.float() - ref).abs().max().item() <= _ulp_tol(ref)
    st = ops.norm_stats(x.contiguous(), eps, rms=rms)                  # statistics-only mode on the same kernel
    xf = x.float()
    rstd = torch.rsqrt((xf.pow(2).mean(-1) if rms else xf.var(-1, unbiased=False)) + eps)
    assert (st.t[:, 1] - rstd).abs().max().item() <= 1e-3 * float(rstd.max())
    if not rms:
        assert (st.t[:, 0] - xf.mean(-1)).abs().max().item() <= 1e-3


def _qkv_setup(B, H, hd, S, g, rope=False):
    from llmseg_b200 import ops
    D = H * hd
    S_pad = (S + 7) // 8 * 8
    x = _bf(torch.randn(B * S, D, generator=g))
    w = _bf(torch.randn(3 * D, D, generator=g) / D ** 0.5)
    bias = None if rope else _bf(torch.randn(3 * D, generator=g) * 0.1)
    q = torch.zeros(B * H, S_pad, hd, device=DEV, dtype=torch.bfloat16)
    k = torch.zeros_like(q)
    vt = torch.zeros(B * H, hd, S_pad, device=DEV, dtype=torch.bfloat16)
    cos = sin = None
    if rope:
        inv = 1.0 / (10000 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
        fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
        cos, sin = _bf(fr.cos()).contiguous(), _bf(fr.sin()).contiguous()
    ops.gemm_qkv(x, w, bias, q, k, vt, heads=H, head_dim=hd, seq_in=S, seq_pad=S_pad, rope_cos=cos, rope_sin=sin)
    ref = x.float() @ w.float().T
    if bias is not None:
        ref = ref + bias.float()
    ref = ref.reshape(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)   # epilogue is fp32 up to the store
    rq, rk, rv = ref[0], ref[1], ref[2]
    if rope:
        c = torch.cat([cos, cos], -1).float()[None, None]
        s_ = torch.cat([sin, sin], -1).float()[None, None]
        rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)
        ap = lambda t: t * c + rot(t) * s_
        rq, rk = ap(rq), ap(rk)
    return q, k, vt, rq, rk, rv, S_pad


def _attn_ref(qf, kf, vf, scale, mask=None, bias=None):
    s = (qf @ kf.transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return torch.softmax(s, -1) @ vf


@pytest.mark.parametrize("B,H,hd,S,causal,rope", [(2, 16, 64, 257, False, False), (1, 4, 80, 300, False, False),
                                                   (2, 32, 128, 319, True, True), (1, 2, 64, 128, False, False),
                                                   (1, 2, 128, 767, True, True)])
def test_qkv_split_and_attention(cuda_lib, B, H, hd, S, causal, rope):
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + S)
    q, k, vt, rq, rk, rv, S_pad = _qkv_setup(B, H, hd, S, g, rope)
    tol = _ulp_tol(rq)
    assert (q[:, :S].float().reshape(B, H, S, hd) - rq).abs().max().item() <= tol
    assert (k[:, :S].float().reshape(B, H, S, hd) - rk).abs().max().item() <= tol
    assert (vt[:, :, :S].float().reshape(B, H, hd, S).transpose(-1, -2) - rv).abs().max().item() <= tol
    kv_len = None
    mask = None
    if causal:
        lens = [S] + [max(S - 69, 1)] * (B - 1)
        kv_len = torch.tensor(lens, dtype=torch.int32, device=DEV)
        i = torch.arange(S, device=DEV)
        mask = (i[None, :] > i[:, None])[None, None] | (i[None, None, None, :] >= kv_len[:, None, None, None])
    qf, kf = q[:, :S].float().reshape(B, H, S, hd), k[:, :S].float().reshape(B, H, S, hd)
    vf = vt[:, :, :S].float().reshape(B, H, hd, S).transpose(-1, -2)
    ref = _attn_ref(qf, kf, vf, hd ** -0.5, mask).permute(0, 2, 1, 3).reshape(B * S, H * hd)
    out = torch.full((B * S, H * hd), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k, vt, out, batch=B, heads=H, head_dim=hd, seq=S, seq_pad=S_pad, scale=hd ** -0.5,
                  causal=causal, kv_len=kv_len)
    valid = torch.ones(B, S, dtype=torch.bool, device=DEV)
    if kv_len is not None:
        valid = torch.arange(S, device=DEV)[None] < kv_len[:, None]   # padded query rows are don't-care
    d = (out.float() - ref).reshape(B, S, -1)[valid]
    assert not torch.isnan(d).any()
    assert d.abs().max().item() <= 2e-2 * max(1.0, float(ref.abs().max()))
    assert d.abs().mean().item() <= 2e-3


@pytest.mark.parametrize("grid,nb", [(14, 25), (64, 1)])
def test_sam_relpos_attention(cuda_lib, grid, nb):
    """decomposed rel-pos through the tensor-core score extension vs the reference formula
    (reference image_encoder.py:354-392, restated in oracle/sam_encoder.py)."""
    from llmseg_b200 import ops
    from oracle import sam_encoder as o_sam
    H, hd, S = 4, 80, grid * grid
    g = torch.Generator().manual_seed(grid)
    q, k, vt, rq, rk, rv, S_pad = _qkv_setup(nb, H, hd, S, g)
    rel_h, rel_w = _bf(torch.randn(2 * grid - 1, hd, generator=g) * 0.2), _bf(torch.randn(2 * grid - 1, hd, generator=g) * 0.2)
    scale = hd ** -0.5
    ext = 32 if grid == 14 else 64
    qext = torch.zeros(nb * H, S_pad, ext, device=DEV, dtype=torch.bfloat16)
    rb = torch.zeros(nb * H, S_pad, 64, device=DEV, dtype=torch.bfloat16) if grid == 64 else None
    ops.relpos_prep(q, ops.make_rel_hw(rel_h, rel_w), bh=nb * H, seq=S, seq_pad=S_pad, head_dim=hd, grid=grid,
                    inv_scale=1 / scale, qext=qext, row_bias=rb)
    out = torch.empty(nb * S, H * hd, device=DEV, dtype=torch.bfloat16)
    ops.attention(q, k, vt, out, batch=nb, heads=H, head_dim=hd, seq=S, seq_pad=S_pad, scale=scale, qext=qext,
                  kext=ops.make_kext(grid, DEV), row_bias=rb, ext_cols=ext)
    qf = q[:, :S].float()
    bias = o_sam.decomposed_rel_pos_bias(qf, rel_h.float(), rel_w.float(), (grid, grid))
    vf = vt[:, :, :S].float().transpose(-1, -2)
    ref = _attn_ref(qf, k[:, :S].float(), vf, scale, bias=bias).reshape(nb, H, S, hd).permute(0, 2, 1, 3).reshape(nb * S, H * hd)
    d = out.float() - ref
    assert d.abs().max().item() <= 6e-2 * max(1.0, float(ref.abs().max()))
    assert d.abs().mean().item() <= 4e-3


def test_window_attention_running_max_rescale(cuda_lib):
    """The window kernel reads every 32-key chunk of a score row once and scales it with a RUNNING row maximum;
    when a later chunk's maximum exceeds it by more than 2^8 the chunks already written are rescaled in place.
    Force that path: keys whose scores climb by ~20 (natural units) from one 32-key chunk to the next for half of
    the rows, and fall for the other half (no rescale: both behaviours inside one warp)."""
    from llmseg_b200 import ops
    H, hd, S, Sp, nb = 2, 80, 196, 200, 3
    g = torch.Generator().manual_seed(11)
    scale = hd ** -0.5
    d = torch.nn.functional.normalize(torch.randn(hd, generator=g), dim=0)
    q = torch.randn(nb * H, S, hd, generator=g) * 0.3
    sign = torch.where(torch.arange(S) % 2 == 0, 1.0, -1.0)                 # even rows climb, odd rows fall
    q = q + sign[None, :, None] * 6.0 * d                                    # q.d = +-6
    step = (torch.arange(S) // 32).float() * (20.0 / (6.0 * scale))          # k.d grows per chunk: score += 20 per chunk
    k = torch.randn(nb * H, S, hd, generator=g) * 0.3 + step[None, :, None] * d
    v = torch.randn(nb * H, S, hd, generator=g)
    qp, kp = torch.zeros(nb * H, Sp, hd), torch.zeros(nb * H, Sp, hd)
    qp[:, :S], kp[:, :S] = q, k
    vt = torch.zeros(nb * H, hd, Sp)
    vt[:, :, :S] = v.transpose(-1, -2)
    qd, kd, vtd = _bf(qp), _bf(kp), _bf(vt)
    qext = torch.zeros(nb * H, Sp, 32, device=DEV, dtype=torch.bfloat16)     # zero rel-pos tables: plain attention
    out = torch.full((nb * S, H * hd), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qd, kd, vtd, out, batch=nb, heads=H, head_dim=hd, seq=S, seq_pad=Sp, scale=scale, qext=qext,
                  kext=ops.make_kext(14, DEV), ext_cols=32)
    qf, kf, vf = qd[:, :S].float(), kd[:, :S].float(), vtd[:, :, :S].float().transpose(-1, -2)
    sc = (qf @ kf.transpose(-1, -2)) * scale
    assert float((sc[:, 0::2, 160:].amax(-1) - sc[:, 0::2, :32].amax(-1)).min()) > 60      # the jump is really there
    ref = (torch.softmax(sc, -1) @ vf).reshape(nb, H, S, hd).permute(0, 2, 1, 3).reshape(nb * S, H * hd)
    dlt = out.float() - ref
    assert not torch.isnan(dlt).any()
    assert dlt.abs().max().item() <= 3e-2 * max(1.0, float(ref.abs().max())) and dlt.abs().mean().item() <= 3e-3


@pytest.mark.parametrize("rows,dim,N,rms", [(1000, 1280, 1024, False), (319, 4096, 512, True), (257, 1024, 768, False)])
def test_norm_folded_into_gemm(cuda_lib, rows, dim, N, rms):
    """y = act(Norm(x) @ W.T + b) with the norm folded into the GEMM (norm_stats + fold_norm + row_stats)
    against fp32 torch, and against the two-kernel path (norm, then GEMM) it replaces."""
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(rows + N)
    x = _bf(torch.randn(rows, dim, generator=g) * 3 + 1.5)     # non-zero mean: must cancel against the centred weights
    x[:, 5] += 40.0                                            # a massive-activation channel
    gm, bt = _bf(1 + 0.2 * torch.randn(dim, generator=g)), _bf(0.2 * torch.randn(dim, generator=g))
    w = _bf(torch.randn(N, dim, generator=g) / dim ** 0.5)
    b = None if rms else _bf(torch.randn(N, generator=g) * 0.1)
    st = ops.norm_stats(x, 1e-6, rms=rms)
    xf = x.float()
    if rms:
        ref_rstd = torch.rsqrt(xf.pow(2).mean(-1) + 1e-6)
        assert (st.t[:, 0] == 0).all() and torch.allclose(st.t[:, 1], ref_rstd, rtol=1e-5)
        normed = xf * ref_rstd[:, None] * gm.float()
    else:
        assert torch.allclose(st.t[:, 0], xf.mean(-1), rtol=1e-5, atol=1e-5)
        assert torch.allclose(st.t[:, 1], torch.rsqrt(xf.var(-1, unbiased=False) + 1e-6), rtol=1e-5)
        normed = torch.nn.functional.layer_norm(xf, (dim,), gm.float(), bt.float(), 1e-6)
    wq, b2 = ops.fold_norm(w, gm, None if rms else bt, b, rms=rms)
    assert (b2 is None) == rms
    if not rms:   # centred rows: the mean of x drops out of x @ wq.T up to the bf16 rounding of wq
        assert wq.float().sum(1).abs().max().item() <= 2 ** -8 * float(wq.float().abs().sum(1).max())
    out = ops.gemm(x, wq, b2, act=None if rms else "gelu", row_stats=st)
    ref = normed @ w.float().T
    if not rms:
        ref = torch.nn.functional.gelu(ref + b.float())
    tol = _ulp_tol(ref) + 2 ** -8 * float(ref.abs().max())    # + bf16 rounding of the gamma-scaled weights
    assert (out.float() - ref).abs().max().item() <= tol
    h = ops.rmsnorm(x, gm, 1e-6) if rms else ops.layernorm(x, gm, bt, 1e-6)
    two = ops.gemm(h, w, b, act=None if rms else "gelu")
    assert (out.float() - ref).abs().mean().item() <= 1.5 * (two.float() - ref).abs().mean().item() + 1e-4


@pytest.mark.parametrize("M,D,rms", [(4096, 1280, False), (319, 4096, True), (2056, 1024, False)])
def test_gemm_epilogue_row_statistics(cuda_lib, M, D, rms):
    """gemm(stats_out=...) leaves (sum, sum of squares) partials of every output row; a following
    gemm(row_stats=...) that finishes them must match the one that reads ops.norm_stats of the same rows."""
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(M + D)
    a, w = _bf(torch.randn(M, 512, generator=g)), _bf(torch.randn(D, 512, generator=g) / 512 ** 0.5)
    res = _bf(torch.randn(M, D, generator=g) * 2 + 0.7)
    st = ops.gemm_stats_buffer(M, D, M, 1e-5, rms=rms)
    st.t.fill_(float("nan"))
    x = ops.gemm(a, w, None, residual=res, stats_out=st)
    ref_x = (a.float() @ w.float().T) + res.float()
    tot = st.t.sum(1)
    assert not torch.isnan(tot).any()
    assert torch.allclose(tot[:, 0], ref_x.sum(-1), rtol=1e-4, atol=2e-2)
    assert torch.allclose(tot[:, 1], ref_x.pow(2).sum(-1), rtol=1e-4, atol=2e-2)
    # ... and the last tile of every 128-row block finished them into (mean, rstd)
    assert st.final is not None and not torch.isnan(st.final).any()
    mean = torch.zeros(M, device=DEV) if rms else ref_x.mean(-1)
    rstd = torch.rsqrt(ref_x.pow(2).mean(-1) - mean * mean + 1e-5)
    assert torch.allclose(st.final[:, 0], mean, rtol=1e-4, atol=1e-4)
    assert torch.allclose(st.final[:, 1], rstd, rtol=1e-4)
    w2 = _bf(torch.randn(256, D, generator=g) / D ** 0.5)
    gm = _bf(1 + 0.1 * torch.randn(D, generator=g))
    wq, _ = ops.fold_norm(w2, gm, rms=rms)
    y_final = ops.gemm(x, wq, None, row_stats=st)
    y_parts = ops.gemm(x, wq, None, row_stats=ops.RowStats(st.t, st.parts, st.dim, st.eps, st.rms))
    y_stats = ops.gemm(x, wq, None, row_stats=ops.norm_stats(x, 1e-5, rms=rms))
    for y in (y_final, y_parts):
        d = (y.float() - y_stats.float()).abs()
        assert d.max().item() <= _ulp_tol(y_stats.float()) and d.mean().item() <= 2e-3
    # the row-block counters re-arm themselves: a second launch gives the same statistics
    first = st.final.clone()
    ops.gemm(a, w, None, residual=res, stats_out=st)
    assert torch.equal(first, st.final)


def test_window_partition_folded_into_qkv_and_attention(cuda_lib):
    """SAM window partition / un-partition (reference image_encoder.py:263-318) as index maps on the QKV
    epilogue and the attention output: must equal the explicit pad -> project -> attend -> crop flow
    bit for bit (padding tokens are zeros after LayerNorm, so their k / v are the projection bias)."""
    from llmseg_b200 import ops
    from llmseg_b200.encoders import SamEncoder
    from llmseg_b200.lisa import SamCfg
    enc = SamEncoder.__new__(SamEncoder)
    enc.cfg, enc.device, enc._maps = SamCfg(), DEV, {}
    B, H, hd, ws = 1, 2, 80, 14
    D, S, sw, sw_pad = H * hd, 4096, 196, 200
    win_map, n_win, tok2win, pad_wins = enc._window_maps(B)
    nb = B * n_win
    g = torch.Generator().manual_seed(7)
    x = _bf(torch.randn(B * S, D, generator=g))
    w = _bf(torch.randn(3 * D, D, generator=g) / D ** 0.5)
    bias = _bf(torch.randn(3 * D, generator=g) * 0.3)
    rel = ops.make_rel_hw(_bf(torch.randn(27, hd, generator=g) * 0.2), _bf(torch.randn(27, hd, generator=g) * 0.2))
    kext, scale = ops.make_kext(14, DEV), hd ** -0.5

    def run(folded):
        q = torch.zeros(nb * H, sw_pad, hd, device=DEV, dtype=torch.bfloat16)
        k, vt = torch.zeros_like(q), torch.zeros(nb * H, hd, sw_pad, device=DEV, dtype=torch.bfloat16)
        qext = torch.zeros(nb * H, sw_pad, 32, device=DEV, dtype=torch.bfloat16)
        if folded:
            ops.gemm_qkv(x, w, bias, q, k, vt, heads=H, head_dim=hd, seq_in=sw, seq_pad=sw_pad, row_map=tok2win)
            ops.fill_kv_rows(k, vt, bias, win_map, batch=nb, heads=H, head_dim=hd, seq_in=sw, seq_pad=sw_pad,
                             seq_ids=pad_wins if folded == "listed" else None)
        else:
            xp = torch.zeros(nb * sw, D, device=DEV, dtype=torch.bfloat16)
            valid = win_map >= 0
            xp[valid] = x[win_map[valid].long()]
            ops.gemm_qkv(xp, w, bias, q, k, vt, heads=H, head_dim=hd, seq_in=sw, seq_pad=sw_pad)
        ops.relpos_prep(q, rel, bh=nb * H, seq=sw, seq_pad=sw_pad, head_dim=hd, grid=ws, inv_scale=1 / scale, qext=qext)
        if folded:
            out = torch.full((B * S, D), float("nan"), device=DEV, dtype=torch.bfloat16)
            ops.attention(q, k, vt, out, batch=nb, heads=H, head_dim=hd, seq=sw, seq_pad=sw_pad, scale=scale,
                          qext=qext, kext=kext, ext_cols=32, out_row_map=win_map)
            return k, vt, out
        o = torch.empty(nb * sw, D, device=DEV, dtype=torch.bfloat16)
        ops.attention(q, k, vt, o, batch=nb, heads=H, head_dim=hd, seq=sw, seq_pad=sw_pad, scale=scale,
                      qext=qext, kext=kext, ext_cols=32)
        out = torch.empty(B * S, D, device=DEV, dtype=torch.bfloat16)
        valid = win_map >= 0
        out[win_map[valid].long()] = o[valid]
        return k, vt, out

    assert pad_wins.numel() == 9      # right column, bottom row and the corner of the 5x5 window grid
    k0, vt0, o0 = run(False)
    for mode in (True, "listed"):
        k1, vt1, o1 = run(mode)
        assert torch.equal(k0, k1) and torch.equal(vt0, vt1)
        assert not torch.isnan(o1.float()).any()
        assert torch.equal(o0, o1)


def test_patchify_embed_splice_im2col(cuda_lib):
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(5)
    img = _bf(torch.randn(2, 3, 224, 224, generator=g))
    out = ops.patchify(img, 14, 592, cls_rows=1)
    ref = torch.nn.functional.unfold(img.float(), 14, stride=14).transpose(1, 2)     # [2,256,588] (c,py,px)
    got = out.float().reshape(2, 257, 592)
    assert torch.equal(got[:, 1:, :588], ref) and float(got[:, 1:, 588:].abs().sum()) == 0
    assert float(got[:, 0, :588].abs().sum()) == 0 and torch.all(got[:, 0, 588] == 1)
    img16 = _bf(torch.randn(1, 3, 64, 64, generator=g))
    o16 = ops.patchify(img16, 16, 768)
    assert torch.equal(o16.float().reshape(1, 16, 768), torch.nn.functional.unfold(img16.float(), 16, stride=16).transpose(1, 2))
    # splice (oracle: lisa_forward.splice_inputs / seg_token_mask)
    from oracle import lisa_forward as o_lf
    V, D, F_ = 100, 64, 256
    emb = _bf(torch.randn(V, D, generator=g))
    feats = _bf(torch.randn(2, F_, D, generator=g))
    ids = torch.tensor([[1, 7, -200, 8, 5, 6, 9, 50, 3, 2], [1, 7, -200, 8, 50, 6, 9, 11, 3, 2]], device=DEV)
    am = torch.ones(2, 10, dtype=torch.bool, device=DEV)
    am[1, 7:] = False
    cfg = o_lf.LisaConfig(seg_token_idx=50)
    e_ref, m_ref = o_lf.splice_inputs(ids, am, feats.float(), emb.float())
    e, kv_len, seg_row = ops.embed_splice(ids, am, emb, feats, image_token=-200, seg_token=50)
    assert torch.equal(e.float().reshape(2, 265, D), e_ref)
    assert kv_len.tolist() == m_ref.sum(-1).tolist()
    sm = o_lf.seg_token_mask(ids, cfg)
    assert seg_row.tolist() == [int(n * 265 + sm[n].nonzero()[0]) for n in range(2)]
    # im2col 3x3
    x = _bf(torch.randn(2 * 8 * 8, 16, generator=g))
    col = ops.im2col3x3(x, 2, 8, 8)
    xr = x.float().reshape(2, 8, 8, 16).permute(0, 3, 1, 2)
    refc = torch.nn.functional.unfold(xr, 3, padding=1).reshape(2, 16, 9, 64).permute(0, 3, 2, 1).reshape(128, 144)
    assert torch.equal(col.float(), refc)


@pytest.mark.parametrize("K", [7, 32, 64])
def test_maskpool_matches_oracle(cuda_lib, K):
    """adjoint upsample∘pool vs the reference order of operations (LISA.py:201-218,350-354)."""
    from llmseg_b200 import ops
    from oracle import selector as o_sel
    emb, segs, _ = o_sel.synthetic_case(100 + K, K)
    emb_b, segs_b = _bf(emb), _bf(segs)
    tok = emb_b[0].permute(1, 2, 0).reshape(1, 4096, 256).contiguous()
    out = ops.maskpool(segs_b, tok, torch.zeros(K, dtype=torch.int32, device=DEV))
    ref = o_sel.mask_pooling(o_sel.upsample_embeddings(emb_b.float())[0], segs_b.float())
    # fp32 up to the single bf16 store: half a bf16 ulp of the largest value (+ the bf16 rounding of the upsampled
    # embedding the reference order of operations has and the adjoint form skips)
    assert (out.float() - ref).abs().max().item() <= 1e-3 + 2 ** -8 * float(ref.abs().max())


def test_select_multi_conversation_and_sentinel(cuda_lib):
    """llmseg_select: fp32 cosine similarity of EVERY conversation of a group against the group's mask embeddings
    ([C,K] per image, reference LISA.py:397-403), predicted IoU and first-argmax from the group's first conversation,
    NaN / -1 sentinel for a conversation without [SEG]; against a torch fp32 reference of the same op."""
    from llmseg_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(11)
    Ks, convs = [5, 64, 33], [2, 1, 3]
    k_off = torch.tensor([0, 5, 69, 102], dtype=torch.int32, device=DEV)
    feat = _bf(torch.randn(102, 256, generator=g, device=DEV))
    text = _bf(torch.randn(6, 256, generator=g, device=DEV))
    h = _bf(torch.randn(102, 128, generator=g, device=DEV).relu())
    w2 = _bf(torch.randn(128, generator=g, device=DEV) * 0.1)
    b2 = torch.zeros(8, dtype=torch.bfloat16, device=DEV)
    b2[0] = 0.25
    grp = torch.tensor([0, 0, 1, 2, 2, 2], dtype=torch.int32, device=DEV)
    valid = torch.tensor([3, 9, 4, -1, 7, 8], dtype=torch.int32, device=DEV)      # conversation 3 (first of group 2): no [SEG]
    sim, iou, best = ops.select(feat, text, h, w2, b2, k_off, batch=3, k_stride=64, conv_group=grp, conv_valid=valid)
    assert sim.shape == (6, 64) and iou.shape == (3, 64) and best.shape == (3,)
    fn = torch.nn.functional.normalize(feat.float(), dim=-1)
    tn = torch.nn.functional.normalize(text.float(), dim=-1)
    r_iou = torch.sigmoid(h.float() @ w2.float() + 0.25)
    for c in range(6):
        gi = int(grp[c])
        lo, K = int(k_off[gi]), Ks[gi]
        if int(valid[c]) < 0:
            assert torch.isnan(sim[c, :K]).all()
        else:
            assert (sim[c, :K] - tn[c] @ fn[lo:lo + K].T).abs().max().item() < 2e-6
        assert torch.isinf(sim[c, K:]).all() and (sim[c, K:] < 0).all()
    for gi, c0 in enumerate((0, 2, 3)):
        lo, K = int(k_off[gi]), Ks[gi]
        if gi == 2:
            assert torch.isnan(iou[gi, :K]).all() and int(best[gi]) == -1
        else:
            assert (iou[gi, :K] - r_iou[lo:lo + K]).abs().max().item() < 2e-6
            assert int(best[gi]) == int(sim[c0, :K].to(torch.bfloat16).float().argmax())
        assert (iou[gi, K:] == 0).all()
    # identity mapping (one conversation per group) is the default
    sim1, iou1, best1 = ops.select(feat, text[[0, 2, 4]].contiguous(), h, w2, b2, k_off, batch=3, k_stride=64)
    assert torch.equal(sim1[0], sim[0]) and torch.equal(sim1[1], sim[2]) and torch.equal(sim1[2], sim[4])
    assert torch.equal(iou1[:2], iou[:2]) and not torch.isnan(iou1[2, :33]).any()


def test_small_attention_vs_fp32(cuda_lib):
    """llmseg_small_attention (8 heads x 32, <= 128 keys, ragged groups) against fp32 softmax attention on the same
    bf16 inputs: fp32 inside, one bf16 rounding at the store (reference transformer.py:319-341)."""
    from llmseg_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(12)
    Ks = [7, 128, 50]
    off = torch.tensor([0, 7, 135, 185], dtype=torch.int32, device=DEV)
    qkv = _bf(torch.randn(185, 768, generator=g, device=DEV))
    out = ops.small_attention(qkv[:, :256], qkv[:, 256:512], qkv[:, 512:], off, off, batch=3, heads=8, max_kv=128)
    for i, K in enumerate(Ks):
        lo = int(off[i])
        q, k, v = (qkv[lo:lo + K, j * 256:(j + 1) * 256].float().reshape(K, 8, 32).transpose(0, 1) for j in range(3))
        ref = (torch.softmax(q @ k.transpose(1, 2) / 32 ** 0.5, dim=-1) @ v).transpose(0, 1).reshape(K, 256)
        assert (out[lo:lo + K].float() - ref).abs().max().item() <= 2 ** -8 * float(ref.abs().max()) + 1e-4


def test_losses_match_golden(cuda_lib, golden_dir):
    """loss kernels vs the values produced by the reference's own model/loss.py (tests/golden/losses.pt)."""
    from llmseg_b200 import ops
    fx = torch.load(golden_dir / "losses.pt", weights_only=False)
    pe, te = fx["pe"], fx["te"]
    sim = ((pe / pe.norm(dim=-1, keepdim=True)) @ (te / te.norm(dim=-1, keepdim=True)).t()).flatten()
    out = ops.align_iou_loss(sim.to(DEV).contiguous(), fx["pred_ious"].flatten().to(DEV).contiguous(),
                             fx["gt_ious"].flatten().to(DEV).contiguous())
    assert abs(out[0].item() - fx["expected"]["softmax_align"]) < 1e-3
    assert abs(out[1].item() - fx["expected"]["iou_regression"]) < 1e-3
    out = ops.dice_bce_loss(fx["logits"].to(DEV).contiguous(), fx["targets"].to(DEV).contiguous(), 3.0)
    assert abs(out[0].item() - fx["expected"]["dice"]) < 1e-4
    assert abs(out[1].item() - fx["expected"]["sigmoid_ce"]) < 1e-4


def test_lm_cross_entropy_and_selector_losses(cuda_lib):
    """CE kernel (label splice by index arithmetic, ignore_index, padded vocab stride) vs F.cross_entropy on the
    oracle's spliced labels; batched align / regression losses vs the oracle's loss functions per group."""
    import torch.nn.functional as F
    from llmseg_b200 import ops
    from oracle import lisa_forward as o_lf
    g = torch.Generator(device=DEV).manual_seed(5)
    N, Tt, F_, V = 3, 12, 7, 1003
    T, ld = Tt + F_ - 1, 1008
    ids = torch.randint(3, 900, (N, Tt), generator=g, device=DEV)
    ids[:, 2] = -200
    labels = ids.clone()
    labels[:, :5] = -100
    labels[1, 9:] = -100
    logits = torch.zeros(N * T, ld, device=DEV, dtype=torch.bfloat16)
    logits[:, :V] = (torch.randn(N * T, V, generator=g, device=DEV) * 3).to(torch.bfloat16)
    logits[:, V:] = 50.0                                     # padding columns must never be read
    out2, row_loss = ops.lm_cross_entropy(logits, ids, labels, n_img_tokens=F_, vocab=V, image_token=-200)
    new_labels = o_lf.splice_labels(ids, labels, F_)
    lg = logits[:, :V].float().view(N, T, V)
    ref = F.cross_entropy(lg[:, :-1].reshape(-1, V), new_labels[:, 1:].reshape(-1))
    n_tgt = int((new_labels[:, 1:] != -100).sum())
    assert int(out2[1]) == n_tgt and n_tgt > 0
    assert abs(float(out2[0]) - float(ref)) < 1e-4 * abs(float(ref))
    assert int((row_loss >= 0).sum()) == n_tgt
    # batched selector losses: 3 groups with ragged K
    Ks, ks = [9, 64, 33], 64
    k_off = torch.tensor([0, 9, 73, 106], dtype=torch.int32, device=DEV)
    sim = torch.rand(3, ks, generator=g, device=DEV) * 2 - 1
    piou, giou, giop = (torch.rand(3, ks, generator=g, device=DEV) for _ in range(3))
    gw = torch.tensor([0.25, 0.25, 0.5], device=DEV)
    ce = torch.tensor([1.25, 7.0], device=DEV)
    out4, per = ops.selector_losses(sim, piou, giou, giop, k_off, gw, ce=ce, weights=(2.0, 0.5, 3.0))
    a_ref = r_ref = 0.0
    for i, K in enumerate(Ks):
        s_, g_ = sim[i, :K, None] / 0.05, giou[i, :K, None] / 0.05
        kl = F.kl_div(F.log_softmax(s_, 0), F.softmax(g_, 0), reduction="sum")
        mse = o_lf.iou_regression_loss(piou[i, :K], giop[i, :K])
        assert abs(float(per[i, 0]) - float(kl)) < 1e-4 * (1 + abs(float(kl)))
        assert abs(float(per[i, 1]) - float(mse)) < 1e-4 * (1 + abs(float(mse)))
        a_ref += float(gw[i]) * float(kl)
        r_ref += float(gw[i]) * float(mse)
    exp = [2.5 + 0.5 * a_ref + 3.0 * r_ref, 2.5, 0.5 * a_ref, 3.0 * r_ref]
    for i in range(4):
        assert abs(float(out4[i]) - exp[i]) < 1e-4 * (1 + abs(exp[i]))


def test_ops_reject_bad_inputs(cuda_lib):
    from llmseg_b200 import ops
    a = torch.zeros(8, 12, dtype=torch.bfloat16, device=DEV)      # K % 8 != 0
    with pytest.raises(RuntimeError):
        ops.gemm(a, a)
    with pytest.raises(TypeError):
        ops.gemm(a.float(), a.float())
    with pytest.raises(RuntimeError):
        ops.gemm(a.cpu(), a.cpu())
