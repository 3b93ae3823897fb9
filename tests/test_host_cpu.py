"""CPU: host-side logic of the product that needs no GPU — checkpoint key handling, the weight-side half of
the folded norms, the bench.py contract of the CPU reference arm."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_strip_peft_prefix_and_lora_merge():
    """Checkpoints saved through PEFT (reference training.py:194-237): `base_model.model.` prefix and unmerged
    LoRA q/v deltas W + (alpha / r) * B @ A with r = 8, alpha = 16."""
    from llmseg_b200.lisa import strip_peft_prefix
    g = torch.Generator().manual_seed(0)
    w = torch.randn(16, 16, generator=g)
    a, b = torch.randn(8, 16, generator=g), torch.randn(16, 8, generator=g)
    sd = {
        "base_model.model.model.layers.0.self_attn.q_proj.weight": w,
        "base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight": a,
        "base_model.model.model.layers.0.self_attn.q_proj.lora_B.default.weight": b,
        "base_model.model.model.norm.weight": torch.ones(16),
        "model.embed_tokens.weight": torch.zeros(4, 16),
    }
    out = strip_peft_prefix(sd)
    assert set(out) == {"model.layers.0.self_attn.q_proj.weight", "model.norm.weight", "model.embed_tokens.weight"}
    assert torch.allclose(out["model.layers.0.self_attn.q_proj.weight"], w + 2.0 * (b @ a), atol=1e-6)
    # a plain state dict passes through untouched
    plain = {"model.norm.weight": torch.ones(3)}
    assert strip_peft_prefix(plain)["model.norm.weight"] is plain["model.norm.weight"]


@pytest.mark.parametrize("rms", [False, True])
def test_fold_norm_identity(rms):
    """ops.fold_norm: Norm(x) @ W.T + b == rstd * (x @ W''.T) + b' with the row statistics applied in the GEMM
    epilogue — the algebra behind every folded LayerNorm / RMSNorm of the encoders (fp32 check of the identity;
    the bf16 rounding of W'' is the only approximation and is covered by the GPU tests)."""
    from llmseg_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 64, generator=g) * 2 + 0.5
    w, bias = torch.randn(24, 64, generator=g) / 8, torch.randn(24, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g)
    eps = 1e-6
    if rms:
        rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
        ref = (x * rstd * gamma) @ w.T + bias
        w2, b2 = ops.fold_norm(w.double(), gamma.double(), None, bias.double(), rms=True)
    else:
        ref = torch.nn.functional.layer_norm(x, (64,), gamma, beta, eps) @ w.T + bias
        rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
        w2, b2 = ops.fold_norm(w.double(), gamma.double(), beta.double(), bias.double(), rms=False)
    got = rstd * (x @ w2.float().T) + b2.float()
    # w2 / b2 come back in bf16: compare against the same identity evaluated with the rounded operands' error bound
    assert (got - ref).abs().max().item() < 0.08 and (got - ref).abs().mean().item() < 0.02
    # exact identity in fp64 without the bf16 rounding (mirrors fold_norm's arithmetic)
    wf = w.double() * gamma.double()[None, :]
    if not rms:
        wf = wf - wf.mean(1, keepdim=True)
    bf = bias.double() + (0 if rms else w.double() @ beta.double())
    exact = rstd.double() * (x.double() @ wf.T) + bf
    assert (exact - ref.double()).abs().max().item() < 1e-5


def test_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver launches next to ours): one JSON line with the same
    metric / unit as the GPU arm, `impl`, a `cpu_baseline` describing the run and an `e2e` repeating the value."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "images/sec fwd 1024px+64tok" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without output."""
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=300, cwd=str(ROOT), env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_select_proposals_rules():
    """training.py:627-629 (argmax of the similarity) and :712-718 (predicted IoU above a threshold), on a dict
    shaped like the forward's output, with and without the fused `best_index`."""
    from llmseg_b200.lisa import select_proposals
    out = {"pred_similarity": [torch.tensor([[0.1, 0.7, 0.3]]), torch.tensor([[0.9, -0.2]])],
           "pred_iou": [torch.tensor([[0.6, 0.4, 0.51]], dtype=torch.bfloat16), torch.tensor([[0.2, 0.1]], dtype=torch.bfloat16)]}
    assert select_proposals(out) == [(1, None), (0, None)]
    assert select_proposals(out, threshold=0.5) == [(1, [0, 2]), (0, [])]
    out["best_index"] = torch.tensor([1, 0], dtype=torch.int32)
    assert select_proposals(out, threshold=0.5) == [(1, [0, 2]), (0, [])]


def test_plan_buckets_and_lru():
    """Host logic of the plan cache (no GPU): prompt lengths bucket to 32 tokens, proposal counts to 32 / 64 / 128,
    the per-stage LRU keeps `cap` plans and evicts the least recently used."""
    from llmseg_b200 import lisa
    assert [lisa.bucket_tokens(t) for t in (1, 32, 33, 64, 65, 512)] == [32, 32, 64, 64, 96, 512]
    assert [lisa.bucket_props(k) for k in (1, 32, 33, 50, 64, 65, 128)] == [32, 32, 64, 64, 64, 128, 128]
    with pytest.raises(ValueError):
        lisa.bucket_props(129)
    lru = lisa._LRU(2)
    lru.put("a", 1)
    lru.put("b", 2)
    assert lru.get("a") == 1          # refreshes "a"
    lru.put("c", 3)                    # evicts "b"
    assert lru.get("b") is None and lru.get("a") == 1 and lru.get("c") == 3 and len(lru) == 2


def test_reference_constructor_surface_without_weights():
    """`LISAForCausalLM(config, **kwargs)` (reference model/LISA.py:144-170) is an nn.Module before any weights exist:
    config mapping, eval/train, no-op device moves, a clear error on forward, LoRA merge rules."""
    from llmseg_b200 import lisa
    hf = {"hidden_size": 4096, "num_hidden_layers": 3, "num_attention_heads": 32, "intermediate_size": 11008,
          "vocab_size": 32003, "rms_norm_eps": 1e-5, "mm_vision_select_layer": -2}
    m = lisa.LISAForCausalLM(hf, seg_token_idx=32001, train_mask_decoder=True, out_dim=256, vision_pretrained=None,
                             vision_tower="openai/clip-vit-large-patch14", use_mm_start_end=True)
    assert isinstance(m, torch.nn.Module) and m.cfg.llama.layers == 3 and m.cfg.llama.eps == 1e-5 and m.cfg.seg_token_idx == 32001
    assert m.training and not m.eval().training and m.to("cpu") is m and m.bfloat16() is m and m.get_model() is m
    assert list(m.parameters()) == []
    with pytest.raises(RuntimeError):
        m(images=None)
    with pytest.raises(RuntimeError):
        m.state_dict()
    with pytest.raises(ValueError):
        lisa.cfg_from_hf_config(dict(hf, num_key_value_heads=8))
    with pytest.raises(ValueError):
        lisa.cfg_from_hf_config(hf, out_dim=128)
    # LoRA pairs: merged with the caller's alpha; a pair without its base weight or its partner is an error
    w, a, b = torch.randn(8, 6), torch.randn(2, 6), torch.randn(8, 2)
    sd = {"base_model.model.x.q_proj.weight": w, "base_model.model.x.q_proj.lora_A.default.weight": a,
          "base_model.model.x.q_proj.lora_B.default.weight": b, "base_model.model.y.weight": torch.ones(1)}
    out = lisa.strip_peft_prefix(sd, lora_alpha=32.0)
    assert set(out) == {"x.q_proj.weight", "y.weight"}
    assert torch.allclose(out["x.q_proj.weight"], w + (32.0 / 2) * (b @ a), atol=1e-5)
    with pytest.raises(KeyError):
        lisa.strip_peft_prefix({k: v for k, v in sd.items() if not k.endswith("q_proj.weight")})
    with pytest.raises(KeyError):
        lisa.strip_peft_prefix({k: v for k, v in sd.items() if "lora_B" not in k})


def test_proposal_point_grid_matches_oracle():
    """The generator's point grid (reference utils/amg.py:179-186) equals the oracle's."""
    from llmseg_b200 import proposals
    from oracle import sam_amg
    for n in (1, 8, 32):
        assert torch.allclose(torch.from_numpy(proposals.point_grid(n)), sam_amg.point_grid(n), atol=1e-4)
