set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "norm" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_proposals.py tests/test_gpu_e2e.py -x -q -m gpu -k "proposal or mask_decoder or sam_encoder or smoke or forward_small" 2>&1 | tail -3
timeout 300 python scripts/gpu_proposals_time.py 256 2>&1 | head -12
for k in amg_tok2img amg_upscale amg_img2tok; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tok2img_attn_mma|upscale_logits|img2tok_attn" -s 1 -c 1 -f -o gpurun_out/r2r_$k python scripts/profile_kernels.py $k 8 3 > gpurun_out/r2r_ncu_$k.log 2>&1; echo exit=$?; tail -1 gpurun_out/r2r_ncu_$k.log
done
