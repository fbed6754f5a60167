set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe25.jsonl > gpurun_out/r2_probe25.log 2>&1
for v in 8 9 10; do
  ZULTRA_CUDA_DP_VAR=$v timeout 600 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe25_var$v.jsonl > /dev/null 2>&1
done
for cd in 768 832 960; do
  ZULTRA_CUDA_PARSE_CD=$cd timeout 600 python tools/gpu_probe.py enwik100m --out gpurun_out/r2_probe25_cd$cd.jsonl > /dev/null 2>&1
done
for cd in 448 512; do
  ZULTRA_CUDA_PARSE_CD=$cd timeout 600 python tools/gpu_probe.py mozilla51m --out gpurun_out/r2_probe25_cd$cd.jsonl > /dev/null 2>&1
done
