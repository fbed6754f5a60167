set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest24.log 2>&1
tail -3 gpurun_out/r2_pytest24.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m batch10k --out gpurun_out/r2_probe24.jsonl > gpurun_out/r2_probe24.log 2>&1
for cd in 384 512 576 640 832 1024; do
  ZULTRA_CUDA_PARSE_CD=$cd timeout 600 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe24_cd$cd.jsonl > /dev/null 2>&1
done
for v in 3 5 6; do
  ZULTRA_CUDA_DP_VAR=$v timeout 600 python tools/gpu_probe.py enwik100m --out gpurun_out/r2_probe24_var$v.jsonl > /dev/null 2>&1
done
