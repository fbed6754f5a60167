set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest26.log 2>&1
tail -3 gpurun_out/r2_pytest26.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe26.jsonl > gpurun_out/r2_probe26.log 2>&1
ZULTRA_CUDA_LANES=2 timeout 600 python tools/gpu_probe.py enwik100m mozilla51m mix256m --out gpurun_out/r2_probe26_lanes2.jsonl > /dev/null 2>&1
ZULTRA_CUDA_DP_VAR=9 timeout 600 python tools/gpu_probe.py mozilla51m mix256m batch10k --out gpurun_out/r2_probe26_var9.jsonl > /dev/null 2>&1
