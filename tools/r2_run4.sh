set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest4.log 2>&1
tail -5 gpurun_out/r2_pytest4.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m --out gpurun_out/r2_probe4.jsonl > gpurun_out/r2_probe4.log 2>&1
for l in 2 3; do ZULTRA_CUDA_LANES=$l timeout 300 python tools/gpu_probe.py mozilla51m enwik100m --out gpurun_out/r2_probe4_lanes$l.jsonl > /dev/null 2>&1; done
for t in "8 6" "12 3" "16 8" "24 6"; do set -- $t; ZULTRA_CUDA_TS_MIN=$1 ZULTRA_CUDA_TS_MUL=$2 timeout 300 python tools/gpu_probe.py mozilla51m enwik100m --out gpurun_out/r2_probe4_ts$1_$2.jsonl > /dev/null 2>&1; done
