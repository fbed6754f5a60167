set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py tests/test_gpu_multi.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest7.log 2>&1
tail -8 gpurun_out/r2_pytest7.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe7.jsonl > gpurun_out/r2_probe7.log 2>&1
