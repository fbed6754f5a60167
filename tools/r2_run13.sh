set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest13.log 2>&1
tail -4 gpurun_out/r2_pytest13.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe13.jsonl > gpurun_out/r2_probe13.log 2>&1
ZULTRA_CUDA_MT_THREADS=256 timeout 300 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe13_mt256.jsonl > /dev/null 2>&1
timeout 900 python tools/bench_batch.py --count 10000 --out gpurun_out/r2_batch10k.jsonl > gpurun_out/r2_batch10k.log 2>&1
tail -2 gpurun_out/r2_batch10k.log
