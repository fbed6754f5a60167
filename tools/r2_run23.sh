set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest23.log 2>&1
tail -3 gpurun_out/r2_pytest23.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe23.jsonl > gpurun_out/r2_probe23.log 2>&1
ZULTRA_CUDA_PARSE_AWU=0 timeout 600 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe23_awu0.jsonl > /dev/null 2>&1
for wu in 352 320 290; do
  ZULTRA_CUDA_PARSE_WU=$wu timeout 600 python tools/gpu_probe.py enwik100m mozilla51m batch10k --out gpurun_out/r2_probe23_wu$wu.jsonl > /dev/null 2>&1
done
