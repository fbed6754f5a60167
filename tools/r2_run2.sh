set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest2.log 2>&1
tail -5 gpurun_out/r2_pytest2.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m batch10k --out gpurun_out/r2_probe2.jsonl > gpurun_out/r2_probe2.log 2>&1
for cd in 512 1024 1536; do ZULTRA_CUDA_PARSE_CD=$cd timeout 300 python tools/gpu_probe.py enwik100m --out gpurun_out/r2_probe2_cd$cd.jsonl > /dev/null 2>&1; done
for wu in 320 448; do ZULTRA_CUDA_PARSE_WU=$wu timeout 300 python tools/gpu_probe.py enwik100m --out gpurun_out/r2_probe2_wu$wu.jsonl > /dev/null 2>&1; done
