set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | head -3 > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest1.log 2>&1
tail -5 gpurun_out/r2_pytest1.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe1.jsonl > gpurun_out/r2_probe1.log 2>&1
ZULTRA_CUDA_DP_SM=1 timeout 300 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe1_dpsm.jsonl > gpurun_out/r2_probe1_dpsm.log 2>&1
ZULTRA_CUDA_DP_SM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_cases or multi_block or runs_and or large_vs" > gpurun_out/r2_pytest1_dpsm.log 2>&1
python -c "
import sys; sys.path.insert(0,'.')
from zultra_b200 import synth
import bench
open('/tmp/js.bin','wb').write(synth.js48k().tobytes())
open('/tmp/enwik.bin','wb').write(bench.gen_workload('enwik100m').tobytes())
"
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -zlib /tmp/js.bin /tmp/js.z ) ; done ) > gpurun_out/r2_cli_js48k.txt 2>&1
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done ) > gpurun_out/r2_cli_enwik.txt 2>&1
( time CUDA_VISIBLE_DEVICES=0 ./zultra_b200/zultra -zlib /tmp/js.bin /tmp/js.z ) >> gpurun_out/r2_cli_js48k.txt 2>&1
( time ./oracle/_ref/zultra_ref -zlib /tmp/js.bin /tmp/js_ref.z ) >> gpurun_out/r2_cli_js48k.txt 2>&1
cmp /tmp/js.z /tmp/js_ref.z && echo same >> gpurun_out/r2_cli_js48k.txt
( time timeout 900 python bench.py ) > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 600 gpurun_out/r2_bench1.err
