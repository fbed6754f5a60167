set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest17.log 2>&1
tail -5 gpurun_out/r2_pytest17.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mozilla:6400000 mozilla:25600000 mix256m --out gpurun_out/r2_probe17.jsonl > gpurun_out/r2_probe17.log 2>&1
python -c "
import sys; sys.path.insert(0,'.')
import bench
open('/tmp/enwik.bin','wb').write(bench.gen_workload('enwik100m').tobytes())
"
( for i in 1 2; do ( time ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done; ( time ZULTRA_CUDA_DEVICES=2 ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik2.gz ); cmp /tmp/enwik.gz /tmp/enwik2.gz && echo same ) > gpurun_out/r2_cli17_enwik.txt 2>&1
