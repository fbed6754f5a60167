set -x
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
import bench
open('/tmp/enwik.bin','wb').write(bench.gen_workload('enwik100m').tobytes())
open('/tmp/js48k.bin','wb').write(bench.gen_workload('js48k').tobytes())
"
( for m in 1 2 1 2; do ( time ZULTRA_CLI_EXIT=$m ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done
  for m in 1 2; do ( time ZULTRA_CLI_EXIT=$m ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -zlib /tmp/js48k.bin /tmp/js48k.z ) ; done ) > gpurun_out/r2_cli20.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest20.log 2>&1
tail -5 gpurun_out/r2_pytest20.log
timeout 600 python bench.py > gpurun_out/r2_bench20_n1.json 2> gpurun_out/r2_bench20_n1.err
tail -c 600 gpurun_out/r2_bench20_n1.json
