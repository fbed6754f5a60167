set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest19.log 2>&1
tail -5 gpurun_out/r2_pytest19.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe19.jsonl > gpurun_out/r2_probe19.log 2>&1
ZULTRA_CUDA_SORT_LEGACY=1 timeout 300 python tools/gpu_probe.py enwik100m mozilla51m --out gpurun_out/r2_probe19_legacy.jsonl > /dev/null 2>&1
python -c "
import sys; sys.path.insert(0,'.')
import bench
open('/tmp/enwik.bin','wb').write(bench.gen_workload('enwik100m').tobytes())
"
( for m in 0 1 2; do ( time ZULTRA_CLI_EXIT=$m ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done ) > gpurun_out/r2_cli19_enwik.txt 2>&1
( for m in 1 0 2; do ( time ZULTRA_CLI_EXIT=$m ZULTRA_CUDA_TRACE=1 ./zultra_b200/zultra -gzip /tmp/enwik.bin /tmp/enwik.gz ) ; done ) >> gpurun_out/r2_cli19_enwik.txt 2>&1
python - <<'PY' >> gpurun_out/r2_cli19_enwik.txt 2>&1
import gzip,hashlib
d=gzip.open('/tmp/enwik.gz','rb').read(); print('roundtrip', len(d), hashlib.sha256(d).hexdigest()==hashlib.sha256(open('/tmp/enwik.bin','rb').read()).hexdigest())
PY
