set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py tests/test_gpu_multi.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest10.log 2>&1
tail -8 gpurun_out/r2_pytest10.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe10.jsonl > gpurun_out/r2_probe10.log 2>&1
python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, bench, zultra_b200 as z
d = np.ascontiguousarray(bench.gen_workload('mozilla51m'))
c = z.CudaCtx(); c.compress_blocks(d, finalize=1, flags=2); c.close()
" > gpurun_out/r2_fixdbg_moz10.txt 2>&1
