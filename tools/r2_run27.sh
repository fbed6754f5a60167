set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zconfigs.py -m gpu -x -q -k "not mix1g and not batch100k" ) > gpurun_out/r2_pytest27.log 2>&1
tail -3 gpurun_out/r2_pytest27.log
timeout 600 python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k --out gpurun_out/r2_probe27.jsonl > gpurun_out/r2_probe27.log 2>&1
ZULTRA_CUDA_FIX_DEBUG=1 python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, bench, zultra_b200 as z
from zultra_b200 import synth
d = synth.mix(256 << 20)
c = z.CudaCtx(); c.compress_blocks(d, finalize=1, flags=2); c.close()
" > gpurun_out/r2_fixdbg_mix27.txt 2>&1
