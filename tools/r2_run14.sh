mkdir -p gpurun_out
ZULTRA_CUDA_FIX_DEBUG=1 python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, bench, zultra_b200 as z
d = np.ascontiguousarray(bench.gen_workload('mozilla51m'))
c = z.CudaCtx(); c.compress_blocks(d, finalize=1, flags=2); c.close()
" > gpurun_out/r2_fixdbg_moz14.txt 2>&1
