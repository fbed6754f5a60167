#!/usr/bin/env python
"""One warm-up call, then one call of a workload between cudaProfilerStart/Stop (run under
`ncu --profile-from-start off -k regex:<kernels> ...`).  usage: ncu_one.py <workload> [flags]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    name = sys.argv[1]
    import bench
    import zultra_b200 as z
    from zultra_b200 import synth
    if name == "js48k":
        data, flags = synth.js48k(), 1
    elif name.startswith("batch"):
        data, flags = synth.batch(int(name[5:].replace("k", "000"))), 1
    else:
        data, flags = np.ascontiguousarray(bench.gen_workload(name)), bench.WORKLOADS[name]["flags"]
    torch.zeros(1, device="cuda")
    ctx = z.CudaCtx()
    run = (lambda: ctx.memory_compress_batch(data, 1)) if isinstance(data, list) else (lambda: ctx.compress_blocks(data, finalize=1, flags=flags))
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.close()


if __name__ == "__main__":
    main()
