#!/usr/bin/env python
"""Per-stage / per-kernel timing of the configurations on one GPU (development aid; writes JSON lines).
  python tools/gpu_probe.py js48k enwik100m mozilla51m mix256m batch10k  [--out gpurun_out/probe.jsonl]
Every workload: 2 warm-up calls, then best-of-3 wall time through the public one-shot API (host to host), then one
profiled call with per-kernel CUDA events."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def kernel_table(L, top=40):
    names = C.create_string_buffer(32 * 256); kms = (C.c_float * 256)(); kcnt = (C.c_int * 256)()
    nk = L.zultra_cuda_profile_collect(names, kms, kcnt, 256)
    tab = sorted([(names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), round(kms[i], 3), kcnt[i]) for i in range(nk)], key=lambda r: -r[1])
    return {r[0]: [r[1], r[2]] for r in tab[:top]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*", default=["js48k", "enwik100m", "mozilla51m"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.jsonl"))
    args = ap.parse_args()
    import bench
    import zultra_b200 as z
    from zultra_b200 import synth
    L = z.load()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fo = open(args.out, "a")
    for name in args.names:
        if name == "js48k":
            data, flags = synth.js48k(), 1
        elif name.startswith("batch"):
            data, flags = None, 1
        elif name == "mix256m":
            data, flags = synth.mix(256 << 20), 2
        elif ":" in name:      # generator:bytes, e.g. mozilla:6400000 (a GPU's share of a strongly scaled configuration)
            g, n = name.split(":")
            data, flags = getattr(synth, g)(int(n)), 2
        else:
            data, flags = np.ascontiguousarray(bench.gen_workload(name)), bench.WORKLOADS[name]["flags"]
        ctx = z.CudaCtx()
        rec = {"workload": name}
        if data is None:
            pay = synth.batch(int(name[5:].replace("k", "000")))
            total = sum(len(p) for p in pay)
            run = lambda: ctx.memory_compress_batch(pay, 1)
        else:
            total = len(data)
            run = lambda: ctx.compress_blocks(data, finalize=1, flags=flags) if total <= (256 << 20) else z.memory_compress(data, flags)
        r0 = run(); run()
        import hashlib
        first = r0[0] if isinstance(r0, tuple) else (b"".join(bytes(x) for x in r0) if isinstance(r0, list) else bytes(r0))
        rec["out_sha"] = hashlib.sha256(first).hexdigest()[:16]      # same for every build variant of the library, or the variant is wrong
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
        rec.update(bytes=total, best_ms=round(1e3 * min(ts), 3), MBps=round(total / min(ts) / 1e6, 1), stages_ms={k: round(v, 3) for k, v in ctx.timings().items()},
                   counters=ctx.counters())
        L.zultra_cuda_profile(1); run(); L.zultra_cuda_profile(0)
        rec["kernels_ms"] = kernel_table(L)
        print(json.dumps(rec), flush=True)
        fo.write(json.dumps(rec) + "\n"); fo.flush()
        ctx.close()
        L.zultra_cuda_release_cached()


if __name__ == "__main__":
    main()
