#!/usr/bin/env python
"""Config 5 of BASELINE.json: a batch of independent 4-64 KB payloads, one zlib stream each
(zultra_cuda_memory_compress_batch).  Prints MB/s, per-stage ms and checks every stream inflates to its payload;
the first few are compared with the compiled reference when it travelled (oracle/_ref)."""
import argparse
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import zultra_b200 as z
    from zultra_b200 import synth
    payloads = synth.batch(args.count)
    total = sum(len(p) for p in payloads)
    ctx = z.CudaCtx()
    outs = ctx.memory_compress_batch(payloads, 1)       # warm-up (allocations)
    for p, o in zip(payloads, outs):
        assert o is not None and zlib.decompress(o) == p.tobytes()
    import refharness
    checked = 0
    if os.path.exists(refharness.REF_SO):
        ref = refharness.Ref()
        for p, o in list(zip(payloads, outs))[:200]:
            assert o == ref.compress(p, flags=1)
            checked += 1
    import ctypes as C
    L = z.load()
    L.zultra_cuda_profile.argtypes = [C.c_int]
    L.zultra_cuda_profile_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        ctx.memory_compress_batch(payloads, 1)
        ts.append(time.perf_counter() - t0)
    dt = min(ts)
    # one more pass with per-kernel CUDA events (not timed above)
    L.zultra_cuda_profile(1)
    ctx.memory_compress_batch(payloads, 1)
    L.zultra_cuda_profile(0)
    names = C.create_string_buffer(32 * 256); kms = (C.c_float * 256)(); kcnt = (C.c_int * 256)()
    nk = L.zultra_cuda_profile_collect(names, kms, kcnt, 256)
    ktab = sorted([(names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), round(kms[i], 2)) for i in range(nk)], key=lambda r: -r[1])[:24]
    print(json.dumps({"workload": "batch", "payloads": args.count, "bytes": total, "MB/s_host_to_host": round(total / dt / 1e6, 2), "ms": round(dt * 1e3, 2),
                      "stages_ms": {k: round(v, 2) for k, v in ctx.timings().items()}, "counters": ctx.counters(), "kernels_ms": dict(ktab),
                      "verified": "all inflate; first %d equal the compiled reference" % checked}))
    ctx.close()


if __name__ == "__main__":
    main()
