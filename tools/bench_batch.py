#!/usr/bin/env python
"""Config 5 of BASELINE.json: a batch of independent 4-64 KB payloads (PNG-IDAT / HTTP-body shaped), one zlib stream each,
through zultra_cuda_memory_compress_batch.  Host buffers in, host buffers out; the argument arrays and the output buffers are
built ONCE (a C caller keeps its buffers too) and only the C call is timed.  Reports MB/s for pageable and page-locked
caller memory, per-stage ms, and checks EVERY stream against the unmodified reference: the sha256 prefixes of
tests/golden/config_golden.npz when the batch is the configuration's own 100 000 payloads, else oracle/_ref run here."""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


class BatchArgs:
    """ctypes argument block over ONE input slab and ONE output slab (pageable numpy or page-locked torch memory)."""

    def __init__(self, L, payloads, flags, pinned):
        n = len(payloads)
        sizes = np.array([len(p) for p in payloads], dtype=np.int64)
        caps = np.array([L.zultra_memory_bound(int(s), flags, 0) for s in sizes], dtype=np.int64)
        ioff = np.concatenate(([0], np.cumsum(sizes)))
        ooff = np.concatenate(([0], np.cumsum(caps)))
        if pinned:
            import torch
            self._ti = torch.empty(int(ioff[-1]), dtype=torch.uint8).pin_memory()
            self._to = torch.empty(int(ooff[-1]), dtype=torch.uint8).pin_memory()
            self.inp, self.out = self._ti.numpy(), self._to.numpy()
        else:
            self.inp, self.out = np.empty(int(ioff[-1]), dtype=np.uint8), np.zeros(int(ooff[-1]), dtype=np.uint8)
        for p, o in zip(payloads, ioff[:-1]):
            self.inp[o:o + len(p)] = p
        ib, ob = self.inp.ctypes.data, self.out.ctypes.data
        self.n, self.sizes, self.ooff = n, sizes, ooff
        self.in_ptrs = (C.c_void_p * n)(*[ib + int(o) for o in ioff[:-1]])
        self.in_sizes = (C.c_size_t * n)(*[int(s) for s in sizes])
        self.out_ptrs = (C.c_void_p * n)(*[ob + int(o) for o in ooff[:-1]])
        self.out_caps = (C.c_size_t * n)(*[int(c) for c in caps])
        self.out_sizes = (C.c_size_t * n)()

    def stream(self, i):
        return self.out[int(self.ooff[i]):int(self.ooff[i]) + int(self.out_sizes[i])].tobytes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import zultra_b200 as z
    from zultra_b200 import synth
    L = z.load()
    t0 = time.time()
    payloads = synth.batch(args.count)
    total = sum(len(p) for p in payloads)
    gen_s = time.time() - t0
    ctx = z.CudaCtx()
    res = {"workload": "batch (config 5)", "payloads": args.count, "bytes": total, "format": "zlib", "generate_s": round(gen_s, 1)}
    for mode in ("pageable", "pinned"):
        a = BatchArgs(L, payloads, 1, mode == "pinned")
        ts = []
        for it in range(1 + args.steps):      # first call: the context sizes its device buffers
            t0 = time.perf_counter()
            rc = L.zultra_cuda_memory_compress_batch(ctx.p, a.in_ptrs, a.in_sizes, a.out_ptrs, a.out_caps, a.out_sizes, a.n, 1, 0)
            assert rc == 0, rc
            if it:
                ts.append(time.perf_counter() - t0)
        res["MB/s_host_to_host_" + mode] = round(total / min(ts) / 1e6, 1)
        res["ms_" + mode] = round(1e3 * min(ts), 1)
        if mode == "pageable":
            res["stages_ms"] = {k: round(v, 1) for k, v in ctx.timings().items()}
            res["counters"] = ctx.counters()
            # parity: every stream against the reference
            gpath = os.path.join(ROOT, "tests", "golden", "config_golden.npz")
            g = np.load(gpath) if os.path.exists(gpath) else None
            if args.count == 100000 and g is not None and "batch100k/out_sha8" in g:
                want, wlen = g["batch100k/out_sha8"], g["batch100k/out_len"]
                for i in range(a.n):
                    s = a.stream(i)
                    assert len(s) == int(wlen[i]) and hashlib.sha256(s).digest()[:8] == want[i].tobytes(), "stream %d differs from the reference" % i
                res["verified"] = "all %d streams == unmodified reference (sha256 prefixes, tests/golden/config_golden.npz)" % a.n
            else:
                import refharness
                k = min(a.n, 500)
                if os.path.exists(refharness.REF_SO):
                    ref = refharness.Ref()
                    for i in range(k):
                        assert a.stream(i) == ref.compress(payloads[i], flags=1), i
                    res["verified"] = "first %d streams == oracle/_ref in this run" % k
            res["compressed_bytes"] = int(sum(int(x) for x in a.out_sizes))
        del a
    print(json.dumps(res))
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(res) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
