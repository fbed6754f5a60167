#!/usr/bin/env python
"""Small inputs through every path of the pipeline, to be run under compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py [which]"""
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    import zultra_b200 as z
    from zultra_b200 import synth
    rng = np.random.default_rng(5)
    if which in ("all", "js"):
        d = synth.js48k()
        out = z.memory_compress(d, 1)
        assert zlib.decompress(out) == d.tobytes()
        print("js48k ok", len(out))
    if which in ("all", "mix"):
        d = np.concatenate([synth.enwik(600000, seed=3), np.zeros(150000, dtype=np.uint8), synth.mozilla(700000, seed=4),
                            np.tile(np.arange(7, dtype=np.uint8), 30000), rng.integers(0, 256, size=100000).astype(np.uint8)])
        c = z.CudaCtx()
        got, bits, ck = c.compress_blocks(d, block=262144, finalize=1, flags=2)
        c.close()
        assert zlib.decompress(got, -15) == d.tobytes() and ck == zlib.crc32(d.tobytes())
        print("mix ok", len(got))
    if which in ("all", "batch"):
        pay = synth.batch(40)
        outs = z.memory_compress_batch(pay, 1)
        for p, o in zip(pay, outs):
            assert zlib.decompress(bytes(o)) == bytes(p)
        print("batch ok", len(outs))


if __name__ == "__main__":
    main()
