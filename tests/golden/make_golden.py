"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference compiled under oracle/_ref.

Run in the build container (needs /root/reference): python tests/golden/make_golden.py
The vectors pin (1) the oracle restatement and (2) the CUDA path on boxes where the reference is absent.
Inputs are regenerated from seeds (tests/cases.py); only reference OUTPUTS are stored:
  <case>/<fmt>  sha256 + length of zultra_memory_compress output for deflate/zlib/gzip
  stage dumps (packed SA|LCP words, match lists, split offsets, code lengths) for two small windows.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
from refharness import Ref  # noqa: E402


def main():
    ref = Ref()
    out = {}
    for name, data in cases.small_cases().items():
        for flags, fmt in ((0, "deflate"), (1, "zlib"), (2, "gzip")):
            r = ref.compress(data, flags=flags)
            out["%s/%s" % (name, fmt)] = np.frombuffer(hashlib.sha256(r).digest() + len(r).to_bytes(8, "little"), dtype=np.uint8)
    for name, data, block in cases.multi_block_cases():
        r = ref.compress(data, flags=2, block=block)
        out["%s/gzip" % name] = np.frombuffer(hashlib.sha256(r).digest() + len(r).to_bytes(8, "little"), dtype=np.uint8)
    # stage dumps
    from zultra_b200 import synth
    w1 = synth.js48k()[:12000]
    w2 = synth.mozilla(60000, seed=77)   # 32768 history + 27232 block bytes
    for tag, win, hist in (("stage_js12k", w1, 0), ("stage_moz60k_h32k", w2, 32768)):
        out[tag + "/sa_lcp"] = ref.sa_lcp(win)
        out[tag + "/match"] = ref.matches(win, hist)
        st = ref.block_stages(win, hist)
        out[tag + "/split"] = st["split"]; out[tag + "/dyn"] = st["dyn"]; out[tag + "/ll"] = st["ll"].astype(np.uint8); out[tag + "/ol"] = st["ol"].astype(np.uint8)
        out[tag + "/bits"] = st["bits"]; out[tag + "/best"] = st["best"][hist:]
    dic = synth.enwik(20000, seed=31)
    body = synth.enwik(90000, seed=32)
    r = ref.compress_dict(body, dic, flags=1)
    out["dict/zlib"] = np.frombuffer(hashlib.sha256(r).digest() + len(r).to_bytes(8, "little"), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", len(out), "entries")


if __name__ == "__main__":
    main()
