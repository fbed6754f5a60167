"""Seeded parity inputs shared by the golden generator, the CPU tests and the GPU tests."""
import numpy as np

from zultra_b200 import synth


def small_cases():
    """name -> uint8 array; every case finishes in well under a second on the reference."""
    rng = np.random.default_rng(7)
    c = {}
    c["js48k"] = synth.js48k()
    c["enwik200k"] = synth.enwik(200000)
    c["moz300k"] = synth.mozilla(300000)
    c["zeros100k"] = np.zeros(100000, dtype=np.uint8)
    c["period7"] = np.tile(np.frombuffer(b"abcdefg", dtype=np.uint8), 20000)
    c["random64k"] = rng.integers(0, 256, size=65536).astype(np.uint8)
    c["alpha2"] = rng.integers(0, 2, size=100000).astype(np.uint8)
    c["alpha3"] = rng.integers(0, 3, size=60000).astype(np.uint8)
    c["one"] = np.array([65], dtype=np.uint8)
    c["abc"] = np.frombuffer(b"abc", dtype=np.uint8).copy()
    c["aaaaa"] = np.frombuffer(b"aaaaa", dtype=np.uint8).copy()
    for a, p in [(1, 0.5), (2, 0.9), (15, 0.5), (96, 0.99), (256, 0.0), (256, 0.995)]:
        c["lz_a%d_p%g" % (a, p)] = synth.lz_selftest(40000, 123 + a, a, p)
    c["rows1000"] = np.tile(rng.integers(0, 256, size=1000).astype(np.uint8), 90)
    c["utf16"] = np.stack([synth.enwik(40000, seed=5), np.zeros(40000, dtype=np.uint8)], axis=1).reshape(-1)
    return c


def multi_block_cases():
    """(name, data, block_size) exercising several max-blocks, history overlap and bit-phase carry."""
    return [
        ("enwik2.5M_b1M", synth.enwik(2500000, seed=21), 0),
        ("moz700k_b32k", synth.mozilla(700000, seed=9), 32768),
        ("moz3.2M_b1M", synth.mozilla(3200000, seed=11), 0),
        ("exact2_b64k", synth.enwik(131072, seed=3), 65536),
        ("mix1.5M_b256k", synth.mix(1500000, seed=4, seg_lo=100000, seg_hi=400000), 262144),
        ("rand+text_b2M", np.concatenate([np.random.default_rng(3).integers(0, 256, size=300000).astype(np.uint8), synth.enwik(400000, seed=8)]), 2097152),
    ]
