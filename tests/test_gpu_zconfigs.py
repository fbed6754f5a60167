"""GPU parity AT THE STATED CONFIG SIZES of BASELINE.json (VERDICT r1 "what's weak" 1-2): the CUDA path against
  (1) tests/golden/config_golden.npz - sha256 of the unmodified reference's complete streams for js48k / enwik100m /
      mozilla51m / mix1g and of every one of the 100 000 batch payloads (made by tests/golden/make_config_golden.py here,
      where the reference compiles; the GPU box needs no reference for these), and
  (2) the compiled reference oracle/_ref run on the box's host threads, for the paths whose code changes with size:
      the 256-block batch boundary of the one-shot call, forced parse chunk lengths, multi-wave tile lists, the batch API.
Bit-exact everywhere (integer / byte work)."""
import hashlib
import os
import threading
import zlib

import numpy as np
import pytest

import refharness
from zultra_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD_PATH = os.path.join(ROOT, "tests", "golden", "config_golden.npz")
CFG = {"js48k": ("js48k", 48944, 1), "enwik100m": ("enwik", 100_000_000, 0), "mozilla51m": ("mozilla", 51_220_480, 2), "mix1g": ("mix", 1 << 30, 2)}
CHUNK = 4 << 20


@pytest.fixture(scope="module")
def z():
    import zultra_b200
    return zultra_b200


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD_PATH):
        pytest.skip("tests/golden/config_golden.npz missing")
    return np.load(GOLD_PATH)


def _workload(name):
    """Same /tmp cache as bench.py (the 1 GiB stream takes over a minute to generate)."""
    import bench
    gen, size, flags = CFG[name]
    if name == "js48k":
        return synth.js48k(), flags
    return bench.gen_workload(name), flags


def _check_stream(gold, name, data, out):
    if name + "/out_sha" not in gold:
        pytest.skip("no golden entry for " + name)
    assert hashlib.sha256(data.tobytes()).digest() == gold[name + "/in_sha"].tobytes(), "synthetic input differs from the one the golden vector was made from"
    assert out is not None
    if hashlib.sha256(out).digest() != gold[name + "/out_sha"].tobytes():
        ch = gold[name + "/chunks"]
        bad = [i for i in range(len(ch)) if hashlib.sha256(out[i * CHUNK:(i + 1) * CHUNK]).digest()[:8] != ch[i].tobytes()]
        raise AssertionError("%s: stream differs from the reference: len %d vs %d, first differing 4 MiB piece %s" % (name, len(out), int(gold[name + "/out_len"][0]), bad[:1]))


@pytest.mark.parametrize("name", ["js48k", "enwik100m", "mozilla51m"])
def test_config_full_stream_vs_reference_golden(z, gold, name):
    """C1-C3: the complete stream of the configuration == the reference's (sha256 of the whole output)."""
    data, flags = _workload(name)
    out = z.memory_compress(data, flags)
    _check_stream(gold, name, data, out)


def test_config_mix1g_full_stream_vs_reference_golden(z, gold):
    """C4: 1 GiB through zultra_memory_compress (4 engine calls of 256 blocks, bit phase carried) == the reference."""
    data, flags = _workload("mix1g")
    out = z.memory_compress(data, flags)
    _check_stream(gold, "mix1g", data, out)
    assert out[-4:] == (len(data) & 0xffffffff).to_bytes(4, "little")


def test_config_batch100k_vs_reference_golden(z, gold):
    """C5: all 100 000 payloads through zultra_cuda_memory_compress_batch, every stream == the reference's."""
    if "batch100k/out_sha8" not in gold:
        pytest.skip("no golden entry for batch100k")
    pay = synth.batch(100000)
    want, wlen, win = gold["batch100k/out_sha8"], gold["batch100k/out_len"], gold["batch100k/in_sha8"]
    ctx = z.CudaCtx()
    try:
        step = 20000
        for lo in range(0, len(pay), step):
            part = pay[lo:lo + step]
            outs = ctx.memory_compress_batch(part, 1)
            for i, (p, o) in enumerate(zip(part, outs)):
                k = lo + i
                if k % 997 == 0:
                    assert hashlib.sha256(p.tobytes()).digest()[:8] == win[k].tobytes(), "payload %d differs from the golden generator's" % k
                assert o is not None and len(o) == int(wlen[k]) and hashlib.sha256(o).digest()[:8] == want[k].tobytes(), "payload %d differs from the reference" % k
    finally:
        ctx.close()


def _ref_many(ref, items, flags, block=0, threads=None):
    """Reference streams of many independent inputs on the host's threads (ctypes releases the GIL)."""
    outs = [None] * len(items)
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(items):
                return
            outs[i] = ref.compress(items[i], flags=flags, block=block)

    ths = [threading.Thread(target=work) for _ in range(threads or min(32, os.cpu_count() or 1))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return outs


def test_batch_api_vs_compiled_reference(z, ref):
    """600 batch-shaped payloads (PNG-IDAT / HTTP body) through the batch API == oracle/_ref, stream by stream."""
    pay = synth.batch(600, seed=4242)
    want = _ref_many(ref, pay, 1)
    got = z.memory_compress_batch(pay, 1)
    for k, (a, b) in enumerate(zip(got, want)):
        assert a == b, k


def test_one_shot_across_256_block_batches_vs_compiled_reference(z, ref):
    """10 MiB at max-block 32768 = 320 blocks: zultra_memory_compress cuts it into two engine calls and carries the pending
    bits of the last byte across (csrc/host/libzultra.c); the joined stream must equal the reference's."""
    rng = np.random.default_rng(11)
    data = np.concatenate([synth.enwik(3 << 20, seed=51), rng.integers(0, 256, size=1 << 20).astype(np.uint8), synth.mozilla(6 << 20, seed=52)])
    assert len(data) // 32768 > 256
    for flags in (0, 1, 2):
        assert z.memory_compress(data, flags, 32768) == ref.compress(data, flags=flags, block=32768), flags


@pytest.mark.parametrize("cd,awu", [(704, 1), (960, 1), (1856, 1), (2048, 1), (64, 1), (128, 1), (200, 1), (320, 1), (512, 0), (128, 0)])
def test_forced_parse_chunk_vs_compiled_reference(z, ref, monkeypatch, cd, awu):
    """The parse chunk length is chosen from the batch size (960 at 100 MB, up to 1856 for 256-block batches, 128 for small
    inputs): force the lengths the configurations use on an input the reference finishes in seconds - with the adaptive
    warm-up (a chunk's warm-up and the verified part of its signature follow what its candidates reach; chunks shorter than
    the 258-position horizon take the propagated form need(c) = max(reach(c), need(c - 1) - CD)) and with it switched off."""
    monkeypatch.setenv("ZULTRA_CUDA_PARSE_CD", str(cd))
    monkeypatch.setenv("ZULTRA_CUDA_PARSE_AWU", str(awu))
    data = synth.mix(6 << 20, seed=300 + cd, seg_lo=400000, seg_hi=2 << 20)
    c = z.CudaCtx()
    try:
        got, bits, ck = c.compress_blocks(data, finalize=1, flags=2)
    finally:
        c.close()
    want = ref.compress(data, flags=2)
    assert got == want[10:-8] and ck == zlib.crc32(data.tobytes())


def test_multi_wave_tile_lists_vs_compiled_reference(z, ref, monkeypatch):
    """More than 16384 match-finder tiles in one call (only > 128 MiB inputs reach that with the default tile): forced with
    512-position tiles on 10 MiB."""
    monkeypatch.setenv("ZULTRA_CUDA_TILE", "512")
    data = synth.mix(10 << 20, seed=77, seg_lo=1 << 20, seg_hi=3 << 20)
    c = z.CudaCtx()
    try:
        got, bits, ck = c.compress_blocks(data, finalize=1, flags=0)
        assert c.counters()["r7"] > 16384
    finally:
        c.close()
    assert got == ref.compress(data, flags=0)


def test_stream_api_large_single_call_vs_golden(z, gold):
    """one zultra_stream_compress(FINALIZE) call holding the whole 100 MB must be cut into bounded engine calls
    (not one unbounded launch) and still produce the reference's stream."""
    data, flags = _workload("enwik100m")
    s = z.Stream(flags)
    st, out = s.compress(data, z.ZULTRA_FINALIZE, out_chunk=8 << 20)
    s.end()
    assert st == z.ZULTRA_STREAM_END
    _check_stream(gold, "enwik100m", data, out)
