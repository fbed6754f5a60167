"""world_size-2 gloo test of the multi-GPU host logic (block-range shards, 8-phase maps, bit-offset stitching) with the
host-emulated pipeline standing in for the GPUs; the stitched stream must equal the reference's one-shot stream."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, block, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import refharness
    from zultra_b200 import shard, synth
    data = np.concatenate([synth.mozilla(150000, seed=51), np.random.default_rng(5).integers(0, 256, size=70000).astype(np.uint8), synth.enwik(130000, seed=52)])
    emu = refharness.Emu()
    lo, hi = shard.plan_shards(len(data), block, world)[rank]
    hist = min(lo, 32768)
    maps = emu.shard_prepare(data[lo - hist:hi], hist, hi - lo, block=block, finalize=1 if hi >= len(data) else 0) if hi > lo else list(range(8))
    mine = torch.tensor(maps, dtype=torch.int64)
    allm = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allm, mine)
    offs, nbits, total = shard.compose([m.tolist() for m in allm])
    buf, bits = emu.shard_emit(offs[rank] & 7, (hi - lo) + 70000) if hi > lo else (b"", offs[rank] & 7)
    assert bits - (offs[rank] & 7) == nbits[rank]
    pad = torch.zeros(len(data) + 70000, dtype=torch.uint8)
    pad[: len(buf)] = torch.from_numpy(np.frombuffer(buf, dtype=np.uint8).copy())
    gl = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, gl, dst=0)
    if rank == 0:
        stream = shard.merge([g.numpy().tobytes() for g in gl], offs, nbits, total)
        ref = refharness.Ref().compress(data, flags=0, block=block) if os.path.exists(refharness.REF_SO) else None
        import oracle_py
        q.put((stream == oracle_py.compress(data, 0, block), ref is None or stream == ref))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("block", [32768, 65536])
def test_two_rank_sharding_stitches_to_the_reference_stream(block):
    import refharness
    if not os.path.exists(refharness.EMU_SO):
        import subprocess
        subprocess.check_call(["make", "-C", ROOT, "emu"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + (1 if block == 65536 else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, block, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    ok_oracle, ok_ref = q.get(timeout=5)
    assert ok_oracle and ok_ref


def _chunk_worker(rank, world, port, block, flags, q):
    """Round-robin chunks (zb_capi.cu run_multi / bench.py under torchrun): every rank prepares ALL its chunks in one
    pipeline pass, the per-chunk 8-phase maps are all-gathered, composed in stream order, every chunk is emitted at its true
    phase and rank 0 ORs the chunk bitstreams together at their absolute bit offsets."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import zlib
    import refharness
    from zultra_b200 import shard, synth
    data = np.concatenate([synth.mozilla(150000, seed=61), np.random.default_rng(6).integers(0, 256, size=70000).astype(np.uint8), synth.enwik(130000, seed=62),
                           np.zeros(40000, dtype=np.uint8), synth.mozilla(60000, seed=63)])
    emu = refharness.Emu()
    plan = shard.plan_chunks(len(data), block, world, g=2)
    mine = [(j, lo, hi) for j, (lo, hi, r) in enumerate(plan) if r == rank]
    chunks = [(lo - min(lo, 32768), min(lo, 32768), hi - lo, 1 if hi >= len(data) else 0) for (_, lo, hi) in mine]
    maps, cks = emu.chunks_prepare(data, chunks, block=block, flags=flags)
    per = max(sum(1 for c in plan if c[2] == r) for r in range(world))
    t = torch.zeros((per, 10), dtype=torch.int64)
    for i, (j, lo, hi) in enumerate(mine):
        t[i, :8] = torch.tensor(maps[i]); t[i, 8] = cks[i]; t[i, 9] = hi - lo
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    idx = [0] * world
    allmaps, allck = [], []
    for (lo, hi, r) in plan:
        row = allt[r][idx[r]]; idx[r] += 1
        allmaps.append(row[:8].tolist()); allck.append((int(row[8]), int(row[9])))
    offs, nbits, total = shard.compose(allmaps)
    bufs, bits = emu.chunks_emit([offs[j] & 7 for (j, _, _) in mine], len(data) + 70000 * len(mine))
    for (j, _, _), b in zip(mine, bits):
        assert b - (offs[j] & 7) == nbits[j]
    # gather: fixed-size rows, one per chunk slot
    cap = max(hi - lo for lo, hi, _ in plan) + 70000
    pad = torch.zeros((per, cap), dtype=torch.uint8)
    for i, b in enumerate(bufs):
        pad[i, : len(b)] = torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy())
    gl = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, gl, dst=0)
    if rank == 0:
        idx = [0] * world
        parts = []
        for (lo, hi, r) in plan:
            parts.append(gl[r][idx[r]].numpy().tobytes()); idx[r] += 1
        stream = shard.merge(parts, offs, nbits, total)
        raw = data.tobytes()
        want_ck = zlib.crc32(raw) if flags == 2 else (zlib.adler32(raw) if flags == 1 else 0)
        ref = refharness.Ref().compress(data, flags=flags, block=block) if os.path.exists(refharness.REF_SO) else None
        import oracle_py
        body = oracle_py.compress(data, flags, block)
        hdr = 0 if flags == 0 else (2 if flags == 1 else 10)
        ftr = 0 if flags == 0 else (4 if flags == 1 else 8)
        # per-chunk checksums are those of the chunk bytes from the initial value
        pos = 0
        ok_ck = True
        for (lo, hi, r), (c, ln) in zip(plan, allck):
            piece = raw[lo:hi]
            ok_ck = ok_ck and ln == hi - lo and c == (zlib.crc32(piece) if flags == 2 else (zlib.adler32(piece) if flags == 1 else c))
        q.put((stream == body[hdr:len(body) - ftr], ref is None or stream == ref[hdr:len(ref) - ftr], ok_ck, want_ck is not None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("block,flags", [(32768, 0), (65536, 2)])
def test_two_rank_round_robin_chunks_stitch_to_the_reference_stream(block, flags):
    import refharness
    if not os.path.exists(refharness.EMU_SO):
        import subprocess
        subprocess.check_call(["make", "-C", ROOT, "emu"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 300) + (1 if block == 65536 else 0)
    procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, block, flags, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert all(q.get(timeout=5))


def test_plan_chunks():
    from zultra_b200 import shard
    plan = shard.plan_chunks(10 * 65536 + 5, 65536, 3, g=2)
    assert [c[2] for c in plan] == [0, 1, 2, 0, 1, 2] and plan[0][:2] == (0, 131072) and plan[-1][:2] == (655360, 655365)
    assert shard.chunk_blocks(1024, 8) == 8 and shard.chunk_blocks(49, 8) == 1 and shard.chunk_blocks(96, 2) == 8


def test_compose_and_plan():
    from zultra_b200 import shard
    assert shard.plan_shards(10 * 1048576 + 5, 1048576, 4) == [(0, 3145728), (3145728, 6291456), (6291456, 9437184), (9437184, 10485765)]
    assert shard.plan_shards(100, 1048576, 2) == [(0, 100), (100, 100)]
    maps = [[10 + p for p in range(8)], [21 + p for p in range(8)]]
    offs, nb, tot = shard.compose(maps)
    assert offs == [0, 10] and nb == [10, 21] and tot == 31


def test_bench_stream_range_is_the_concatenation_of_segments(tmp_path, monkeypatch):
    """bench.py's weak-scaling stream: segment r of a workload comes from its own seed, a rank's byte range [lo, hi) may
    span two segments, and the pieces of all ranks (minus their 32 KiB history overlap) tile the stream exactly."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from zultra_b200 import shard
    size, world, block = 300_000, 3, 65536
    whole = bench.stream_range("enwik100m", size, 0, size * world)
    assert len(whole) == size * world
    segs = [bench.gen_workload("enwik100m", size, s) for s in range(world)]
    assert bytes(whole) == b"".join(s.tobytes() for s in segs)
    assert bytes(segs[0]) != bytes(segs[1])
    got = bytearray()
    for lo, hi in shard.plan_shards(size * world, block, world):
        hist = min(lo, 32768)
        piece = bench.stream_range("enwik100m", size, lo - hist, hi)
        assert len(piece) == hi - lo + hist
        got += piece[hist:].tobytes()
    assert bytes(got) == bytes(whole)
