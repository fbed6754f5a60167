#!/usr/bin/env python
"""bench.py - headline metric of BASELINE.json: input MB/s of the zultra compression hot path on N B200s.

One "step" = one pass of the whole pipeline (suffix array + LCP, match lists, block split, optimal parse, Huffman,
bit emission) over the workload.  N=1 workload: configs[1] of BASELINE.json, a synthetic 100 000 000 B
enwik8-shaped XML/wiki text, deflate format, default 1 MiB max-block (96 blocks).
  value      MB/s (10^6 B/s) with the input already resident in HBM (zultra_cuda_compress_blocks_device), timed
             with CUDA events on the library's stream, max over ranks
  e2e        same metric through the public zultra_memory_compress call with pinned HOST buffers, H2D/D2H inside
  roofline   dominant kernel: algorithmic bytes / its average CUDA-event duration vs the measured HBM copy peak
  cpu_baseline   the reference's own CPU code (oracle/_ref) on a bounded sample, one thread
--impl reference times the reference CPU implementation on the host cores instead (no GPU work).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "input MB/s (zultra compression hot path, byte-identical to CPU zultra)"   # both arms report this metric
WORKLOADS = {
    "enwik100m": dict(size=100_000_000, flags=0, fmt="deflate", gen="enwik"),
    "mozilla51m": dict(size=51_220_480, flags=2, fmt="gzip", gen="mozilla"),
    "mix1g": dict(size=1 << 30, flags=2, fmt="gzip", gen="mix"),
}
# algorithmic bytes per unit for the kernels that can dominate
KERNEL_BYTES = {
    "mf_build_walk": ("per block byte: 4 B SA|LCP word in (x1.031 window) + 32 B match list out", lambda n, P: 4.0 * P + 32.0 * n),
    "mf_tile_filter": ("per window position: 4 B word read + 4 B tile word written", lambda n, P: 8.0 * P),
    "parse_dp": ("per block byte and pass: 32 B match list + 1 B text read, 4 B choice written; 4 passes", lambda n, P: 4 * 37.0 * n),
    "rs_scatter": ("per sorted element and pass: 12 B key+index read, 12 B written", lambda n, P: 24.0 * P),
    "rs_hist": ("per sorted element and pass: 8 B key read", lambda n, P: 8.0 * P),
    "lcp_pack": ("per window position: 4 B SA read, 1 B text, 4 B word written", lambda n, P: 9.0 * P),
}


def gen_workload(name, size=None, seg=0):
    """Segment `seg` of a workload (segment 0 is the configuration itself; further segments, other seeds, make up the
    N x larger stream of a weak-scaling run).  Cached under /tmp; written atomically since ranks may race."""
    from zultra_b200 import synth
    w = WORKLOADS[name]
    n = size or w["size"]
    cache = os.path.join("/tmp", "zb_%s_%d_s%d.npy" % (name, n, seg))
    if os.path.exists(cache):
        return np.load(cache)
    fn = getattr(synth, w["gen"])
    d = fn(n) if seg == 0 else fn(n, seed=0x5A170100 + 1009 * seg)
    try:
        tmp = cache + ".%d.tmp.npy" % os.getpid()
        np.save(tmp, d)
        os.replace(tmp, cache)
    except OSError:
        pass
    return d


def stream_range(name, size, lo, hi):
    """Bytes [lo, hi) of the stream made of consecutive segments of `size` bytes each."""
    parts = []
    for sg in range(lo // size, (max(lo, hi - 1)) // size + 1):
        d = gen_workload(name, size, sg)
        a, b = max(lo, sg * size) - sg * size, min(hi, (sg + 1) * size) - sg * size
        if b > a:
            parts.append(d[a:b])
    return np.concatenate(parts) if len(parts) != 1 else np.ascontiguousarray(parts[0])


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture of this workload (profiles/), or None."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return j.get(kernel)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is started
    before the warm-up (nvidia-smi needs a few hundred ms to deliver its first row); every row is stamped on arrival and only
    rows that fall inside [mark_begin, mark_end] are reported (all rows under load if the region was shorter than a period)."""

    def __init__(self, index=0):
        self.rows, self.p, self.index, self.t0, self.t1 = [], None, index, None, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index),
                                       "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)   # let the row sampled at the end of the region arrive
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        rows = inside if inside else [r for (t, r) in self.rows if self.t0 is None or t >= self.t0 - 1.0]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(rows)}


def cpu_reference_timing(data, flags, threads, slice_bytes):
    """Reference CPU code (oracle/_ref): `threads` workers, each compressing its own slice of the workload with the stock
    zultra_memory_compress.  Returns (MB/s aggregate, seconds)."""
    import refharness
    if not os.path.exists(refharness.REF_SO):
        return None, None, "oracle/_ref missing"
    lib = C.CDLL(refharness.REF_SO)
    lib.zultra_memory_compress.restype = C.c_size_t
    lib.zultra_memory_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
    lib.zultra_memory_bound.restype = C.c_size_t
    lib.zultra_memory_bound.argtypes = [C.c_size_t, C.c_uint, C.c_uint]
    slices = [np.ascontiguousarray(data[i * slice_bytes:(i + 1) * slice_bytes]) for i in range(threads)]
    slices = [s for s in slices if len(s)]
    outs = [np.empty(lib.zultra_memory_bound(len(s), flags, 0), dtype=np.uint8) for s in slices]

    def work(i):
        lib.zultra_memory_compress(slices[i].ctypes.data, len(slices[i]), outs[i].ctypes.data, len(outs[i]), flags, 0)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(slices))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = sum(len(s) for s in slices)
    return total / dt / 1e6, dt, "%d thread(s) x %d B slices of the workload, stock zultra_memory_compress" % (len(slices), slice_bytes)


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    data = gen_workload(name, args.size)
    world = max(1, world)
    cores = max(1, min(os.cpu_count() or 1, 64))
    slice_bytes = 2 << 20
    vals = []
    for it in range(args.warmup + args.steps):
        v, dt, sample = cpu_reference_timing(data[(it * cores * slice_bytes) % max(1, len(data) - cores * slice_bytes):], w["flags"], cores, slice_bytes)
        if v is None:
            print(json.dumps({"impl": "reference", "unavailable": sample}))
            return
        if it >= args.warmup:
            vals.append((v, dt))
    v = sum(x[0] for x in vals) / len(vals)
    ms = 1000.0 * sum(x[1] for x in vals) / len(vals)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "MB/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": {"workload": name if world == 1 or name != "enwik100m" else "%s x %d (one stream of %d segments of the configuration's shape, sharded by block range)" % (name, world, world),
                                 "bytes": int(len(data)) * (world if name == "enwik100m" else 1), "format": w["fmt"], "max_block": 1048576,
                                 "blocks": int((int(len(data)) * (world if name == "enwik100m" else 1) + 1048575) // 1048576),
                                 "sample": "CPU threads each compress a 2 MiB slice of segment 0 per step"},
                      "cpu_baseline": {"value": round(v, 3), "unit": "MB/s", "cores": cores, "kind": "reference", "sample": sample},
                      "e2e": {"value": round(v, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="enwik100m")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"], help="N > 1: weak = one stream of N x the configuration (default for enwik100m), strong = the configuration split N ways")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="no per-kernel CUDA events in the timed steps")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import zultra_b200 as z
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload
    w = WORKLOADS[name]
    seg_size = args.size or w["size"]
    # weak scaling (default for the headline workload): the job is ONE stream of N segments of the configuration's size,
    # sharded by contiguous max-block ranges; strong: the configuration itself split N ways
    weak = world > 1 and (args.scaling == "weak" or (args.scaling == "auto" and name == "enwik100m"))
    n = seg_size * (world if weak else 1)
    block = 1 << 20
    nblocks = (n + block - 1) // block
    # shard by contiguous max-block ranges; every rank keeps the 32 KiB before its first block as history
    from zultra_b200 import shard
    lo, hi = shard.plan_shards(n, block, world)[rank]
    b0, b1 = lo // block, (hi + block - 1) // block
    hist = min(lo, 32768)
    shard_bytes = stream_range(name, seg_size, lo - hist, hi)
    L = z.load()
    L.zultra_cuda_profile.argtypes = [C.c_int]
    L.zultra_cuda_profile_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    ctx = z.CudaCtx(local)
    from bench_shard import ShardRunner
    runner = ShardRunner(z, ctx, shard_bytes, hist, lo, hi, n, w["flags"], block, rank, world, dist, torch)
    if dist is not None:
        dist.barrier()     # every segment is in the /tmp cache now
    data = stream_range(name, seg_size, 0, n) if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(profile):
        flush.fill_(1)   # L2 flush between iterations
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        L.zultra_cuda_profile(1 if profile else 0)
        ms = runner.step_device()
        L.zultra_cuda_profile(0)
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(False)
    # parity gate before any timing counts: the stream must inflate back to the input (and equal the 1-GPU stream)
    if rank == 0:
        import hashlib
        import zlib
        stream = runner.final_stream()
        raw = data.tobytes()
        assert zlib.decompress(stream, {0: -15, 1: 15, 2: 31}[w["flags"]]) == raw, "output does not inflate to the input"
        stream_sha = hashlib.sha256(stream).hexdigest()
        if world > 1 and len(data) <= (1 << 30):
            one = z.memory_compress(data, w["flags"], block)
            assert one == stream, "sharded stream differs from the single-GPU stream"
        del raw
    times, launches = [], 0
    sampler.mark_begin()
    for _ in range(args.steps):
        times.append(step(not args.no_profile))
        launches += ctx.counters()["launches"]
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    names = C.create_string_buffer(32 * 256); kms = (C.c_float * 256)(); kcnt = (C.c_int * 256)()
    nk = L.zultra_cuda_profile_collect(names, kms, kcnt, 256)
    ktab = sorted([(names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), kms[i], kcnt[i]) for i in range(nk)], key=lambda r: -r[1])
    if world > 1:
        ktab = [r for r in ktab]
    t = torch.tensor([sum(times)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step / 1e3) / 1e6
    stages = ctx.timings()
    # end to end through the public API with pinned host buffers
    e2e_ms, h2d, d2h = runner.e2e(args.steps)
    if dist is not None:
        t2 = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item())
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        top = next((r for r in ktab if r[0] in KERNEL_BYTES), ktab[0] if ktab else ("none", 0.0, 1))
        P_local = (hi - lo) + (b1 - b0) * 32768
        desc, fn = KERNEL_BYTES.get(top[0], ("unknown", lambda a, b: 0.0))
        alg_bytes_per_step = fn(hi - lo, P_local)
        per_launch_ms = top[1] / max(1, top[2])
        launches_per_step = top[2] / args.steps
        achieved = (alg_bytes_per_step / launches_per_step) / (per_launch_ms / 1e3) / 1e9 if per_launch_ms > 0 else 0.0
        out = {"metric": METRIC, "value": round(value, 2), "unit": "MB/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
               "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "config": {"workload": name if not weak else "%s x %d (one stream of %d segments of the configuration's shape, sharded by block range)" % (name, world, world),
                          "bytes": int(n), "format": w["fmt"], "max_block": block, "blocks": int(nblocks),
                          "parallelism": "block-range shards x%d, %d concurrent lanes (streams) per GPU" % (world, max(1, ctx.counters()["r5"])), "l2": "256 MiB flush write between iterations"},
               "clocks": clocks,
               "e2e": {"value": round(n / (e2e_ms / 1e3) / 1e6, 2), "unit": "MB/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": top[0], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                            "traffic": ncu_traffic(top[0]), "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, enwik100m at N=1)",
                            "algorithmic_bytes": desc, "kernel_ms_per_step": round(top[1] / args.steps, 3),
                            "kernel_share_of_step": round(top[1] / max(1e-9, sum(r[1] for r in ktab)), 4), "kernel_share_basis": "sum of all kernel durations (lanes overlap, so wall time is shorter)",
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6.65 TB/s"},
               "stages_ms": {k: round(v, 3) for k, v in stages.items()},
               "kernels_ms_per_step": {r[0]: round(r[1] / args.steps, 3) for r in ktab[:48]},
               "counters": ctx.counters(), "compressed_bytes": runner.last_out_bytes, "stream_sha256": stream_sha, "verified": "inflate(stream) == input" + (" and == 1-GPU stream" if world > 1 else "")}
        if not args.no_cpu_baseline and world == 1:
            v, dt, sample = cpu_reference_timing(data, w["flags"], 1, 24 << 20)
            if v is not None:
                out["cpu_baseline"] = {"value": round(v, 3), "unit": "MB/s", "cores": 1, "kind": "reference",
                                       "sample": "first 24 MiB of the workload, " + sample + ", %.1f s" % dt}
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
