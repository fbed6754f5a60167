"""Per-rank work of bench.py: one contiguous range of max-blocks per GPU."""
import numpy as np


class ShardRunner:
    def __init__(self, z, ctx, data, lo, hi, flags, block, rank, world, dist, torch):
        self.z, self.ctx, self.flags, self.block, self.rank, self.world, self.dist, self.torch = z, ctx, flags, block, rank, world, dist, torch
        self.lo, self.hi = lo, hi
        self.data = data
        hist = min(lo, 32768)
        self.hist = hist
        shard = np.ascontiguousarray(data[lo - hist:hi])
        self.host = torch.from_numpy(shard).pin_memory()
        self.dev_in = self.host.cuda()
        self.dev_out = torch.empty(max(1, (hi - lo) + (hi - lo) // 8 + 65536), dtype=torch.uint8, device="cuda")
        self.last_out_bytes = 0
        self.host_out = torch.empty(self.dev_out.numel() + 64, dtype=torch.uint8).pin_memory()

    def step_device(self):
        """Input resident in HBM; returns device milliseconds (CUDA events on the library stream)."""
        if self.hi <= self.lo:
            return 0.0
        if self.world == 1:
            bits, ck = self.ctx.compress_blocks_device(self.dev_in.data_ptr(), self.hi - self.lo, self.dev_out.data_ptr(), self.dev_out.numel(),
                                                       block=self.block, finalize=1, flags=self.flags)
            self.last_out_bytes = (bits + 7) // 8
            return self.ctx.timings()["total"]
        raise NotImplementedError("multi-GPU sharded step is wired in bench_multi")

    def e2e(self, steps):
        """Public API, pinned host input -> host output; returns (ms per step, h2d bytes, d2h bytes)."""
        import time
        import ctypes as C
        L = self.z.load()
        n = self.hi - self.lo
        src = self.host.data_ptr() + self.hist
        best = []
        for _ in range(max(1, steps)):
            self.torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = L.zultra_memory_compress(C.c_void_p(src), n, C.c_void_p(self.host_out.data_ptr()), self.host_out.numel(), self.flags, self.block)
            dt = time.perf_counter() - t0
            assert r != C.c_size_t(-1).value
            best.append(dt)
            out_bytes = r
        return 1000.0 * sum(best) / len(best), n, out_bytes
