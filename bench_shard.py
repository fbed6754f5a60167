"""Per-rank work of bench.py: one contiguous range of max-blocks per GPU.

N = 1: the whole stream in one zultra_cuda_compress_blocks_device call.
N > 1: every rank runs the phase-independent pipeline on its block range (zultra_cuda_shard_prepare) and learns its
shard size for each of the 8 possible entering bit phases; one NCCL all_gather of those 8-entry maps (+ the shard
checksum) lets every rank derive its true entering phase and absolute bit offset; it then emits its bitstream
(zultra_cuda_shard_emit) and the shard bitstreams are gathered on rank 0 over NCCL, where boundary bytes are OR-merged.
There is no other data-path collective.
"""
import ctypes as C
import time

import numpy as np


class ShardRunner:
    def __init__(self, z, ctx, shard, hist, lo, hi, n_total, flags, block, rank, world, dist, torch):
        """shard = bytes [lo - hist, hi) of the stream (hist <= 32 KiB of preceding input)."""
        self.z, self.ctx, self.flags, self.block, self.rank, self.world, self.dist, self.torch = z, ctx, flags, block, rank, world, dist, torch
        self.lo, self.hi, self.n_total = lo, hi, n_total
        self.hist = hist
        shard = np.ascontiguousarray(shard)
        assert len(shard) == hi - lo + hist
        self.host = torch.from_numpy(shard).pin_memory()
        self.dev_in = self.host.cuda()
        self.cap = max(1, (hi - lo) + (hi - lo) // 8 + 65536)
        self.dev_out = torch.zeros(self.cap, dtype=torch.uint8, device="cuda")
        self.last_out_bytes = 0
        self.host_out = torch.empty(self.cap + 64, dtype=torch.uint8).pin_memory()
        self.final = None
        self.checksum = 0
        if world > 1:
            # every shard's output is bounded by the largest shard's capacity
            caps = torch.tensor([self.cap], dtype=torch.int64, device="cuda")
            allc = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(allc, caps)
            self.maxcap = int(max(int(c.item()) for c in allc))
            if self.maxcap > self.cap:
                self.dev_out = torch.zeros(self.maxcap, dtype=torch.uint8, device="cuda")
            self.gather_list = [torch.zeros(self.maxcap, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
            self.final_buf = torch.zeros(n_total + n_total // 8 + 65536 * world, dtype=torch.uint8, device="cuda") if rank == 0 else None

    def step_device(self):
        """Input resident in HBM.  Returns milliseconds between two CUDA events around the step."""
        torch, dist = self.torch, self.dist
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = self.hi - self.lo
        if self.world == 1:
            bits, ck = self.ctx.compress_blocks_device(self.dev_in.data_ptr(), n, self.dev_out.data_ptr(), self.dev_out.numel(),
                                                       block=self.block, finalize=1, flags=self.flags)
            self.last_out_bytes = (bits + 7) // 8
            self.checksum = ck
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        if n > 0:
            maps, ck = self.ctx.shard_prepare(self.dev_in.data_ptr(), self.hist, n, block=self.block, finalize=1 if self.hi >= self.n_total else 0, flags=self.flags)
        else:
            maps, ck = list(range(8)), (1 if self.flags == 1 else 0)
        mine = torch.tensor(maps + [ck, n], dtype=torch.int64, device="cuda")
        allm = [torch.zeros(10, dtype=torch.int64, device="cuda") for _ in range(self.world)]
        dist.all_gather(allm, mine)
        allm = [m.tolist() for m in allm]
        from zultra_b200 import shard
        offs, nbits, abs_bits = shard.compose([m[:8] for m in allm])
        in_bits = offs[self.rank] & 7
        if n > 0:
            self.dev_out[: 8].zero_()
            bits = self.ctx.shard_emit(in_bits, self.dev_out.data_ptr(), self.dev_out.numel())
            assert bits - in_bits == nbits[self.rank]
        dist.gather(self.dev_out, self.gather_list, dst=0)
        if self.rank == 0:
            fb = self.final_buf
            total_bytes = (abs_bits + 7) // 8
            fb[: total_bytes + 8].zero_()
            for r in range(self.world):
                if nbits[r] == 0:
                    continue
                start = offs[r] >> 3
                nb = ((offs[r] & 7) + nbits[r] + 7) // 8
                fb[start:start + nb] |= self.gather_list[r][:nb]   # boundary byte OR-merged, the rest lands on zeros
            self.last_out_bytes = total_bytes
            ck = allm[0][8]
            for r in range(1, self.world):
                ck = self.z.load().zultra_cuda_checksum_combine(self.flags, ck, allm[r][8], allm[r][9])
            self.checksum = ck
            self.final = fb[:total_bytes]
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def final_stream(self):
        """Rank 0: the complete framed stream of the last step (header + deflate data + trailer) as bytes."""
        import zultra_b200 as z
        L = z.load()
        body = (self.final if self.world > 1 else self.dev_out[: self.last_out_bytes]).cpu().numpy().tobytes()
        hdr = (C.c_ubyte * 16)(); ftr = (C.c_ubyte * 16)()
        nh = L.zultra_frame_encode_header(hdr, 16, self.flags, None, 0)
        nf = L.zultra_frame_encode_footer(ftr, 16, C.c_uint(self.checksum), C.c_longlong(self.n_total), self.flags)
        return bytes(hdr[:nh]) + body + bytes(ftr[:nf])

    def e2e(self, steps):
        """Public API, pinned host input -> host output; returns (ms per step, h2d bytes, d2h bytes)."""
        L = self.z.load()
        n = self.hi - self.lo
        if n <= 0:
            return 0.0, 0, 0
        src = self.host.data_ptr() + self.hist
        times, out_bytes = [], 0
        warm = 2   # untimed: the pooled context of the public API allocates its device buffers on first use
        for it in range(warm + max(1, steps)):
            self.torch.cuda.synchronize()
            if self.dist is not None:
                self.dist.barrier()
            t0 = time.perf_counter()
            if self.world == 1:
                r = L.zultra_memory_compress(C.c_void_p(src), n, C.c_void_p(self.host_out.data_ptr()), self.host_out.numel(), self.flags, self.block)
                assert r != C.c_size_t(-1).value
                out_bytes = r
            else:
                # each rank: its block range from pinned host memory through the C-ABI, bitstream back to host memory
                bits = C.c_ulonglong(0); ck = C.c_uint(1 if self.flags == 1 else 0)
                rc = L.zultra_cuda_compress_blocks(self.ctx.p, C.c_void_p(self.host.data_ptr()), self.hist, C.c_void_p(src), n, self.block,
                                                   1 if self.hi >= self.n_total else 0, 0, self.flags, C.byref(ck), C.c_void_p(self.host_out.data_ptr()),
                                                   self.host_out.numel(), C.byref(bits))
                assert rc == 0
                out_bytes = (bits.value + 7) // 8
            if it >= warm:
                times.append(time.perf_counter() - t0)
        return 1000.0 * sum(times) / len(times), n, out_bytes
