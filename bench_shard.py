"""Per-rank work of bench.py: one process per GPU, the stream cut into chunks of a few max-blocks dealt round-robin to the
ranks (zultra_b200/shard.py plan_chunks - the same scheme the library uses in-process, zb_capi.cu run_multi).

N = 1: the whole stream in one zultra_cuda_compress_blocks_device call.
N > 1, every step:
  compute     zultra_cuda_chunks_prepare: ONE pipeline pass over all chunks of this rank (input resident in HBM), giving per
              chunk its size in bits for each of the 8 possible entering bit phases + the checksum of its bytes
  all_gather  NCCL all_gather of those maps (10 int64 per chunk): every rank composes them, in stream order, into every
              chunk's true phase and absolute bit offset - the bit-offset scan
  emit        zultra_cuda_chunks_emit: the bitstreams for the true phases
  gather      NCCL send/recv of exactly the produced bytes of every rank to rank 0
  merge       rank 0: zultra_cuda_stitch_device, one kernel that puts every chunk at its bit offset (boundary bytes OR-merged)
There is no other data-path collective.  Times are per rank (ms, host clock around device-synchronous phases).
A pipeline pass is bounded to 256 max-blocks per GPU (~28 GB of working buffers): longer shares (1 GiB on 1-2 GPUs) take
several passes per rank; the chunks of pass p precede those of pass p + 1 in the stream, so the scan simply carries on.

"""
import ctypes as C
import time

import numpy as np

from zultra_b200 import shard


class _DevMem:
    """View of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def span_layout(phases, nbits):
    """Byte offsets of a rank's chunk bitstreams inside its pipeline's output buffer (stage_emit_finish: every stream gets
    a word-aligned span followed by two spare words) and the total span in bytes."""
    offs, w = [], 0
    for ph, nb in zip(phases, nbits):
        offs.append(4 * w)
        w += (ph + nb + 31) // 32 + 2
    return offs, 4 * w


class ShardRunner:
    PASS_BLOCKS = 256

    def __init__(self, z, ctx, get_range, n_total, flags, block, rank, world, dist, torch):
        """get_range(lo, hi) -> uint8 array of stream bytes [lo, hi)."""
        self.z, self.ctx, self.flags, self.block, self.rank, self.world, self.dist, self.torch = z, ctx, flags, block, rank, world, dist, torch
        self.n_total = n_total
        nblocks = (n_total + block - 1) // block
        self.single = world == 1 and nblocks <= self.PASS_BLOCKS      # the whole stream in one zultra_cuda_compress_blocks_device call
        self.plan = shard.plan_chunks(n_total, block, world)
        g = shard.chunk_blocks(nblocks, world)
        per_pass = max(1, self.PASS_BLOCKS // g) * world               # chunks of the stream per pass (all ranks together)
        self.passes = [list(range(a, min(len(self.plan), a + per_pass))) for a in range(0, len(self.plan), per_pass)]
        self.mine = [j for j, c in enumerate(self.plan) if c[2] == rank]
        parts, self.chunk_of, at = [], {}, 0
        if self.single:
            parts.append(get_range(0, n_total))
        else:
            for j in self.mine:
                lo, hi, _ = self.plan[j]
                h = min(lo, 32768)
                parts.append(get_range(lo - h, hi))
                self.chunk_of[j] = (at, h, hi - lo, 1 if hi >= n_total else 0)
                at += h + hi - lo
        self.my_bytes = sum(c[1] - c[0] for c in self.plan if c[2] == rank)
        buf = np.concatenate(parts) if len(parts) != 1 else np.ascontiguousarray(parts[0])
        self.dev_in = torch.from_numpy(buf).cuda()
        del buf, parts
        self.cap = max(1, self.my_bytes + self.my_bytes // 8 + 65536 * max(1, len(self.mine)))
        self.last_out_bytes, self.final, self.checksum = 0, None, 0
        self.breakdown = {}
        self.per = max(sum(1 for j in ps if self.plan[j][2] == r) for ps in self.passes for r in range(world))
        if self.single:
            self.dev_out = torch.zeros(self.cap, dtype=torch.uint8, device="cuda")
            return
        self.acc = torch.empty(self.cap, dtype=torch.uint8, device="cuda")      # this rank's chunk bitstreams, pass after pass
        if world > 1:
            caps = torch.tensor([self.cap], dtype=torch.int64, device="cuda")
            allc = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(allc, caps)
            self.stage = [torch.empty(int(c.item()), dtype=torch.uint8, device="cuda") if r != 0 else None for r, c in enumerate(allc)] if rank == 0 else None
        if rank == 0:
            self.final_buf = torch.zeros(n_total + n_total // 8 + 65536 * world, dtype=torch.uint8, device="cuda")

    def step_device(self):
        """Input resident in HBM.  Returns milliseconds (CUDA events around the whole step on torch's stream; every library
        call inside is synchronous on its own stream, so the events bracket all of it)."""
        torch, dist = self.torch, self.dist
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if self.single:
            t0 = time.perf_counter()
            bits, ck = self.ctx.compress_blocks_device(self.dev_in.data_ptr(), self.n_total, self.dev_out.data_ptr(), self.dev_out.numel(),
                                                       block=self.block, finalize=1, flags=self.flags)
            self.last_out_bytes = (bits + 7) // 8
            self.checksum = ck
            e1.record(); torch.cuda.synchronize()
            self.breakdown = {"compute": 1e3 * (time.perf_counter() - t0)}
            return e0.elapsed_time(e1)
        tm = {"compute": 0.0, "all_gather": 0.0, "emit": 0.0, "gather": 0.0, "merge": 0.0}
        pos = 0                                    # running bit position of the scan
        spans = [[] for _ in range(self.world)]    # per rank: (chunk index, byte offset in that rank's acc buffer)
        acc_at = [0] * self.world
        offs, nbits, cks_all = {}, {}, {}
        for ps in self.passes:
            t0 = time.perf_counter()
            js = [j for j in ps if self.plan[j][2] == self.rank]
            maps, cks = self.ctx.chunks_prepare(self.dev_in.data_ptr(), [self.chunk_of[j] for j in js], block=self.block, flags=self.flags) if js else ([], [])
            t1 = time.perf_counter()
            rows = {}
            if self.world > 1:
                mine = torch.zeros((self.per, 9), dtype=torch.int64)
                for i, (m, c) in enumerate(zip(maps, cks)):
                    mine[i, :8] = torch.tensor(m); mine[i, 8] = c
                mine = mine.cuda()
                allm = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(allm, mine)
                allm = [m.cpu() for m in allm]
                idx = [0] * self.world
                for j in ps:
                    r = self.plan[j][2]
                    rows[j] = allm[r][idx[r]].tolist(); idx[r] += 1
            else:
                for i, j in enumerate(js):
                    rows[j] = maps[i] + [cks[i]]
            for j in ps:                            # the bit-offset scan carries on through the passes
                ph = pos & 7
                offs[j] = pos; nbits[j] = int(rows[j][ph]) - ph; cks_all[j] = int(rows[j][8]); pos += nbits[j]
            t2 = time.perf_counter()
            for r in range(self.world):             # every rank knows every rank's layout
                rj = [j for j in ps if self.plan[j][2] == r]
                o, span = span_layout([offs[j] & 7 for j in rj], [nbits[j] for j in rj])
                for i, j in enumerate(rj):
                    spans[r].append((j, acc_at[r] + o[i]))
                if r == self.rank and rj:
                    ptr, off, bits = self.ctx.chunks_emit([offs[j] & 7 for j in rj])
                    assert off == o and all(bits[i] - (offs[j] & 7) == nbits[j] for i, j in enumerate(rj))
                    self.acc[acc_at[r]: acc_at[r] + span].copy_(torch.as_tensor(_DevMem(ptr, span), device="cuda"))
                acc_at[r] += span
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            tm["compute"] += 1e3 * (t1 - t0); tm["all_gather"] += 1e3 * (t2 - t1); tm["emit"] += 1e3 * (t3 - t2)
        t3 = time.perf_counter()
        if self.world > 1:                          # exactly the produced bytes travel
            ops = []
            if self.rank == 0:
                for r in range(1, self.world):
                    if acc_at[r]:
                        ops.append(dist.P2POp(dist.irecv, self.stage[r][: acc_at[r]], r))
            elif acc_at[self.rank]:
                ops.append(dist.P2POp(dist.isend, self.acc[: acc_at[self.rank]], 0))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            torch.cuda.synchronize()
        t4 = time.perf_counter()
        if self.rank == 0:
            total_bytes = (pos + 7) // 8
            self.final_buf[: total_bytes + 8].zero_()
            torch.cuda.synchronize()
            src, dbit, nb = [], [], []
            for r in range(self.world):
                base = self.acc.data_ptr() if r == 0 else self.stage[r].data_ptr()
                for (j, o) in spans[r]:
                    src.append(base + o); dbit.append(offs[j]); nb.append(nbits[j])
            self.ctx.stitch_device(self.final_buf.data_ptr(), src, dbit, nb)
            self.last_out_bytes = total_bytes
            L = self.z.load()
            ck = cks_all[0]
            for j in range(1, len(self.plan)):
                ck = L.zultra_cuda_checksum_combine(self.flags, ck, cks_all[j], self.plan[j][1] - self.plan[j][0])
            self.checksum = ck
            self.final = self.final_buf[:total_bytes]
        e1.record(); torch.cuda.synchronize()
        t5 = time.perf_counter()
        tm["gather"] = 1e3 * (t4 - t3); tm["merge"] = 1e3 * (t5 - t4)
        self.breakdown = tm
        return e0.elapsed_time(e1)

    def final_stream(self):
        """Rank 0: the complete framed stream of the last step (header + deflate data + trailer) as bytes."""
        L = self.z.load()
        body = (self.dev_out[: self.last_out_bytes] if self.single else self.final).cpu().numpy().tobytes()
        hdr = (C.c_ubyte * 16)(); ftr = (C.c_ubyte * 16)()
        nh = L.zultra_frame_encode_header(hdr, 16, self.flags, None, 0)
        nf = L.zultra_frame_encode_footer(ftr, 16, C.c_uint(self.checksum), C.c_longlong(self.n_total), self.flags)
        return bytes(hdr[:nh]) + body + bytes(ftr[:nf])
