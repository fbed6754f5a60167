/* Host build of the pipeline (tests only): C entry points for ctypes. */
#include "zb_engine.h"

static ZbPipe g_shard_pipe;
extern "C" {
/* sharded operation, host build: same two calls as zultra_cuda_shard_prepare / _emit */
int emu_shard_prepare(const uint8_t *buf /* [hist|data] */, int hist_len, long n, unsigned block_size, int finalize, unsigned long long *bits8) {
   g_shard_pipe.st = 0;
   ZbStreamIn s = {0, (size_t)n, 0, (uint32_t)hist_len, finalize, 0, 0};
   std::vector<uint8_t> o; std::vector<ZbStreamRes> r;
   ZbRunOpts opt; opt.dev_in = buf; opt.phase = 1;
   if (zb_run_batch(g_shard_pipe, &s, 1, block_size, o, r, opt)) return -1;
   memcpy(bits8, opt.phase_bits, sizeof(opt.phase_bits));
   return 0;
}
long emu_shard_emit(int in_bits, uint8_t *out, long cap, unsigned long long *bits) {
   if (zb_finish_shard(g_shard_pipe, (uint32_t)in_bits, out, (size_t)cap, bits)) return -1;
   return (long)((*bits + 7) / 8);
}

/* several chunks of one buffer per call, host build: same two calls as zultra_cuda_chunks_prepare / _emit */
static ZbPipe g_chunk_pipe;
int emu_chunks_prepare(const uint8_t *buf, int nchunks, const size_t *off, const int *hist, const size_t *len, const int *fin, unsigned block_size, unsigned flags,
                       unsigned *cks, unsigned long long *maps8) {
   g_chunk_pipe.st = 0;
   const int kind = (flags & 2) ? 2 : ((flags & 1) ? 1 : 0);
   std::vector<ZbStreamIn> s((size_t)nchunks);
   for (int i = 0; i < nchunks; i++) { ZbStreamIn t = {0, len[i], 0, (uint32_t)hist[i], fin[i], 0, kind == 1 ? 1u : 0u, off[i]}; s[i] = t; }
   std::vector<uint8_t> o; std::vector<ZbStreamRes> r;
   ZbRunOpts opt; opt.dev_in = buf; opt.dev_offsets = 1; opt.phase = 1; opt.checksum_kind = kind;
   if (zb_run_batch(g_chunk_pipe, s.data(), nchunks, block_size, o, r, opt)) return -1;
   memcpy(maps8, opt.phase_maps.data(), sizeof(unsigned long long) * 8 * (size_t)nchunks);
   for (int i = 0; i < nchunks; i++) cks[i] = r[i].checksum;
   return 0;
}
long emu_chunks_emit(const unsigned *in_bits, uint8_t *out, long cap, size_t *out_off, unsigned long long *bits) {
   if (zb_finish_chunks(g_chunk_pipe, in_bits, out_off, bits)) return -1;
   size_t end = 0;
   for (size_t i = 0; i < g_chunk_pipe.plan.size(); i++) end = std::max(end, out_off[i] + (size_t)((bits[i] + 7) / 8));
   if ((long)end > cap) return -2;
   memcpy(out, g_chunk_pipe.out.p, end);
   return (long)end;
}

/* single stream, one call; returns bytes written (ceil(bits/8)), *bits = total bits */
long emu_compress(const uint8_t *data, long n, const uint8_t *hist, int hist_len, unsigned block_size, int finalize, int in_bits,
                  uint8_t *out, long out_cap, unsigned long long *bits, unsigned tile_main,
                  /* optional dumps, sized by caller: */ uint32_t *sa_lcp, uint16_t *match, int *nsub, int *sub_info /* 8 ints per sub */,
                  int *lit_len, int *off_len, uint16_t *best, unsigned *crc) {
   ZbPipe p; p.st = 0;
   if (getenv("ZB_EMU_PARSE_CD")) p.parse_cd = atoi(getenv("ZB_EMU_PARSE_CD"));      /* analysis aid: force the parse chunk length */
   ZbStreamIn s = {data, (size_t)n, hist, (uint32_t)hist_len, finalize, (uint32_t)in_bits, 0};
   std::vector<uint8_t> o; std::vector<ZbStreamRes> r; ZbDump d;
   ZbRunOpts opt; opt.tile_main = tile_main; opt.dump = &d; opt.checksum_kind = 2;
   if (zb_run_batch(p, &s, 1, block_size, o, r, opt)) return -1;
   if (crc) *crc = r[0].checksum;
   if ((long)o.size() > out_cap) return -2;
   memcpy(out, o.data(), o.size());
   *bits = r[0].total_bits;
   if (sa_lcp) memcpy(sa_lcp, d.sa_lcp.data(), d.sa_lcp.size() * 4);
   if (match) memcpy(match, d.match.data(), d.match.size() * 4);
   if (best) memcpy(best, d.best.data(), d.best.size() * 4);
   if (nsub) {
      *nsub = (int)d.sub.size();
      for (size_t i = 0; i < d.sub.size(); i++) {
         const ZbSub &b = d.sub[i];
         int *q = sub_info + 8 * i;
         q[0] = b.win; q[1] = b.ps; q[2] = b.pe; q[3] = b.is_dyn; q[4] = b.static_cost; q[5] = b.dynamic_cost; q[6] = b.body_bits; q[7] = b.stored | (b.ub_hit << 8);
         for (int j = 0; j < 288; j++) lit_len[288 * i + j] = d.tabs[i].llen[j];
         for (int j = 0; j < 32; j++) off_len[32 * i + j] = d.tabs[i].olen[j];
      }
   }
   p.release_all();
   return (long)o.size();
}
}


