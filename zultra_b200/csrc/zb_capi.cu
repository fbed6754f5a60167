/*
 * zb_capi.cu - the C-ABI of include/zultra_cuda.h over the CUDA pipeline.
 */
#include "zb_engine.h"
#include "../../include/zultra_cuda.h"
#include <new>
#include <mutex>

extern int g_zb_cuda_error;

struct zultra_cuda_ctx_s {
   int device;
   unsigned tile = 0;
   ZbPipe pipe;
   float ms[8];
   long long counters[8];
   std::vector<uint8_t> out;
};

static int ctx_enter(zultra_cuda_ctx_t *c) {
   if (!c) return ZULTRA_CUDA_ERR_ARG;
   if (cudaSetDevice(c->device) != cudaSuccess) return ZULTRA_CUDA_ERR_CUDA;
   g_zb_cuda_error = 0;
   return 0;
}
static int ctx_leave(zultra_cuda_ctx_t *c, int rc) {
   if (g_zb_cuda_error) { cudaGetLastError(); return ZULTRA_CUDA_ERR_CUDA; }
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) { fprintf(stderr, "zultra-b200: CUDA error %s\n", cudaGetErrorString(e)); return ZULTRA_CUDA_ERR_CUDA; }
   (void)c;
   return rc;
}
static void fill_counters(zultra_cuda_ctx_t *c, long long launches0) {
   ZbPipe &p = c->pipe;
   c->counters[0] = p.nwin; c->counters[1] = p.nsub; c->counters[2] = p.stat_sa_rounds;
   c->counters[4] = g_zb_launches - launches0;
   c->counters[3] = p.stat_redo; c->counters[7] = p.stat_tiles;
}

extern "C" {

int zultra_cuda_device_count(void) {
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
   return n;
}

int zultra_cuda_ctx_create(zultra_cuda_ctx_t **pp, int device) {
   if (!pp) return ZULTRA_CUDA_ERR_ARG;
   *pp = 0;
   int n = zultra_cuda_device_count();
   if (n <= 0) return ZULTRA_CUDA_ERR_NODEVICE;
   if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
   if (device >= n) return ZULTRA_CUDA_ERR_NODEVICE;
   if (cudaSetDevice(device) != cudaSuccess) return ZULTRA_CUDA_ERR_CUDA;
   zultra_cuda_ctx_t *c = new (std::nothrow) zultra_cuda_ctx_t();
   if (!c) return ZULTRA_CUDA_ERR_ARG;
   c->device = device;
   memset(c->ms, 0, sizeof(c->ms)); memset(c->counters, 0, sizeof(c->counters));
   if (cudaStreamCreateWithFlags(&c->pipe.st, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ZULTRA_CUDA_ERR_CUDA; }
   {  /* tuning knobs (never change the output) */
      const char *e;
      if ((e = getenv("ZULTRA_CUDA_PARSE_CD")) && atoi(e) >= 512) c->pipe.parse_cd = atoi(e);
      if ((e = getenv("ZULTRA_CUDA_PARSE_WU")) && atoi(e) >= 258) c->pipe.parse_wu = atoi(e);
      if ((e = getenv("ZULTRA_CUDA_TILE")) && atoi(e) >= 256) c->tile = (unsigned)atoi(e);
   }
   *pp = c;
   return 0;
}

void zultra_cuda_ctx_destroy(zultra_cuda_ctx_t *c) {
   if (!c) return;
   cudaSetDevice(c->device);
   c->pipe.release_all();
   cudaStreamDestroy(c->pipe.st);
   delete c;
}

static int run_one(zultra_cuda_ctx_t *c, const ZbStreamIn &s, unsigned block, ZbRunOpts &o, std::vector<ZbStreamRes> &res) {
   long long l0 = g_zb_launches;
   zb_memset(c->pipe.st, 0, 0, 0);
   c->pipe.counters.need(64);
   zb_memset(c->pipe.st, c->pipe.counters.p, 0, 64 * 4);
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   cudaEventRecord(e0, c->pipe.st);
   if (!o.tile_main) o.tile_main = c->tile;
   c->pipe.stat_redo = 0;
   int rc = zb_run_batch(c->pipe, &s, 1, block, c->out, res, o);
   cudaEventRecord(e1, c->pipe.st);
   cudaEventSynchronize(e1);
   memcpy(c->ms, o.ms, sizeof(c->ms));
   { float t = 0; if (cudaEventElapsedTime(&t, e0, e1) == cudaSuccess) c->ms[7] = t; }   /* device-side total, CUDA events on the launching stream */
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   fill_counters(c, l0);
   return rc;
}

static unsigned clamp_block(unsigned b) {
   if (!b) b = 1048576;
   if (b < 32768) b = 32768;
   if (b > 2097152) b = 2097152;
   return b;
}

int zultra_cuda_compress_blocks(zultra_cuda_ctx_t *c, const unsigned char *hist, int hist_size, const unsigned char *in, size_t n,
                                unsigned int block, int finalize, unsigned int in_bits, unsigned int flags, unsigned int *checksum,
                                unsigned char *out, size_t out_cap, unsigned long long *out_bits) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   if (!in || !out || !out_bits || in_bits > 7 || hist_size < 0 || hist_size > ZB_HISTORY) return ZULTRA_CUDA_ERR_ARG;
   ZbStreamIn s = {in, n, hist, (uint32_t)hist_size, finalize, in_bits, checksum ? *checksum : 0};
   ZbRunOpts o;
   o.checksum_kind = (flags & 2) ? 2 : ((flags & 1) ? 1 : 0);
   std::vector<ZbStreamRes> res;
   rc = run_one(c, s, clamp_block(block), o, res);
   if (rc == 0) {
      if (c->out.size() > out_cap) rc = ZULTRA_CUDA_ERR_DST;
      else { memcpy(out, c->out.data(), c->out.size()); *out_bits = res[0].total_bits; if (checksum) *checksum = res[0].checksum; }
   } else rc = ZULTRA_CUDA_ERR_CUDA;
   return ctx_leave(c, rc);
}

int zultra_cuda_shard_prepare(zultra_cuda_ctx_t *c, const void *dev_in, int hist_size, size_t n, unsigned int block, int finalize, unsigned int flags,
                              unsigned int *checksum, unsigned long long *phase_bits8) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   if (!dev_in || !phase_bits8 || hist_size < 0 || hist_size > ZB_HISTORY) return ZULTRA_CUDA_ERR_ARG;
   ZbStreamIn s = {0, n, 0, (uint32_t)hist_size, finalize, 0, checksum ? *checksum : 0};
   ZbRunOpts o;
   o.dev_in = (const uint8_t *)dev_in; o.phase = 1;
   o.checksum_kind = (flags & 2) ? 2 : ((flags & 1) ? 1 : 0);
   std::vector<ZbStreamRes> res;
   rc = run_one(c, s, clamp_block(block), o, res);
   if (rc == 0) { memcpy(phase_bits8, o.phase_bits, sizeof(o.phase_bits)); if (checksum) *checksum = res[0].checksum; }
   return ctx_leave(c, rc ? ZULTRA_CUDA_ERR_CUDA : 0);
}

int zultra_cuda_shard_emit(zultra_cuda_ctx_t *c, unsigned int in_bits, void *dev_out, size_t out_cap, unsigned long long *out_bits) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   if (!dev_out || !out_bits || in_bits > 7) return ZULTRA_CUDA_ERR_ARG;
   long long l0 = g_zb_launches;
   rc = zb_finish_shard(c->pipe, in_bits, (uint8_t *)dev_out, out_cap, out_bits);
   c->counters[4] += g_zb_launches - l0;
   return ctx_leave(c, rc == -2 ? ZULTRA_CUDA_ERR_DST : (rc ? ZULTRA_CUDA_ERR_CUDA : 0));
}

unsigned int zultra_cuda_checksum_combine(unsigned int flags, unsigned int ck1, unsigned int ck2, unsigned long long len2) {
   if (flags & 2) return zb_crc32_combine(ck1, ck2, len2);
   if (flags & 1) {   /* adler32 of A||B from adler(A), adler(B) (B started from 1), len(B) */
      const unsigned long long BASE = 65521;
      unsigned long long rem = len2 % BASE, sum1 = ck1 & 0xffff, sum2 = (rem * sum1) % BASE;
      sum1 += (ck2 & 0xffff) + BASE - 1;
      sum2 += ((ck1 >> 16) & 0xffff) + ((ck2 >> 16) & 0xffff) + BASE - rem;
      if (sum1 >= BASE) sum1 -= BASE;
      if (sum1 >= BASE) sum1 -= BASE;
      if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
      if (sum2 >= BASE) sum2 -= BASE;
      return (unsigned int)(sum1 | (sum2 << 16));
   }
   return 0;
}

int zultra_cuda_compress_blocks_device(zultra_cuda_ctx_t *c, const void *dev_in, size_t n, unsigned int block, int finalize,
                                       unsigned int flags, unsigned int *checksum, void *dev_out, size_t out_cap, unsigned long long *out_bits) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   if (!dev_in || !dev_out || !out_bits) return ZULTRA_CUDA_ERR_ARG;
   ZbStreamIn s = {0, n, 0, 0, finalize, 0, checksum ? *checksum : 0};
   ZbRunOpts o;
   o.dev_in = (const uint8_t *)dev_in; o.dev_out = (uint8_t *)dev_out; o.dev_out_cap = out_cap;
   o.checksum_kind = (flags & 2) ? 2 : ((flags & 1) ? 1 : 0);
   std::vector<ZbStreamRes> res;
   rc = run_one(c, s, clamp_block(block), o, res);
   if (rc == 0) { *out_bits = res[0].total_bits; if (checksum) *checksum = res[0].checksum; }
   else rc = rc == -2 ? ZULTRA_CUDA_ERR_DST : ZULTRA_CUDA_ERR_CUDA;
   return ctx_leave(c, rc);
}

/* ---- context pool: contexts (device buffers, stream) are expensive to build; the libzultra front end borrows them ---- */
static std::mutex g_pool_mu;
static std::vector<zultra_cuda_ctx_t *> g_pool;

int zultra_cuda_ctx_acquire(zultra_cuda_ctx_t **pp, int device) {
   if (!pp) return ZULTRA_CUDA_ERR_ARG;
   if (device < 0) { if (zultra_cuda_device_count() <= 0) return ZULTRA_CUDA_ERR_NODEVICE; if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
   {
      std::lock_guard<std::mutex> g(g_pool_mu);
      for (size_t i = 0; i < g_pool.size(); i++)
         if (g_pool[i]->device == device) { *pp = g_pool[i]; g_pool.erase(g_pool.begin() + i); return 0; }
   }
   return zultra_cuda_ctx_create(pp, device);
}
void zultra_cuda_ctx_release(zultra_cuda_ctx_t *c) {
   if (!c) return;
   std::lock_guard<std::mutex> g(g_pool_mu);
   if (g_pool.size() < 4) g_pool.push_back(c); else zultra_cuda_ctx_destroy(c);
}
void zultra_cuda_release_cached(void) {
   std::lock_guard<std::mutex> g(g_pool_mu);
   for (size_t i = 0; i < g_pool.size(); i++) zultra_cuda_ctx_destroy(g_pool[i]);
   g_pool.clear();
}

/* frame bytes written on the host side of the batch call (frame.c:387-452, :509-547) */
static size_t put_header(unsigned char *o, unsigned flags) {
   if (flags & 2) { const unsigned char h[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 2, 255}; memcpy(o, h, 10); return 10; }
   if (flags & 1) { o[0] = 0x78; o[1] = 0xda; return 2; }
   return 0;
}
static size_t put_footer(unsigned char *o, unsigned flags, unsigned ck, unsigned long long n) {
   if (flags & 2) { for (int i = 0; i < 4; i++) { o[i] = (unsigned char)(ck >> (8 * i)); o[4 + i] = (unsigned char)(n >> (8 * i)); } return 8; }
   if (flags & 1) { for (int i = 0; i < 4; i++) o[i] = (unsigned char)(ck >> (8 * (3 - i))); return 4; }
   return 0;
}

int zultra_cuda_memory_compress_batch(zultra_cuda_ctx_t *c, const unsigned char *const *in, const size_t *in_sizes, unsigned char *const *outp,
                                      const size_t *out_caps, size_t *out_sizes, size_t nstreams, unsigned int flags, unsigned int block) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   block = clamp_block(block);
   long long l0 = g_zb_launches;
   const unsigned hdr = (flags & 2) ? 10 : ((flags & 1) ? 2 : 0), ftr = (flags & 2) ? 8 : ((flags & 1) ? 4 : 0);
   float ms[8] = {0};
   size_t i0 = 0;
   while (i0 < nstreams) {
      /* sub-batch: bounded positions */
      size_t i1 = i0, bytes = 0;
      std::vector<ZbStreamIn> s;
      while (i1 < nstreams && (i1 == i0 || bytes + in_sizes[i1] + ((in_sizes[i1] / block) * ZB_HISTORY) < ((size_t)768 << 20))) {
         if (in_sizes[i1] > 0) {   /* empty input yields (size_t)-1 like the reference (libzultra.c:275,617) */
            ZbStreamIn t = {in[i1], in_sizes[i1], 0, 0, 1, 0, (flags & 2) ? 0u : 1u};
            s.push_back(t);
            bytes += in_sizes[i1] + (in_sizes[i1] / block) * ZB_HISTORY;
         }
         i1++;
      }
      std::vector<ZbStreamRes> res;
      ZbRunOpts o;
      o.checksum_kind = (flags & 2) ? 2 : ((flags & 1) ? 1 : 0);
      c->pipe.counters.need(64);
      zb_memset(c->pipe.st, c->pipe.counters.p, 0, 64 * 4);
      if (!s.empty() && zb_run_batch(c->pipe, s.data(), (int)s.size(), block, c->out, res, o)) return ctx_leave(c, ZULTRA_CUDA_ERR_CUDA);
      for (int k = 0; k < 8; k++) ms[k] += o.ms[k];
      size_t k = 0;
      for (size_t i = i0; i < i1; i++) {
         if (in_sizes[i] == 0) { out_sizes[i] = (size_t)-1; continue; }
         const size_t nb = (size_t)((res[k].total_bits + 7) / 8);
         if (hdr + nb + ftr > out_caps[i]) out_sizes[i] = (size_t)-1;
         else {
            size_t w = put_header(outp[i], flags);
            memcpy(outp[i] + w, c->out.data() + res[k].out_off, nb); w += nb;
            w += put_footer(outp[i] + w, flags, res[k].checksum, in_sizes[i]);
            out_sizes[i] = w;
         }
         k++;
      }
      i0 = i1;
   }
   memcpy(c->ms, ms, sizeof(ms));
   fill_counters(c, l0);
   return ctx_leave(c, 0);
}

int zultra_cuda_checksum_device(zultra_cuda_ctx_t *c, const void *dev, size_t n, unsigned int flags, unsigned int *ck) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   const int kind = (flags & 2) ? 2 : ((flags & 1) ? 1 : 0);
   if (!kind) { *ck = 0; return 0; }
   c->pipe.in_ptr = (const uint8_t *)dev;
   std::vector<uint64_t> off(1, 0), len(1, n); std::vector<uint32_t> sums;
   c->pipe.stage_checksum(kind, off, len, sums, *ck);
   *ck = sums[0];
   return ctx_leave(c, 0);
}

int zultra_cuda_window_sa_lcp(zultra_cuda_ctx_t *c, const unsigned char *win, int n, unsigned int *words) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   ZbStreamIn s = {win, (size_t)n, 0, 0, 1, 0, 0};
   ZbRunOpts o; ZbDump d; o.dump = &d; o.stop_after = 1;
   std::vector<ZbStreamRes> res;
   rc = run_one(c, s, 2097152u + 65536u, o, res);   /* one window: block size above any window */
   if (rc == 0) memcpy(words, d.sa_lcp.data(), (size_t)n * 4);
   return ctx_leave(c, rc ? ZULTRA_CUDA_ERR_CUDA : 0);
}

int zultra_cuda_window_matches(zultra_cuda_ctx_t *c, const unsigned char *win, int hist, int n, unsigned short *matches, unsigned int tile) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   ZbStreamIn s = {win + hist, (size_t)(n - hist), win, (uint32_t)hist, 1, 0, 0};
   ZbRunOpts o; ZbDump d; o.dump = &d; o.stop_after = 2; o.tile_main = tile;
   std::vector<ZbStreamRes> res;
   rc = run_one(c, s, 2097152u + 65536u, o, res);
   if (rc == 0) memcpy(matches, d.match.data() + (size_t)hist * 8, (size_t)(n - hist) * 8 * sizeof(zb_match_t));
   return ctx_leave(c, rc ? ZULTRA_CUDA_ERR_CUDA : 0);
}

int zultra_cuda_block_stages(zultra_cuda_ctx_t *c, const unsigned char *win, int hist, int n, int *info, int *lit_len, int *off_len, unsigned short *best) {
   int rc = ctx_enter(c);
   if (rc) return rc;
   ZbStreamIn s = {win + hist, (size_t)(n - hist), win, (uint32_t)hist, 0, 0, 0};
   ZbRunOpts o; ZbDump d; o.dump = &d;
   std::vector<ZbStreamRes> res;
   rc = run_one(c, s, 2097152u + 65536u, o, res);
   if (rc) return ctx_leave(c, ZULTRA_CUDA_ERR_CUDA);
   for (size_t i = 0; i < d.sub.size(); i++) {
      const ZbSub &b = d.sub[i];
      int *q = info + 8 * i;
      q[0] = b.ps; q[1] = b.pe; q[2] = b.is_dyn; q[3] = b.static_cost; q[4] = b.dynamic_cost; q[5] = b.body_bits; q[6] = b.stored; q[7] = b.mask | (b.ub_hit << 8);
      for (int j = 0; j < 288; j++) lit_len[288 * i + j] = d.tabs[i].llen[j];
      for (int j = 0; j < 32; j++) off_len[32 * i + j] = d.tabs[i].olen[j];
   }
   if (best) memcpy(best, d.best.data(), (size_t)n * sizeof(zb_match_t));
   return ctx_leave(c, (int)d.sub.size());
}

int zultra_cuda_last_timings(zultra_cuda_ctx_t *c, float *ms) { if (!c) return ZULTRA_CUDA_ERR_ARG; memcpy(ms, c->ms, sizeof(c->ms)); return 0; }
int zultra_cuda_last_counters(zultra_cuda_ctx_t *c, long long *v) { if (!c) return ZULTRA_CUDA_ERR_ARG; memcpy(v, c->counters, sizeof(c->counters)); return 0; }

}
