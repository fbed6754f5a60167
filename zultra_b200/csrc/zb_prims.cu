/*
 * zb_prims.cu - cooperative sm_100a primitives: stable LSD radix sort (8-bit digits, warp-match ranking,
 * shared-memory histograms), device-wide scans (warp-shuffle), and the warp-per-tile suffix filter.
 *
 * These replace what libdivsufsort does on the CPU (reference divsufsort.c / sssort.c / trsort.c build the
 * suffix array with induced copying; here it is radix sort + prefix doubling, see zb_pipeline.h) and feed
 * the tile match finder.
 */
#include "zb_rt.h"

void zb_cuda_fail(cudaError_t e) {
   (void)e;
   /* no CPU fallback: a CUDA failure is fatal for the engine call; the pipeline stops at the next stage boundary
      (zb_failed) and the C-ABI layer reports it.  The flag belongs to the calling host thread: concurrent streams / lanes
      cannot clear or read each other's errors. */
   g_zb_cuda_error = 1;
}
thread_local int g_zb_cuda_error = 0;
long long g_zb_launches = 0;

/* ---- per-kernel timing ---- */
#include <vector>
#include <string>
#include <map>
#include <mutex>
int g_zb_prof_on = 0;
struct ZbProfRec { std::string tag; cudaEvent_t a, b; };
static std::vector<ZbProfRec> g_prof;
static std::mutex g_prof_mu;
/* several host threads (one per lane, zb_capi.cu) launch concurrently: the pending tag / record are per thread */
static thread_local const char *g_next_tag = 0;
static thread_local ZbProfRec g_cur;
void zb_tag(const char *tag) { g_next_tag = tag; }
void zb_prof_begin(int line, cudaStream_t st) {
   if (g_next_tag) g_cur.tag = g_next_tag; else { char b[32]; snprintf(b, sizeof b, "L%d", line); g_cur.tag = b; }
   g_next_tag = 0;
   cudaEventCreate(&g_cur.a); cudaEventCreate(&g_cur.b);
   cudaEventRecord(g_cur.a, st);
}
void zb_prof_end(cudaStream_t st) {
   cudaEventRecord(g_cur.b, st);
   std::lock_guard<std::mutex> g(g_prof_mu);
   g_prof.push_back(g_cur);
}
/* aggregate and clear; writes up to cap rows: name (32 bytes each), total ms, launches */
extern "C" int zultra_cuda_profile_collect(char *names, float *ms, int *counts, int cap) {
   std::map<std::string, std::pair<float, int> > agg;
   cudaDeviceSynchronize();
   for (size_t i = 0; i < g_prof.size(); i++) {
      float t = 0;
      cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b);
      agg[g_prof[i].tag].first += t; agg[g_prof[i].tag].second += 1;
      cudaEventDestroy(g_prof[i].a); cudaEventDestroy(g_prof[i].b);
   }
   g_prof.clear();
   int n = 0;
   for (std::map<std::string, std::pair<float, int> >::iterator it = agg.begin(); it != agg.end() && n < cap; ++it, ++n) {
      snprintf(names + 32 * n, 32, "%s", it->first.c_str());
      ms[n] = it->second.first; counts[n] = it->second.second;
   }
   return n;
}
extern "C" void zultra_cuda_profile(int on) { g_zb_prof_on = on; }
#define PROF_K(tag, st) do { if (g_zb_prof_on) { zb_tag(tag); zb_prof_begin(0, st); } } while (0)
#define PROF_E(st) do { if (g_zb_prof_on) zb_prof_end(st); } while (0)

void *zb_dev_alloc(size_t n) {
   void *p = 0;
   cudaError_t e = cudaMalloc(&p, n ? n : 256);
   if (e != cudaSuccess) { fprintf(stderr, "zultra-b200: cudaMalloc(%zu) failed: %s\n", n, cudaGetErrorString(e)); g_zb_cuda_error = 1; return 0; }
   return p;
}
void zb_dev_free(void *p) { if (p) cudaFree(p); }

/* ------------------------------------------------------------------ scans ------------------------------------------------------------------ */

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

struct OpSum { __device__ static uint32_t id() { return 0; } __device__ static uint32_t op(uint32_t a, uint32_t b) { return a + b; } };
struct OpMax { __device__ static uint32_t id() { return 0; } __device__ static uint32_t op(uint32_t a, uint32_t b) { return a > b ? a : b; } };

/* block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block total (all threads) */
template <class Op>
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *total) {
   __shared__ uint32_t wsum[SCAN_THREADS / 32];
   __shared__ uint32_t wtot;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   uint32_t inc = v;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = Op::op(inc, o);
   }
   if (lane == 31) wsum[warp] = inc;
   __syncthreads();
   if (warp == 0) {
      uint32_t w = lane < SCAN_THREADS / 32 ? wsum[lane] : Op::id();
      uint32_t winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         uint32_t o = __shfl_up_sync(0xffffffffu, winc, d);
         if (lane >= d) winc = Op::op(winc, o);
      }
      uint32_t wexc = __shfl_up_sync(0xffffffffu, winc, 1);
      if (lane == 0) wexc = Op::id();
      if (lane < SCAN_THREADS / 32) wsum[lane] = wexc;
      if (lane == SCAN_THREADS / 32 - 1) wtot = winc;
   }
   __syncthreads();
   uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
   if (lane == 0) exc = Op::id();
   exc = Op::op(exc, wsum[warp]);
   *total = wtot;
   __syncthreads();
   return exc;
}

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_k(const uint32_t *in, uint32_t *partial, long n) {
   const long base = (long)blockIdx.x * SCAN_TILE + (long)threadIdx.x * SCAN_ITEMS;
   uint32_t acc = Op::id();
#pragma unroll
   for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) acc = Op::op(acc, in[base + k]);
   uint32_t tot;
   block_exclusive<Op>(acc, &tot);
   if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

/* out = scan(in) seeded with prefix[blockIdx] (or identity if prefix == 0); optionally writes the grand total */
template <class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_k(const uint32_t *in, uint32_t *out, const uint32_t *prefix, long n, uint32_t *total_dev) {
   const long base = (long)blockIdx.x * SCAN_TILE + (long)threadIdx.x * SCAN_ITEMS;
   uint32_t v[SCAN_ITEMS];
   uint32_t acc = Op::id();
#pragma unroll
   for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : Op::id(); acc = Op::op(acc, v[k]); }
   uint32_t tot;
   uint32_t exc = block_exclusive<Op>(acc, &tot);
   uint32_t seed = prefix ? prefix[blockIdx.x] : Op::id();
   exc = Op::op(exc, seed);
#pragma unroll
   for (int k = 0; k < SCAN_ITEMS; k++) {
      uint32_t inc = Op::op(exc, v[k]);
      if (base + k < n) out[base + k] = INCLUSIVE ? inc : exc;
      exc = inc;
   }
   if (total_dev && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_dev = exc;
}

size_t zb_scan_scratch_words(long n) {
   size_t tot = 0;
   long m = n;
   while (m > SCAN_TILE) { m = (m + SCAN_TILE - 1) / SCAN_TILE; tot += (size_t)((m + 63) & ~63L); }
   return tot + 64;
}

template <class Op, bool INCLUSIVE>
static void scan_impl(cudaStream_t st, const uint32_t *in, uint32_t *out, long n, uint32_t *total_dev, uint32_t *scratch) {
   if (n <= 0) { if (total_dev) ZB_CUDA_CHECK(cudaMemsetAsync(total_dev, 0, 4, st)); return; }
   long nb = (n + SCAN_TILE - 1) / SCAN_TILE;
   if (nb == 1) {
      scan_apply_k<Op, INCLUSIVE><<<1, SCAN_THREADS, 0, st>>>(in, out, 0, n, total_dev);
      zb_count_launch(1);
   } else {
      scan_reduce_k<Op><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, scratch, n);
      scan_impl<Op, false>(st, scratch, scratch, nb, 0, scratch + ((nb + 63) & ~63L));
      scan_apply_k<Op, INCLUSIVE><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, scratch, n, total_dev);
      zb_count_launch(2);
   }
   ZB_CUDA_CHECK(cudaGetLastError());
}

void zb_exclusive_sum(zb_stream_t st, const uint32_t *in, uint32_t *out, long n, uint32_t *total_dev, uint32_t *scratch) {
   scan_impl<OpSum, false>(st, in, out, n, total_dev, scratch);
}
void zb_inclusive_max(zb_stream_t st, const uint32_t *in, uint32_t *out, long n, uint32_t *scratch) {
   scan_impl<OpMax, true>(st, in, out, n, 0, scratch);
}

/* --------------------------------------------------------------- radix sort --------------------------------------------------------------- */

#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_WARPS (RS_THREADS / 32)

__global__ void __launch_bounds__(RS_THREADS) rs_hist_k(const uint64_t *keys, long n, int shift, uint32_t mask, uint32_t *hist, int ntiles) {
   __shared__ uint32_t h[256];
   h[threadIdx.x] = 0;
   __syncthreads();
   const long base = (long)blockIdx.x * RS_TILE;
#pragma unroll
   for (int k = 0; k < RS_ITEMS; k++) {
      long idx = base + k * RS_THREADS + threadIdx.x;
      if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & mask], 1u);
   }
   __syncthreads();
   hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

/* One tile (2048 pairs) per CTA: rank every pair inside the tile (warp-level match.any, per-warp digit counters), move the
   pairs to their tile-local sorted place in shared memory, then write them out: pairs of one digit are consecutive there,
   so the global stores go out in runs instead of 32 scattered 8-byte writes per warp (the L2 sector rate was the limit). */
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_k(const uint64_t *kin, const uint32_t *vin, uint64_t *kout, uint32_t *vout, long n,
                                                          int shift, uint32_t mask, const uint32_t *hist_scanned, int ntiles) {
   __shared__ uint32_t cnt[RS_WARPS][256];
   __shared__ uint32_t gbase[256];      /* global start of this tile's run of digit d, minus the run's tile-local start */
   __shared__ uint32_t dstart[256];
   __shared__ uint64_t skey[RS_TILE];
   __shared__ uint32_t sval[RS_TILE];
   __shared__ uint32_t wsum[RS_WARPS];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
   for (int w = 0; w < RS_WARPS; w++) cnt[w][tid] = 0;
   const uint32_t gb = hist_scanned[(size_t)tid * ntiles + blockIdx.x];
   __syncthreads();
   const long base = (long)blockIdx.x * RS_TILE + (long)warp * (32 * RS_ITEMS);
   uint64_t key[RS_ITEMS];
   uint32_t val[RS_ITEMS], rk[RS_ITEMS];
   const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
   for (int k = 0; k < RS_ITEMS; k++) {
      const long idx = base + k * 32 + lane;
      const bool valid = idx < n;
      key[k] = valid ? kin[idx] : 0;
      val[k] = valid ? vin[idx] : 0;
   }
#pragma unroll
   for (int k = 0; k < RS_ITEMS; k++) {
      const long idx = base + k * 32 + lane;
      const bool valid = idx < n;
      const uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & mask) : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t pre = __popc(peers & lt);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers); }
      old = __shfl_sync(0xffffffffu, old, leader);
      rk[k] = old + pre;
      __syncwarp();
   }
   __syncthreads();
   /* digit tid: exclusive prefix over the warps, tile total, then exclusive scan of the totals over the digits */
   uint32_t tot = 0;
#pragma unroll
   for (int w = 0; w < RS_WARPS; w++) { uint32_t t = cnt[w][tid]; cnt[w][tid] = tot; tot += t; }
   uint32_t inc = tot;
#pragma unroll
   for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
   if (lane == 31) wsum[warp] = inc;
   __syncthreads();
   uint32_t woff = 0;
#pragma unroll
   for (int w = 0; w < RS_WARPS; w++) if (w < warp) woff += wsum[w];
   const uint32_t ds = woff + inc - tot;
   dstart[tid] = ds;
   gbase[tid] = gb - ds;
   __syncthreads();
#pragma unroll
   for (int k = 0; k < RS_ITEMS; k++) {
      const long idx = base + k * 32 + lane;
      if (idx < n) {
         const uint32_t d = (uint32_t)(key[k] >> shift) & mask;
         const uint32_t lp = dstart[d] + cnt[warp][d] + rk[k];
         skey[lp] = key[k]; sval[lp] = val[k];
      }
   }
   __syncthreads();
   const long tile_n = n - (long)blockIdx.x * RS_TILE < RS_TILE ? n - (long)blockIdx.x * RS_TILE : RS_TILE;
#pragma unroll
   for (int k = 0; k < RS_ITEMS; k++) {
      const int e = k * RS_THREADS + tid;
      if (e < tile_n) {
         const uint64_t kk = skey[e];
         const uint32_t d = (uint32_t)(kk >> shift) & mask;
         const size_t dst = (size_t)(uint32_t)(gbase[d] + (uint32_t)e);   /* 32-bit wrap: gbase holds start - local start */
         kout[dst] = kk;
         vout[dst] = sval[e];
      }
   }
}

size_t zb_sort_scratch_words(long n) {
   long ntiles = (n + RS_TILE - 1) / RS_TILE;
   if (ntiles < 1) ntiles = 1;
   return (size_t)256 * ntiles + zb_scan_scratch_words(256 * ntiles) + 64;
}

void zb_sort_pairs(zb_stream_t st, uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, long n, int bit_lo, int bit_hi, uint32_t *scratch) {
   if (n <= 1 || bit_hi <= bit_lo) return;
   const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
   uint32_t *hist = scratch;
   uint32_t *scan_scratch = scratch + (size_t)256 * ntiles;
   uint64_t *kin = keys, *kout = keys_tmp;
   uint32_t *vin = vals, *vout = vals_tmp;
   int npass = 0;
   for (int shift = bit_lo; shift < bit_hi; shift += 8, npass++) {
      const int w = (bit_hi - shift) < 8 ? (bit_hi - shift) : 8;
      const uint32_t mask = (1u << w) - 1u;
      PROF_K("rs_hist", st);
      rs_hist_k<<<ntiles, RS_THREADS, 0, st>>>(kin, n, shift, mask, hist, ntiles);
      PROF_E(st);
      scan_impl<OpSum, false>(st, hist, hist, (long)256 * ntiles, 0, scan_scratch);
      PROF_K("rs_scatter", st);
      rs_scatter_k<<<ntiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, mask, hist, ntiles);
      PROF_E(st);
      zb_count_launch(2);
      uint64_t *tk = kin; kin = kout; kout = tk;
      uint32_t *tv = vin; vin = vout; vout = tv;
   }
   ZB_CUDA_CHECK(cudaGetLastError());
   if (npass & 1) { /* result sits in the tmp buffers */
      ZB_CUDA_CHECK(cudaMemcpyAsync(keys, keys_tmp, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
      ZB_CUDA_CHECK(cudaMemcpyAsync(vals, vals_tmp, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
   }
}

/* --------------------------------------------------------------- tile filter --------------------------------------------------------------- */

/* Every tile's source list is cut into nseg segments, one warp each, so that the sequential part of a stream is short:
   pass 1 counts the kept entries of every segment and the running LCP minimum it leaves behind, a per-tile scan turns that
   into output offsets and carried-in minima, pass 2 streams again and writes. */
__device__ __forceinline__ void tile_seg_range(const ZbTileDesc &t, const uint32_t *src_cnt, int nseg, int sg, uint32_t &r_lo, uint32_t &r_hi) {
   const uint32_t n = t.src_cnt_idx >= 0 ? src_cnt[t.src_cnt_idx] : t.src_n;
   const uint32_t seglen = (((n + (uint32_t)nseg - 1u) / (uint32_t)nseg) + 31u) & ~31u;
   r_lo = (uint32_t)sg * seglen; if (r_lo > n) r_lo = n;
   r_hi = r_lo + seglen; if (r_hi > n) r_hi = n;
}

__global__ void __launch_bounds__(128) tile_filter_count_k(const uint32_t *srcw, const uint32_t *src_cnt, const ZbTileDesc *tiles, int ntiles, int first_tile, int nseg, uint32_t *seg) {
   const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
   if (wid >= (long)ntiles * nseg) return;
   const int lane = threadIdx.x & 31;
   const ZbTileDesc t = tiles[first_tile + (int)(wid / nseg)];
   const uint32_t *src = srcw + t.src_base;
   uint32_t r_lo, r_hi;
   tile_seg_range(t, src_cnt, nseg, (int)(wid % nseg), r_lo, r_hi);
   const uint32_t lo = t.lo - t.src_lo, span = t.hi - t.lo;
   uint32_t count = 0, tail = 0x1ffu, has = 0;
   for (uint32_t r0 = r_lo; r0 < r_hi; r0 += 32) {
      const uint32_t r = r0 + lane;
      const bool valid = r < r_hi;
      const uint32_t w = valid ? __ldg(src + r) : 0u;
      const uint32_t v = valid ? ((w >> ZB_POS_BITS) & 0x1ffu) : 0x1ffu;
      const bool keep = valid && ((w & ZB_POS_MASK) - lo) < span;
      const uint32_t kmask = __ballot_sync(0xffffffffu, keep);
      if (kmask) {
         const int lastk = 31 - __clz((int)kmask);
         tail = __reduce_min_sync(0xffffffffu, lane > lastk ? v : 0x1ffu);
         has = 1; count += __popc(kmask);
      } else {
         const uint32_t mn = __reduce_min_sync(0xffffffffu, v);
         tail = mn < tail ? mn : tail;
      }
   }
   if (lane == 0) { seg[2 * wid] = count; seg[2 * wid + 1] = tail | (has << 16); }
}

__global__ void tile_filter_offsets_k(uint32_t *seg, int ntiles, int nseg, uint32_t *cnt) {
   const int k = blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= ntiles) return;
   uint32_t acc = 0, carry = 0x1ffu;
   for (int sg = 0; sg < nseg; sg++) {
      const size_t i = (size_t)k * nseg + sg;
      const uint32_t c = seg[2 * i], th = seg[2 * i + 1];
      seg[2 * i] = acc; seg[2 * i + 1] = carry;
      acc += c;
      const uint32_t tl = th & 0xffffu;
      carry = (th >> 16) ? tl : (tl < carry ? tl : carry);
   }
   cnt[k] = acc;
}

__global__ void __launch_bounds__(128) tile_filter_k(const uint32_t *srcw, const uint32_t *src_cnt, const ZbTileDesc *tiles, int ntiles, int first_tile, uint32_t *out, size_t stride,
                                                     int nseg, const uint32_t *seg) {
   const long wid = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
   if (wid >= (long)ntiles * nseg) return;
   const int lane = threadIdx.x & 31;
   const int tile = (int)(wid / nseg);
   const ZbTileDesc t = tiles[first_tile + tile];
   const uint32_t *src = srcw + t.src_base;
   uint32_t *dst = out + (size_t)tile * stride;
   uint32_t r_lo, r_hi;
   tile_seg_range(t, src_cnt, nseg, (int)(wid % nseg), r_lo, r_hi);
   const uint32_t lo = t.lo - t.src_lo, span = t.hi - t.lo;
   const uint32_t lt = (1u << lane) - 1u;
   uint32_t carry = seg[2 * wid + 1], count = seg[2 * wid];
   for (uint32_t r0 = r_lo; r0 < r_hi; r0 += 32) {
      const uint32_t r = r0 + lane;
      const bool valid = r < r_hi;
      const uint32_t w = valid ? __ldg(src + r) : 0u;
      const uint32_t pos = (w & ZB_POS_MASK) - lo;
      uint32_t v = valid ? ((w >> ZB_POS_BITS) & 0x1ffu) : 0x1ffu;
      const bool keep = valid && pos < span;
      const uint32_t kmask = __ballot_sync(0xffffffffu, keep);
      const uint32_t below = kmask & lt;
      const int start = below ? (32 - __clz((int)below)) : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
         uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
         if (lane >= start + d) v = min(v, o);
      }
      if (start == 0) v = min(v, carry);
      if (keep) dst[count + __popc(below)] = pos | (v << ZB_POS_BITS);
      const uint32_t v31 = __shfl_sync(0xffffffffu, v, 31);
      carry = (kmask >> 31) ? 0x1ffu : v31;
      count += __popc(kmask);
   }
}

void zb_tile_filter(zb_stream_t st, const uint32_t *src, const uint32_t *src_cnt, const ZbTileDesc *tiles, int ntiles, int first_tile, uint32_t *out, size_t stride, uint32_t *cnt,
                    int nseg, uint32_t *seg_scratch) {
   if (ntiles <= 0) return;
   if (nseg < 1) nseg = 1;
   const int wpb = 4;
   const long nwarp = (long)ntiles * nseg;
   PROF_K("mf_tile_filter", st);
   tile_filter_count_k<<<(unsigned)((nwarp + wpb - 1) / wpb), wpb * 32, 0, st>>>(src, src_cnt, tiles, ntiles, first_tile, nseg, seg_scratch);
   tile_filter_offsets_k<<<(ntiles + 127) / 128, 128, 0, st>>>(seg_scratch, ntiles, nseg, cnt);
   tile_filter_k<<<(unsigned)((nwarp + wpb - 1) / wpb), wpb * 32, 0, st>>>(src, src_cnt, tiles, ntiles, first_tile, out, stride, nseg, seg_scratch);
   PROF_E(st);
   zb_count_launch(3);
   ZB_CUDA_CHECK(cudaGetLastError());
}
