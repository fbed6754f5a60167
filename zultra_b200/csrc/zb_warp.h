/*
 * zb_warp.h - warp-cooperative forms of the per-sub-block / per-splitter-node table work (CUDA build only).
 *
 * The scalar task bodies of zb_core.h (zb_huff_lengths, zb_huff_build, zb_dynamic_cost, range histograms) are exact
 * restatements of the reference's sequential code (huffencoder.c:157-375, blockdeflate.c:577-618), but one GPU thread runs
 * ~50 000 dependent instructions per task at one instruction every few cycles, and these tasks sit on the critical path of
 * every splitter level and every parse pass.  Here ONE WARP runs a task: everything that is a map, a reduction, a compaction
 * or a sort goes over the 32 lanes; what is inherently a chain (the Moffat-Katajainen pairing, the run-length tokeniser, the
 * length-limit adjustment) stays on lane 0 but works on shared memory.  Same integer results as the scalar forms - the host
 * build keeps running those, and the -m gpu parity tests diff split offsets, cost estimates and code lengths of this path
 * against the reference.
 */
#ifndef ZB_WARP_H
#define ZB_WARP_H
#ifndef ZB_EMU

struct ZbWarpScratch {            /* per warp, in shared memory */
   int h[ZB_NH];                  /* a histogram being worked on: 288 literal/length + 32 distance counts */
   int llen[ZB_NLIT], olen[ZB_NLIT];
   uint32_t key[ZB_NLIT], tmp[ZB_NLIT];
   int16_t order[ZB_NLIT];
   uint32_t bins[32];
   ZbScratch sc;                  /* scalar scratch for the tails that stay on lane 0 */
};

#define ZBW_FULL 0xffffffffu

__device__ __forceinline__ int zbw_sum(int v) {
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(ZBW_FULL, v, d);
   return v;
}

/* stable sort of key[0..n) (n <= 288) ascending on bits [9, 9 + bits): LSD, 5 bits a pass = one bin per lane */
__device__ __forceinline__ void zbw_sort(uint32_t *key, uint32_t *tmp, uint32_t *bins, int n, int bits, const int lane) {
   uint32_t *src = key, *dst = tmp;
   const uint32_t lt = (1u << lane) - 1u;
   for (int sh = 9; sh < 9 + bits; sh += 5) {
      bins[lane] = 0;
      __syncwarp();
      for (int b = 0; b < n; b += 32) {
         const int e = b + lane;
         const uint32_t d = e < n ? ((src[e] >> sh) & 31u) : 0xffffffffu;
         const uint32_t peers = __match_any_sync(ZBW_FULL, d);
         if (e < n && lane == __ffs((int)peers) - 1) bins[d] += (uint32_t)__popc(peers);
         __syncwarp();
      }
      /* exclusive scan of the 32 bins */
      const uint32_t c = bins[lane];
      uint32_t inc = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(ZBW_FULL, inc, d); if (lane >= d) inc += o; }
      __syncwarp();
      bins[lane] = inc - c;
      __syncwarp();
      for (int b = 0; b < n; b += 32) {
         const int e = b + lane;
         const uint32_t v = e < n ? src[e] : 0u;
         const uint32_t d = e < n ? ((v >> sh) & 31u) : 0xffffffffu;
         const uint32_t peers = __match_any_sync(ZBW_FULL, d);
         if (e < n) dst[bins[d] + (uint32_t)__popc(peers & lt)] = v;
         __syncwarp();
         if (e < n && lane == __ffs((int)peers) - 1) bins[d] += (uint32_t)__popc(peers);
         __syncwarp();
      }
      uint32_t *t = src; src = dst; dst = t;
   }
   if (src != key) { for (int e = lane; e < n; e += 32) key[e] = src[e]; __syncwarp(); }
}

/* zb_huff_lengths (huffencoder.c:157-270): cnt[0..nsym) -> len[0..288), all in shared (or cnt in global) memory */
__device__ __forceinline__ void zbw_lengths(const int *cnt, int nsym, int *len, ZbWarpScratch &s, const int lane) {
   uint32_t *key = s.key;
   const uint32_t lt = (1u << lane) - 1u;
   for (int i = lane; i < ZB_NLIT; i += 32) len[i] = 0;
   int n = 0;
   uint32_t top = 0;
   for (int b = 0; b < nsym; b += 32) {      /* used symbols, in symbol order */
      const int i = b + lane;
      const int c = i < nsym ? cnt[i] : 0;
      const uint32_t mk = __ballot_sync(ZBW_FULL, c != 0);
      if (c) { key[n + __popc(mk & lt)] = ((uint32_t)c << 9) | (uint32_t)i; top |= (uint32_t)c; }
      n += __popc(mk);
   }
   __syncwarp();
   if (n <= 1) { if (lane == 0) len[0] = 1; __syncwarp(); return; }
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) top |= __shfl_xor_sync(ZBW_FULL, top, d);
   zbw_sort(key, s.tmp, s.bins, n, 32 - __clz((int)top), lane);
   if (lane == 0) {
      /* Moffat-Katajainen in place, as zb_huff_lengths: w[] in the upper bits of key[], symbol ids in the low 9 */
#define ZB_W(i) (key[i] >> 9)
#define ZB_SETW(i, v) (key[i] = ((uint32_t)(v) << 9) | (key[i] & 511u))
      int leaf = 0, inode = 0;
      for (int t = 0; t < n - 1; t++) {
         uint32_t w;
         if (leaf >= n || (inode < t && ZB_W(inode) < ZB_W(leaf))) { w = ZB_W(inode); ZB_SETW(inode, t); inode++; }
         else { w = ZB_W(leaf); leaf++; }
         if (leaf >= n || (inode < t && ZB_W(inode) < ZB_W(leaf))) { w += ZB_W(inode); ZB_SETW(inode, t); inode++; }
         else { w += ZB_W(leaf); leaf++; }
         ZB_SETW(t, w);
      }
      ZB_SETW(n - 2, 0);
      for (int t = n - 3; t >= 0; t--) ZB_SETW(t, ZB_W(ZB_W(t)) + 1);
      int avail = 1, used = 0, depth = 0, t = n - 2, x = n - 1;
      while (avail > 0) {
         while (t >= 0 && (int)ZB_W(t) == depth) { used++; t--; }
         while (avail > used) { ZB_SETW(x, depth); x--; avail--; }
         avail = used << 1; depth++; used = 0;
      }
   }
   __syncwarp();
   for (int i = lane; i < n; i += 32) len[key[i] & 511u] = (int)ZB_W(i);
#undef ZB_W
#undef ZB_SETW
   __syncwarp();
}

/* zb_huff_build without codewords (huffencoder.c:279-345): lengths, then the length limit.  The limit is needed only when the
   longest code exceeds maxlen - rare - and then runs as the scalar zb_huff_limit on lane 0. */
__device__ __forceinline__ void zbw_build(const int *cnt, int nsym, int maxlen, int *len, ZbWarpScratch &s, int *ub_hit, const int lane) {
   zbw_lengths(cnt, nsym, len, s, lane);
   int mx = 0;
   for (int i = lane; i < nsym; i += 32) mx = max(mx, len[i]);
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(ZBW_FULL, mx, d));
   if (mx > maxlen) {
      if (lane == 0) { int ub = 0; zb_huff_limit(len, nsym, maxlen, s.order, &ub); if (ub && ub_hit) *ub_hit = 1; }
      __syncwarp();
   }
}

/* zb_dynamic_cost (blockdeflate.c:577-618) */
__device__ __forceinline__ int zbw_dynamic_cost(const int *lcnt, const int *llen, const int *ocnt, const int *olen, ZbWarpScratch &s, const int lane) {
   int cost = 0;
   for (int i = lane; i < 286; i += 32) cost += lcnt[i] * (llen[i] + (i >= 257 ? zb_lensym_extra(i - 257) : 0));
   if (lane < ZB_NOFF) cost += ocnt[lane] * (olen[lane] + zb_offsym_extra(lane));
   cost = zbw_sum(cost);
   /* HLIT / HDIST: last defined length (zb_defined_count) */
   int nl = 257, no = 1;
   for (int b = 256; b < ZB_NLIT; b += 32) { const int i = b + lane; const uint32_t mk = __ballot_sync(ZBW_FULL, i < ZB_NLIT && i >= 257 && llen[i] != 0); if (mk) nl = b + 32 - __clz((int)mk); }
   { const uint32_t mk = __ballot_sync(ZBW_FULL, lane >= 1 && olen[lane] != 0); if (mk) no = 32 - __clz((int)mk); }
   for (int i = lane; i < nl; i += 32) s.sc.cl[i] = (uint8_t)(llen[i] > 255 ? 255 : llen[i]);
   for (int i = lane; i < no; i += 32) s.sc.cl[nl + i] = (uint8_t)(olen[i] > 255 ? 255 : olen[i]);
   __syncwarp();
   int tail = 0;
   if (lane == 0) {      /* the code-length alphabet: 19 symbols, scalar */
      for (int i = 0; i < ZB_NCL; i++) s.sc.clcnt[i] = 0;
      ZbRleCount cv = {s.sc.clcnt};
      zb_rle_scan(s.sc.cl, nl + no, 7, cv);
      zb_huff_lengths(s.sc.clcnt, ZB_NCL, s.sc.cllen, s.sc.key);
      ZbRleSize sv = {s.sc.cllen, 0};
      zb_rle_scan(s.sc.cl, nl + no, 31, sv);
      tail = 5 + 5 + 4 + 3 * zb_raw_table_size(s.sc.cllen) + sv.bits + 3;
   }
   tail = __shfl_sync(ZBW_FULL, tail, 0);
   __syncwarp();
   return cost + tail;
}

/* ZbGreedyView::range_hist: histogram of greedy tokens [t1, t2) of window w into h[ZB_NH] (shared) */
template <class GV>
__device__ __forceinline__ void zbw_add_tokens(const GV &gv, int w, uint32_t t1, uint32_t t2, int *h, const int lane) {
   const uint32_t gb = gv.wbs[w]; const uint8_t *t = gv.T + gv.wd[w].in_off;
   for (uint32_t q = gv.wtb[w] + t1 + (uint32_t)lane; q < gv.wtb[w] + t2; q += 32) {
      const uint32_t p = gv.tp[q];
      const uint32_t l = gv.gl[gb + p];
      if (l >= ZB_MIN_MATCH) { atomicAdd(h + zb_len_sym(l - ZB_MIN_MATCH), 1); atomicAdd(h + ZB_NLIT + zb_off_sym(gv.go[gb + p]), 1); }
      else atomicAdd(h + t[p], 1);
   }
}
template <class GV>
__device__ __forceinline__ void zbw_range_hist(const GV &gv, int w, uint32_t t1, uint32_t t2, int *h, const int lane) {
   const uint32_t k1 = (t1 + ZB_TOKI - 1) / ZB_TOKI, k2 = t2 / ZB_TOKI;
   if (k1 <= k2) {
      const int *a = gv.PH + (size_t)(gv.wib[w] + k1) * ZB_NH, *b = gv.PH + (size_t)(gv.wib[w] + k2) * ZB_NH;
      for (int i = lane; i < ZB_NH; i += 32) h[i] = b[i] - a[i];
      __syncwarp();
      zbw_add_tokens(gv, w, t1, k1 * ZB_TOKI, h, lane);
      zbw_add_tokens(gv, w, k2 * ZB_TOKI, t2, h, lane);
   } else {
      for (int i = lane; i < ZB_NH; i += 32) h[i] = 0;
      __syncwarp();
      zbw_add_tokens(gv, w, t1, t2, h, lane);
   }
   __syncwarp();
}

/* zb_make_costtab over the lanes; zero lengths read as `dl` / `dd` when fill is set (blockdeflate.c:873-881) */
__device__ __forceinline__ void zbw_make_costtab(const int *llen, const int *olen, bool fill, ZbCostTab &t, const int lane) {
   for (int i = lane; i < 256; i += 32) {
      const int a = llen[i];
      t.lit[i] = (uint8_t)((fill && !a) ? 9 : a);
      const int sy = zb_len_sym((uint32_t)i), b = llen[sy];
      t.len[i] = (uint8_t)(((fill && !b) ? 9 : b) + zb_len_extra_bits((uint32_t)i));
   }
   if (lane < 32) { const int o = olen[lane]; t.off[lane] = (uint8_t)(((fill && !o) ? 6 : o) + zb_offsym_extra(lane)); }
}

#endif /* ZB_EMU */
#endif
