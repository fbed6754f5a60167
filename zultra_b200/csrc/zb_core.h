/*
 * zb_core.h - per-task device logic of the zultra-b200 compression path.
 *
 * Everything here is a __host__ __device__ function that one GPU thread runs for one task (one Huffman
 * build, one parse chunk, one match-finder tile ...).  The CUDA kernels in zb_kernels.cu are thin
 * task->thread mappings around these functions.  tests/emu/ compiles the same functions for the host so
 * that the logic can be diffed against the reference without a GPU; the product library never runs them
 * on the CPU.
 *
 * Reference behaviour each function reproduces is cited as file:line of emmanuel-marty/zultra.
 */
#ifndef ZB_CORE_H
#define ZB_CORE_H
#include <stdint.h>

#ifdef __CUDACC__
#define ZB_HD __host__ __device__ __forceinline__
#define ZB_HDN static __host__ __device__ __noinline__
#else
#define ZB_HD inline
#define ZB_HDN inline
#endif

/* format constants (format.h:37-50, private.h:41-56) */
#define ZB_MIN_MATCH 3
#define ZB_MAX_MATCH 258
#define ZB_MAX_OFFSET 32768
#define ZB_HISTORY 32768
#define ZB_NLIT 288
#define ZB_NOFF 32
#define ZB_NCL 19
#define ZB_EOB 256
#define ZB_NMATCH 8
#define ZB_LEAVE_ALONE 40
#define ZB_MAX_SPLITS 64
#define ZB_POS_BITS 22
#define ZB_POS_MASK 0x3fffffu

typedef struct { uint16_t length, offset; } zb_match_t;

ZB_HD int zb_ilog2(uint32_t v) {
#ifdef __CUDA_ARCH__
   return 31 - __clz((int)v);
#else
   return 31 - __builtin_clz(v);
#endif
}

/* ---- deflate symbol arithmetic (RFC 1951 3.2.5; tables at blockdeflate.c:45-85) ---- */

/* lenidx = length-3 clamped to 255 (blockdeflate.c:203-205) -> symbol 257..285 */
ZB_HD int zb_len_sym(uint32_t lenidx) {
   if (lenidx > 255) lenidx = 255;
   if (lenidx < 8) return 257 + (int)lenidx;
   if (lenidx == 255) return 285;
   int e = zb_ilog2(lenidx) - 2;
   return 257 + 4 * (e + 1) + (int)((lenidx >> e) & 3);
}
ZB_HD int zb_len_extra_bits(uint32_t lenidx) {
   if (lenidx > 255) lenidx = 255;
   if (lenidx < 8 || lenidx == 255) return 0;
   return zb_ilog2(lenidx) - 2;
}
/* number of extra bits of length symbol s-257 (g_nRevMatchSymbolBits) */
ZB_HD int zb_lensym_extra(int s) { return (s < 8 || s >= 28) ? 0 : ((s - 4) >> 2); }
/* offset 1..32768 -> distance symbol 0..29 */
ZB_HD int zb_off_sym(uint32_t offset) {
   uint32_t d = offset - 1;
   if (d < 4) return (int)d;
   int b = zb_ilog2(d);
   return 2 * b + (int)((d >> (b - 1)) & 1);
}
ZB_HD int zb_off_extra_bits(uint32_t offset) {
   uint32_t d = offset - 1;
   if (d < 4) return 0;
   return zb_ilog2(d) - 1;
}
ZB_HD int zb_offsym_extra(int s) { return (s < 4 || s >= 30) ? 0 : ((s - 2) >> 1); }

/* ---- Huffman code construction (huffencoder.c) ---- */

/* Sort keys ascending, n <= 288.  The keys are distinct and arrive ordered by their low 9 bits (the symbol index), so a
   stable LSD radix sort of the bits above them (the count) is a full sort: 1..3 passes of <= 7 bits over the bits the largest
   key uses, instead of a comparison sort's long chain of dependent local-memory moves (these builds run one per thread, and
   their latency is what the splitter and the table stages wait for).  Small inputs: insertion sort. */
ZB_HD void zb_sort_u32(uint32_t *a, int n) {
   if (n <= 24) {
      for (int i = 1; i < n; i++) {
         uint32_t v = a[i];
         int j = i;
         while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; j--; }
         a[j] = v;
      }
      return;
   }
   uint32_t tmp[ZB_NLIT];
   uint16_t bin[128];
   uint32_t top = 0;
   for (int i = 0; i < n; i++) top |= a[i];
   int bits = 0;
   while ((top >> 9) >> bits) bits++;
   if (bits == 0) return;
   const int passes = (bits + 6) / 7, w = (bits + passes - 1) / passes, nb = 1 << w;
   uint32_t *src = a, *dst = tmp;
   for (int p = 0; p < passes; p++) {
      const int sh = 9 + p * w;
      for (int i = 0; i < nb; i++) bin[i] = 0;
      for (int i = 0; i < n; i++) bin[(src[i] >> sh) & (uint32_t)(nb - 1)]++;
      uint32_t run = 0;
      for (int i = 0; i < nb; i++) { const uint32_t c = bin[i]; bin[i] = (uint16_t)run; run += c; }
      for (int i = 0; i < n; i++) { const uint32_t v = src[i]; dst[bin[(v >> sh) & (uint32_t)(nb - 1)]++] = v; }
      uint32_t *t = src; src = dst; dst = t;
   }
   if (src != a) for (int i = 0; i < n; i++) a[i] = src[i];
}

/*
 * Minimum-redundancy code lengths, unlimited (huffencoder.c:157-270: symbols with a non-zero count sorted by
 * (count, index) ascending, Moffat-Katajainen in place; zero or one used symbol -> len[0] = 1 only).
 * cnt[0..nsym) -> len[0..288).  scratch: key[288].
 */
ZB_HD void zb_huff_lengths(const int *cnt, int nsym, int *len, uint32_t *key) {
   int n = 0;
   for (int i = 0; i < ZB_NLIT; i++) len[i] = 0;
   for (int i0 = 0; i0 < nsym; i0 += 8) {      /* the counts may sit in global memory: 8 loads in flight, then the compaction */
      int c[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 8; j++) c[j] = i0 + j < nsym ? cnt[i0 + j] : 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 8; j++) if (c[j]) key[n++] = ((uint32_t)c[j] << 9) | (uint32_t)(i0 + j);
   }
   if (n <= 1) { len[0] = 1; return; }
   zb_sort_u32(key, n);
   /* w[] lives in the upper bits of key[]; keep symbol ids aside in the low 9 bits */
   /* phase 1: pair up the two lightest nodes; w[t] becomes the weight of internal node t, and consumed
      internal nodes get their parent index */
   int leaf = 0, inode = 0;
#define ZB_W(i) (key[i] >> 9)
#define ZB_SETW(i, v) (key[i] = ((uint32_t)(v) << 9) | (key[i] & 511u))
   /* weights can reach 2*sum; sum <= 2^21+1 fits in 23 bits */
   for (int t = 0; t < n - 1; t++) {
      uint32_t w;
      if (leaf >= n || (inode < t && ZB_W(inode) < ZB_W(leaf))) { w = ZB_W(inode); ZB_SETW(inode, t); inode++; }
      else { w = ZB_W(leaf); leaf++; }
      if (leaf >= n || (inode < t && ZB_W(inode) < ZB_W(leaf))) { w += ZB_W(inode); ZB_SETW(inode, t); inode++; }
      else { w += ZB_W(leaf); leaf++; }
      ZB_SETW(t, w);
   }
   /* phase 2: internal node depths */
   ZB_SETW(n - 2, 0);
   for (int t = n - 3; t >= 0; t--) ZB_SETW(t, ZB_W(ZB_W(t)) + 1);
   /* phase 3: leaf depths */
   int avail = 1, used = 0, depth = 0, t = n - 2, x = n - 1;
   while (avail > 0) {
      while (t >= 0 && (int)ZB_W(t) == depth) { used++; t--; }
      while (avail > used) { ZB_SETW(x, depth); x--; avail--; }
      avail = used << 1; depth++; used = 0;
   }
   for (int i = 0; i < n; i++) len[key[i] & 511u] = (int)ZB_W(i);
#undef ZB_W
#undef ZB_SETW
}

/* order[] = symbols with len != 0 sorted by (len, index) ascending; returns count.  Lengths <= 287.  Counting sort (one
   pass to count, one to place): limited codes have <= 15 distinct lengths, the per-length rescans of a naive selection were a
   fifth of a table build. */
ZB_HD int zb_order_by_len(const int *len, int nsym, int16_t *order) {
   int maxl = 0;
   for (int i = 0; i < nsym; i++) if (len[i] > maxl) maxl = len[i];
   if (maxl <= 31) {
      uint16_t start[32];
      for (int l = 0; l <= maxl; l++) start[l] = 0;
      for (int i = 0; i < nsym; i++) start[len[i]]++;
      int run = 0;
      for (int l = 1; l <= maxl; l++) { const int c = start[l]; start[l] = (uint16_t)run; run += c; }
      for (int i = 0; i < nsym; i++) { const int l = len[i]; if (l) order[start[l]++] = (int16_t)i; }
      return run;
   }
   int n = 0;
   for (int l = 1; l <= maxl; l++)
      for (int i = 0; i < nsym; i++)
         if (len[i] == l) order[n++] = (int16_t)i;
   return n;
}

/*
 * Length limiting exactly as huffencoder.c:303-345 (clamp, lengthen from the long end until Kraft fits,
 * shorten from the short end while it still fits).  The reference's last loop runs one index past the
 * sorted list (i <= nNumSorted, huffencoder.c:334) which is undefined behaviour when Kraft slack remains
 * after the last symbol; we stop at the last symbol and report that case through *ub_hit.
 * Returns number of coded symbols; order[] is the final (len,index) order.
 */
ZB_HD int zb_huff_limit(int *len, int nsym, int maxlen, int16_t *order, int *ub_hit) {
   int n = zb_order_by_len(len, nsym, order);
   if (n > 0 && maxlen > 0 && len[order[n - 1]] > maxlen) {
      int k = 0, maxk = 1 << maxlen;
      for (int i = n - 1; i >= 0; i--) {
         int s = order[i];
         if (len[s] > maxlen) len[s] = maxlen;
         k += maxk >> len[s];
      }
      for (int i = n - 1; k > maxk && i >= 0; i--) {
         int s = order[i];
         while (len[s] < maxlen && k > maxk) { len[s]++; k -= maxk >> len[s]; }
      }
      int i = 0;
      for (; k < maxk && i < n; i++) {
         int s = order[i];
         while (k + (maxk >> len[s]) <= maxk) { k += maxk >> len[s]; len[s]--; }
      }
      if (k < maxk && i == n && ub_hit) *ub_hit = 1;
      n = zb_order_by_len(len, nsym, order);
   }
   return n;
}

ZB_HD uint32_t zb_rev16(uint32_t v) {
   v = ((v & 0x5555u) << 1) | ((v & 0xaaaau) >> 1);
   v = ((v & 0x3333u) << 2) | ((v & 0xccccu) >> 2);
   v = ((v & 0x0f0fu) << 4) | ((v & 0xf0f0u) >> 4);
   v = ((v & 0x00ffu) << 8) | ((v & 0xff00u) >> 8);
   return v;
}

/* canonical, bit-reversed codewords from (len,index)-ordered symbols (huffencoder.c:350-372, :120-145) */
ZB_HD void zb_huff_codes(const int *len, const int16_t *order, int n, uint16_t *code) {
   if (n <= 0) return;
   uint32_t c = 0;
   int cl = len[order[0]];
   for (int i = 0; i < n; i++) {
      int s = order[i];
      code[s] = (uint16_t)(zb_rev16(c) >> (16 - cl));
      if (i + 1 < n) { int nl = len[order[i + 1]]; c = (c + 1) << (nl - cl); cl = nl; }
   }
}

/* build_dynamic_codewords (huffencoder.c:279-375): lengths + limit (+ codes if code != 0). key: 288 u32, order: 288 i16 */
ZB_HD void zb_huff_build(const int *cnt, int nsym, int maxlen, int *len, uint16_t *code, uint32_t *key, int16_t *order, int *ub_hit) {
   zb_huff_lengths(cnt, nsym, len, key);
   int n = zb_huff_limit(len, nsym, maxlen, order, ub_hit);
   if (code) zb_huff_codes(len, order, n, code);
}

/* HLIT / HDIST style trailing-zero trim (huffencoder.c:532-538) */
ZB_HD int zb_defined_count(const int *len, int nsym, int minsym) {
   int i = nsym;
   while (i > minsym && !len[i - 1]) i--;
   return i;
}

/* ---- run-length coding of the code-length sequence (huffencoder.c:446-735) ---- */

/* visitor interface: v.sym(s) for a plain code-length symbol, v.rep(s, extra_value, extra_bits) for 16/17/18 */
template <class V>
ZB_HD void zb_rle_scan(const uint8_t *cl, int n, unsigned mask, V &v) {
   int i = 0;
   while (i < n) {
      int run = 1;
      while (i + run < n && cl[i + run] == cl[i]) run++;
      /* What the repeat codes leave of a run is re-scanned by the reference one symbol per round (huffencoder.c:450-453), and no
         repeat code can apply to it any more (the loops below only stop once the remainder is too short for the codes the mask
         allows): it comes out as plain symbols.  They are emitted here at once - same tokens, but linear instead of quadratic in
         the run length when the mask disables a repeat code (a 128-long zero run under mask 0 cost 8 000 steps per scan). */
      if (cl[i] == 0) {
         if (run >= 3) {
            while (run >= 11 && (mask & 4)) { int m = run > 138 ? 138 : run; v.rep(18, m - 11, 7); run -= m; i += m; }
            while (run >= 3 && (mask & 2)) { int m = run > 10 ? 10 : run; v.rep(17, m - 3, 3); run -= m; i += m; }
         }
         for (; run > 0; run--) { v.sym(0); i++; }
      } else {
         int c = cl[i] > 15 ? 15 : cl[i];
         run--; v.sym(c); i++;
         if (run == 7 && (mask & 1) && !(mask & 8)) { v.rep(16, 1, 2); v.rep(16, 0, 2); run -= 7; i += 7; }
         else if (run == 8 && (mask & 1) && !(mask & 16)) { v.rep(16, 1, 2); v.rep(16, 1, 2); run -= 8; i += 8; }
         while (run >= 3 && (mask & 1)) { int m = run > 6 ? 6 : run; v.rep(16, m - 3, 2); run -= m; i += m; }
         for (; run > 0; run--) { v.sym(c); i++; }
      }
   }
}
struct ZbRleCount { int *cnt; ZB_HD void sym(int s) { cnt[s]++; } ZB_HD void rep(int s, int, int) { cnt[s]++; } };
struct ZbRleSize { const int *len; int bits; ZB_HD void sym(int s) { bits += len[s]; } ZB_HD void rep(int s, int, int eb) { bits += len[s] + eb; } };

/* RFC 1951 3.2.7 transmission order of the code-length alphabet (huffencoder.c:30) */
ZB_HD int zb_clorder(int i) {
   const uint8_t o[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
   return o[i];
}
/* HCLEN+4 (huffencoder.c:400-406) */
ZB_HD int zb_raw_table_size(const int *cllen) {
   int i = ZB_NCL;
   while (i > 4 && !cllen[zb_clorder(i - 1)]) i--;
   return i;
}

/* scratch shared by the cost / build helpers of one thread */
struct ZbScratch {
   uint32_t key[ZB_NLIT];
   int16_t order[ZB_NLIT];
   uint8_t cl[ZB_NLIT + ZB_NOFF];
   int clcnt[ZB_NCL];
   int cllen[ZB_NLIT]; /* zb_huff_lengths clears 288 entries */
};

/*
 * zultra_block_evaluate_dynamic_cost (blockdeflate.c:577-618): data bits with the given lengths, plus
 * 14 header bits, 3 x HCLEN, the RLE-coded lengths sized with mask 31 under an UNLIMITED code for the
 * code-length alphabet whose histogram was taken with mask 7, plus 3.
 */
ZB_HD int zb_dynamic_cost(const int *lcnt, const int *llen, const int *ocnt, const int *olen, ZbScratch &s) {
   int cost = 0;
   for (int i = 0; i < 257; i++) cost += lcnt[i] * llen[i];
   for (int i = 257; i < 286; i++) cost += lcnt[i] * (llen[i] + zb_lensym_extra(i - 257));
   for (int i = 0; i < ZB_NOFF; i++) cost += ocnt[i] * (olen[i] + zb_offsym_extra(i));
   int nl = zb_defined_count(llen, ZB_NLIT, 257), no = zb_defined_count(olen, ZB_NOFF, 1);
   /* lengths above 255 cannot occur (<= 287 symbols, but 288 > 255): estimate lengths are < 288; keep them
      in a byte only when they fit, else clamp to 255 which still compares unequal to every real neighbour
      only if that neighbour is not also >= 255 - depths that large need > 2^255 total count, impossible */
   for (int i = 0; i < nl; i++) s.cl[i] = (uint8_t)(llen[i] > 255 ? 255 : llen[i]);
   for (int i = 0; i < no; i++) s.cl[nl + i] = (uint8_t)(olen[i] > 255 ? 255 : olen[i]);
   for (int i = 0; i < ZB_NCL; i++) s.clcnt[i] = 0;
   ZbRleCount cv = {s.clcnt};
   zb_rle_scan(s.cl, nl + no, 7, cv);
   zb_huff_lengths(s.clcnt, ZB_NCL, s.cllen, s.key);
   cost += 5 + 5 + 4;
   cost += 3 * zb_raw_table_size(s.cllen);
   ZbRleSize sv = {s.cllen, 0};
   zb_rle_scan(s.cl, nl + no, 31, sv);
   return cost + sv.bits + 3;
}

/* zultra_block_evaluate_static_cost (blockdeflate.c:538-566) */
ZB_HD int zb_static_lit_len(int i) { return i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8)); }
ZB_HD int zb_static_cost(const int *lcnt, const int *ocnt) {
   int cost = 0;
   for (int i = 0; i < 257; i++) cost += lcnt[i] * zb_static_lit_len(i);
   for (int i = 257; i < 286; i++) cost += lcnt[i] * (zb_static_lit_len(i) + zb_lensym_extra(i - 257));
   for (int i = 0; i < ZB_NOFF; i++) cost += ocnt[i] * (5 + zb_offsym_extra(i));
   return cost + 3;
}

/* zultra_huffman_encoder_optimize_for_rle (huffutils.c:34-114, from Zopfli): smooth counts so that the
   code lengths RLE-compress better.  good: scratch of >= length bytes. */
ZB_HD void zb_smooth_counts(int length, int *counts, uint8_t *good) {
   while (length > 0 && counts[length - 1] == 0) length--;
   if (length == 0) return;
   for (int i = 0; i < length; i++) good[i] = 0;
   /* protect existing long runs: >= 5 zeros, >= 7 equal non-zeros */
   {
      int cur = counts[0], stride = 0;
      for (int i = 0; i <= length; i++) {
         if (i == length || counts[i] != cur) {
            if ((cur == 0 && stride >= 5) || (cur != 0 && stride >= 7))
               for (int k = 0; k < stride; k++) good[i - k - 1] = 1;
            stride = 1;
            if (i != length) cur = counts[i];
         } else stride++;
      }
   }
   /* collapse strides of near-equal counts to their rounded mean */
   {
      int stride = 0;
      long long limit = counts[0], sum = 0;
      for (int i = 0; i <= length; i++) {
         long long d = 0;
         if (i != length) { d = (long long)counts[i] - limit; if (d < 0) d = -d; }
         if (i == length || good[i] || d >= 4) {
            if (stride >= 4 || (stride >= 3 && sum == 0)) {
               int c = (int)((sum + stride / 2) / stride);
               if (c < 1) c = 1;
               if (sum == 0) c = 0;
               for (int k = 0; k < stride; k++) counts[i - k - 1] = c;
            }
            stride = 0; sum = 0;
            if (i < length - 3) limit = ((long long)counts[i] + counts[i + 1] + counts[i + 2] + counts[i + 3] + 2) / 4;
            else if (i < length) limit = counts[i];
            else limit = 0;
         }
         ++stride;
         if (i != length) sum += counts[i];
      }
   }
}

/* ---- LSB-first bit writer into a zero-initialised 32-bit word buffer (bitwriter.c:63-98 semantics) ----
 * Words that may be shared with a neighbouring writer (first/last of a span) are merged with atomicOr on the
 * device; on the host plain OR.
 */
struct ZbBitSink {
   uint32_t *words; /* base of the output, word 0 = bits 0..31 */
   uint64_t acc;    /* pending bits */
   int nacc;        /* number of pending bits (< 32 after flush) */
   uint64_t wpos;   /* index of the word the pending bits start in */
   uint64_t first_word;
   ZB_HD void init(uint32_t *w, uint64_t bitpos) {
      words = w; wpos = bitpos >> 5; first_word = wpos; nacc = (int)(bitpos & 31); acc = 0;
   }
   ZB_HD void orword(uint64_t idx, uint32_t v, bool shared) {
      if (!v) return;
#ifdef __CUDA_ARCH__
      if (shared) atomicOr(words + idx, v); else words[idx] = v;
#else
      (void)shared; words[idx] |= v;
#endif
   }
   ZB_HD void put(uint32_t value, int nbits) {
      acc |= (uint64_t)value << nacc;
      nacc += nbits;
      if (nacc >= 32) {
         orword(wpos, (uint32_t)acc, wpos == first_word);
         acc >>= 32; nacc -= 32; wpos++;
      }
   }
   ZB_HD void finish() { if (nacc > 0) orword(wpos, (uint32_t)acc, true); }
};

/* ---- match finder ---- */

/*
 * Match list of one main position by direct scan of the tile's suffix array - no interval tree, no mutation, so every
 * position of a tile is independent (one GPU thread each, the tile's list held in shared memory).
 *
 * words[0..n): the tile's suffixes in suffix-array order, pos | lcp<<22 (lcp with the previous entry, already clamped to
 * {0,3..258}).  r = index of the suffix starting at tile position i.  Walk outwards from r, always taking the side whose
 * running LCP with i is larger, so candidates arrive by non-increasing LCP.  best = nearest earlier position (within
 * 32768) seen so far; whenever the LCP level drops and best moved during the level just finished, (level, i - best) is
 * the next entry of the reference's list: exactly the positions that zultra_find_matches_at (matchfinder.c:171-234)
 * reports, longest first - for each length the nearest earlier occurrence, each position once, with its own length
 * (SURVEY 8(a)-M1).  The reference stops storing after 8 entries (matchfinder.c:217); nothing nearer than offset 1 exists.
 *
 * Rank walk + text walk.  The rank walk costs one step per member of the enclosing LCP >= 3 group, which is heavy-tailed
 * (common trigrams, byte runs).  Every suffix the walk has seen lies outside the position interval (best, i) - otherwise
 * it would have become `best` - so what the remaining walk could still add is exactly the Pareto frontier of the positions
 * j in (best, i), all with lcp(i, j) <= bound = the LCP level the walk has reached.  Once that interval is short compared
 * with the steps already spent, it is cheaper to read it directly: j = i-1 .. best+1, byte-compare up to `bound`, keep
 * every j that is strictly longer than all nearer ones, stop as soon as one reaches `bound`.  Same list, and the cost of
 * a position becomes O(min(group size, distance to the nearest long match)).
 */
#ifndef ZB_TS_MIN
#define ZB_TS_MIN 12      /* rank-walk steps before the text walk is considered */
#define ZB_TS_MUL 6       /* switch when (i - best - 1) <= ZB_TS_MUL * steps */
#endif
#ifdef ZB_SCAN_STATS
extern long g_zb_rank_steps, g_zb_text_steps, g_zb_cmp_bytes;
#define ZB_STAT(x) (x)
#else
#define ZB_STAT(x) ((void)0)
#endif

/* T[p] = byte at tile position p (readable up to the window end) */
ZB_HD int zb_mf_scan(const uint32_t *words, int n, int r, uint32_t i, const uint8_t *T, zb_match_t *out) {
   int L = r - 1, R = r + 1;
   uint32_t lL = r > 0 ? (words[r] >> ZB_POS_BITS) : 0u;
   uint32_t lR = R < n ? (words[R] >> ZB_POS_BITS) : 0u;
   int best = (i > ZB_MAX_OFFSET ? (int)(i - ZB_MAX_OFFSET) : 0) - 1;   /* p > best also enforces the 32768 limit */
   int nm = 0, steps = 0;
   uint32_t lvl = 0;
   bool moved = false;
   for (;;) {
      const uint32_t l = lL > lR ? lL : lR;
      if (l < lvl && moved) {
         out[nm].length = (uint16_t)lvl; out[nm].offset = (uint16_t)(i - (uint32_t)best); nm++;
         moved = false;
         if (nm == ZB_NMATCH) return nm;
      }
      if (l < ZB_MIN_MATCH) return nm;
      if (steps >= ZB_TS_MIN && (int)i - 1 - best <= ZB_TS_MUL * steps) {
         /* text walk over (best, i): records arrive nearest first = shortest first; tr[0] is the longest so far */
         uint32_t tr[ZB_NMATCH]; int nt = 0; uint32_t curmax = 0;
         const uint8_t c0 = T[i], c1 = T[i + 1], c2 = T[i + 2];
         for (int j = (int)i - 1; j > best; j--) {
            ZB_STAT(g_zb_text_steps++);
            if (T[j] != c0 || T[j + 1] != c1 || T[j + 2] != c2) continue;
            uint32_t len = ZB_MIN_MATCH;
            while (len < l && T[j + len] == T[i + len]) len++;
            ZB_STAT(g_zb_cmp_bytes += len);
            if (len > curmax) {
               for (int z = ZB_NMATCH - 1; z > 0; z--) tr[z] = tr[z - 1];
               tr[0] = len | ((i - (uint32_t)j) << 16);
               if (nt < ZB_NMATCH) nt++;
               curmax = len;
               if (len == l) break;
            }
         }
         /* the pending record of the rank walk stands unless a nearer position reached the same level */
         if (moved && !(nt && curmax == lvl)) { out[nm].length = (uint16_t)lvl; out[nm].offset = (uint16_t)(i - (uint32_t)best); nm++; }
         for (int z = 0; z < nt && nm < ZB_NMATCH; z++) { out[nm].length = (uint16_t)(tr[z] & 0xffffu); out[nm].offset = (uint16_t)(tr[z] >> 16); nm++; }
         return nm;
      }
      lvl = l; steps++;
      ZB_STAT(g_zb_rank_steps++);
      int p;
      if (lL > lR || (lL == lR && (steps & 1))) {      /* ties alternate: see zb_mf_scan_k */
         const uint32_t w = words[L];
         p = (int)(w & ZB_POS_MASK);
         const uint32_t wl = w >> ZB_POS_BITS;
         lL = L > 0 ? (wl < lL ? wl : lL) : 0u;
         L--;
      } else {
         p = (int)(words[R] & ZB_POS_MASK);
         R++;
         if (R < n) { const uint32_t wl = words[R] >> ZB_POS_BITS; lR = wl < lR ? wl : lR; } else lR = 0u;
      }
      if (p < (int)i && p > best) {
         best = p; moved = true;
         if (best == (int)i - 1) {   /* nothing nearer exists and the level of this record is fixed: done */
            out[nm].length = (uint16_t)lvl; out[nm].offset = 1; nm++;
            return nm;
         }
      }
   }
}

/* ---- optimal parse (blockdeflate.c:254-323) ---- */

/* per sub-block bit costs fed to the parse: lit[b], len[length-3] (symbol + extra bits), off[dist symbol] (+extra) */
struct ZbCostTab { uint8_t lit[256]; uint8_t len[256]; uint8_t off[32]; };

ZB_HD void zb_make_costtab(const int *llen, const int *olen, ZbCostTab &t) {
   for (int i = 0; i < 256; i++) t.lit[i] = (uint8_t)llen[i];
   for (int i = 0; i < 256; i++) t.len[i] = (uint8_t)(llen[zb_len_sym((uint32_t)i)] + zb_len_extra_bits((uint32_t)i));
   for (int i = 0; i < 32; i++) t.off[i] = (uint8_t)(olen[i] + zb_offsym_extra(i));
}

#define ZB_RING 260 /* cost[i+1 .. i+258] plus cost[i] */

/*
 * Backward cost recurrence over positions [from-1 .. lo] of one sub-block ending at `end` (exclusive).
 * Costs are kept modulo 2^16 in a ring; every comparison is made on differences, which stay far below 2^15
 * inside the 258-position horizon (<= 258*15 + 48), so the choices equal the reference's int arithmetic.
 * ring slot of position p is (slot_from + (from - p)) mod ZB_RING going DOWN in p, i.e. we keep `s` = slot of i.
 * Ring must hold valid costs for positions [from, min(from+258,end)] on entry (slot of `from` = slot0).
 * Writes best[i] for i in [lo, keep_hi).  If sig != 0, stores cost[lo+k]-cost[lo] for k=0..258 (0 beyond end).
 * RING: accessor with get(slot)/set(slot,v).
 */
struct ZbMatchRec { uint32_t w[8]; };   /* the 8 candidates of one position: length | offset << 16 */
ZB_HD ZbMatchRec zb_load_rec(const zb_match_t *match, int i) {
   ZbMatchRec r;
#ifdef __CUDA_ARCH__
   const uint4 *q = (const uint4 *)(match + ((size_t)i << 3));
   const uint4 a = __ldg(q), b = __ldg(q + 1);
   r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w; r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
#else
   const zb_match_t *pm = match + ((size_t)i << 3);
   for (int m = 0; m < 8; m++) r.w[m] = (uint32_t)pm[m].length | ((uint32_t)pm[m].offset << 16);
#endif
   return r;
}

template <class RING>
ZB_HD void zb_parse_range(const uint8_t *T, const zb_match_t *match, const ZbCostTab &tab, int lo, int from, int end,
                          int keep_hi, zb_match_t *best, RING &ring, int &slot /* in: slot of `from`; out: slot of lo */) {
   int s = slot;
   if (from - 1 < lo) return;
   ZbMatchRec nxt = zb_load_rec(match, from - 1);
   uint32_t nlit = T[from - 1];
   for (int i = from - 1; i >= lo; i--) {
      const ZbMatchRec rec = nxt;
      const uint32_t lit = nlit;
      if (i - 1 >= lo) { nxt = zb_load_rec(match, i - 1); nlit = T[i - 1]; }   /* independent of the recurrence: overlaps with it */
      int s1 = s;                 /* slot of i+1 */
      s = s1 + 1; if (s >= ZB_RING) s -= ZB_RING; /* slot of i: going down in position = going up in slot */
      const uint16_t base = ring.get(s1);
      int bestc = tab.lit[lit];
      int bestlen = 0, bestoff = 0;
      /* Candidates in the reference's order are: literal; then for m = 0..7, lengths ml..3 (only ml for a >= 40 match),
         replaced on strictly lower cost (blockdeflate.c:272-312).  lencost(k) + cost[i+k] does not depend on the match,
         so the best length of a short match is a prefix minimum over k (largest k on ties = first met going down), and
         the matches are combined shortest first with "<=" so that the earlier (longer) match keeps ties.  Same choice,
         at most 37 + 8 evaluations instead of 8 x 37. */
      int M = 0;
#pragma unroll
      for (int m = 0; m < ZB_NMATCH; m++) if (M == m && (rec.w[m] & 0xffffu) >= ZB_MIN_MATCH) M = m + 1;
      if (M) {
         int bt = 0x7fffffff, bl = 0, bo = 0;
         int k = ZB_MIN_MATCH, curmin = 0x7fffffff, curk = 0;
         int sk = s1 - (ZB_MIN_MATCH - 1); if (sk < 0) sk += ZB_RING;   /* slot of i+3 */
#pragma unroll
         for (int m = ZB_NMATCH - 1; m >= 0; m--) {
            if (m < M) {
               const int mlen0 = (int)(rec.w[m] & 0xffffu), moff = (int)(rec.w[m] >> 16);
               const int offc = tab.off[zb_off_sym((uint32_t)moff)];
               int ml = mlen0;
               if (i + ml > end) ml = end - i;
               int total, kk;
               if (mlen0 >= ZB_LEAVE_ALONE) {
                  int sl = s1 - (ml - 1); if (sl < 0) sl += ZB_RING;
                  int lidx = ml - ZB_MIN_MATCH; if (lidx < 0 || lidx > 255) lidx = 255;
                  total = tab.len[lidx] + offc + (int16_t)(uint16_t)(ring.get(sl) - base);
                  kk = ml;
               } else {
                  while (k <= ml) {
                     const int c = tab.len[k - ZB_MIN_MATCH] + (int16_t)(uint16_t)(ring.get(sk) - base);
                     if (c <= curmin) { curmin = c; curk = k; }
                     k++; sk--; if (sk < 0) sk += ZB_RING;
                  }
                  total = curk ? curmin + offc : 0x7fffffff;
                  kk = curk;
               }
               if (total <= bt && total != 0x7fffffff) { bt = total; bl = kk; bo = moff; }
            }
         }
         if (bt < bestc) { bestc = bt; bestlen = bl; bestoff = bo; }
      }
      ring.set(s, (uint16_t)(base + (uint16_t)bestc));
      if (i < keep_hi) { best[i].length = (uint16_t)bestlen; best[i].offset = (uint16_t)bestoff; }
   }
   slot = s;
}

struct ZbRingLocal {
   uint16_t v[ZB_RING];
   ZB_HD uint16_t get(int s) const { return v[s]; }
   ZB_HD void set(int s, uint16_t x) { v[s] = x; }
};
/* ring of one thread inside a shared-memory array laid out [slot][thread] */
struct ZbRingStrided {
   uint16_t *base; int stride;
   ZB_HD uint16_t get(int s) const { return base[s * stride]; }
   ZB_HD void set(int s, uint16_t x) { base[s * stride] = x; }
};

/* ------------------------------------------------------------------------------------------------------------------------
 * Compact candidate records for the parse kernel (zb_parse_dp_k).  The match list of a position does not change over the four
 * parse passes (blockdeflate.c:871-920 only changes the code lengths), so everything about it that the recurrence needs is
 * worked out ONCE per batch and packed into 16 bytes - half the traffic of the 32-byte match record in each of the four
 * passes, and no offset -> symbol arithmetic, clamping or validity tests left on the recurrence's serial chain.
 *
 * A record is up to eight 16-bit entries in PROCESSING order, terminated by a zero entry:
 *   first the leave-alone matches (original length >= 40: tried at their full clamped length only, blockdeflate.c:286),
 *   in list order (m = 0, 1, ..):        1 . | ml : 9 (bits 13..5, clamped to the sub-block end, may be 1 or 2: SURVEY A-3) | sym : 5
 *   then the short matches, SHORTEST first (descending m), those clamped below 3 dropped:
 *                                          0 1 | m : 3 (bits 13..11) | ml : 6 (bits 10..5) | sym : 5
 * sym = distance symbol of the match's offset.  Candidate order of the reference (literal; m = 0.. longest first; lengths
 * descending; strict '<' to replace) = minimal cost, ties to the smallest m, then to the largest length: leave-alone matches
 * (the smallest m's) are compared in list order with '<', the short ones shortest first with '<=' over a shared prefix
 * minimum of lencost(k) + cost[i + k], and a short match beats the best leave-alone match only when strictly cheaper.
 *
 * The kernel's choice word: k | sym << 9 | m << 14 (0 = literal); zb_choice_k turns the final pass's choices into {length,
 * offset} by fetching the one offset from the match list.
 */
struct ZbCand { uint32_t w[4]; };

ZB_HD ZbCand zb_cand_pack(const ZbMatchRec &rec, int rem /* sub-block end - position */) {
   uint64_t lo = 0, hi = 0;
   int n = 0, M = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
   for (int m = 0; m < ZB_NMATCH; m++) if (M == m && (rec.w[m] & 0xffffu) >= ZB_MIN_MATCH) M = m + 1;
#define ZB_CAND_PUSH(e_) do { const uint64_t e__ = (uint64_t)(e_); if (n < 4) lo |= e__ << (16 * n); else hi |= e__ << (16 * (n - 4)); n++; } while (0)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
   for (int m = 0; m < ZB_NMATCH; m++) {
      const int len0 = (int)(rec.w[m] & 0xffffu);
      if (m < M && len0 >= ZB_LEAVE_ALONE) {
         const int ml = len0 < rem ? len0 : rem;
         ZB_CAND_PUSH(0x8000u | ((uint32_t)ml << 5) | (uint32_t)zb_off_sym(rec.w[m] >> 16));
      }
   }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
   for (int m = ZB_NMATCH - 1; m >= 0; m--) {
      const int len0 = (int)(rec.w[m] & 0xffffu);
      if (m < M && len0 < ZB_LEAVE_ALONE) {
         const int ml = len0 < rem ? len0 : rem;
         if (ml >= ZB_MIN_MATCH) ZB_CAND_PUSH(0x4000u | ((uint32_t)m << 11) | ((uint32_t)ml << 5) | (uint32_t)zb_off_sym(rec.w[m] >> 16));
      }
   }
#undef ZB_CAND_PUSH
   ZbCand c;
   c.w[0] = (uint32_t)lo; c.w[1] = (uint32_t)(lo >> 32); c.w[2] = (uint32_t)hi; c.w[3] = (uint32_t)(hi >> 32);
   return c;
}

/* One position of the recurrence from its compact record, written for clarity: the host build runs it (ZB_EMU_DP_LEAN) to pin
   the record format and the tie-breaking against the golden vectors; the CUDA kernel implements the same arithmetic on its
   shared-memory ring.  cost(k) = cost of position i + k (0 above the chunk's start).  Returns cost[i]; *choice as above. */
template <class COSTAT>
ZB_HD uint32_t zb_dp_eval(const ZbCand &c, uint32_t lit_cost, uint32_t cprev, const ZbCostTab &tab, const COSTAT &cost, uint32_t *choice) {
   uint32_t btla = 0xffffffffu, bwla = 0, bts = 0xffffffffu, bws = 0, curmin = 0xffffffffu;
   int k = ZB_MIN_MATCH, mla = 0;
   for (int j = 0; j < 8; j++) {
      const uint32_t e = (c.w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
      if (!e) break;
      const uint32_t sym = e & 31u;
      if (e & 0x8000u) {
         const int ml = (int)((e >> 5) & 511u);
         const int lidx = ml >= ZB_MIN_MATCH ? ml - ZB_MIN_MATCH : 255;
         const uint32_t total = (uint32_t)tab.len[lidx] + tab.off[sym] + cost(ml);
         if (total < btla) { btla = total; bwla = (uint32_t)ml | (sym << 9) | ((uint32_t)mla << 14); }
         mla++;
      } else {
         const int ml = (int)((e >> 5) & 63u);
         for (; k <= ml; k++) {
            const uint32_t key = ((cost(k) + tab.len[k - ZB_MIN_MATCH]) << 6) | (uint32_t)(63 - k);
            if (key < curmin) curmin = key;
         }
         const uint32_t total = (curmin >> 6) + tab.off[sym];
         if (total <= bts) { bts = total; bws = (63u - (curmin & 63u)) | (sym << 9) | (((e >> 11) & 7u) << 14); }
      }
   }
   uint32_t bt = btla, bw = bwla;
   if (bts < btla) { bt = bts; bw = bws; }
   uint32_t bestc = cprev + lit_cost;
   *choice = 0;
   if (bt < bestc) { bestc = bt; *choice = bw; }
   return bestc;
}

#endif /* ZB_CORE_H */
