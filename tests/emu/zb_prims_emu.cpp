/* Host stand-ins for the cooperative GPU primitives of zb_prims.cu.  TEST BUILD ONLY (-DZB_EMU): lets the
 * per-task pipeline logic run in a host loop so it can be diffed against the reference without a GPU. */
#include "zb_rt.h"
#include <vector>
#include <algorithm>
#include <numeric>

size_t zb_sort_scratch_words(long) { return 64; }
size_t zb_scan_scratch_words(long) { return 64; }

void zb_sort_pairs(zb_stream_t, uint64_t *keys, uint32_t *vals, uint64_t *, uint32_t *, long n, int lo, int hi, uint32_t *) {
   if (n <= 1 || hi <= lo) return;
   const uint64_t mask = (hi - lo >= 64) ? ~0ull : (((1ull << (hi - lo)) - 1) << lo);
   std::vector<long> idx(n);
   std::iota(idx.begin(), idx.end(), 0L);
   std::stable_sort(idx.begin(), idx.end(), [&](long a, long b) { return (keys[a] & mask) < (keys[b] & mask); });
   std::vector<uint64_t> k(n); std::vector<uint32_t> v(n);
   for (long i = 0; i < n; i++) { k[i] = keys[idx[i]]; v[i] = vals[idx[i]]; }
   std::copy(k.begin(), k.end(), keys); std::copy(v.begin(), v.end(), vals);
}
void zb_exclusive_sum(zb_stream_t, const uint32_t *in, uint32_t *out, long n, uint32_t *total, uint32_t *) {
   uint32_t acc = 0;
   for (long i = 0; i < n; i++) { uint32_t v = in[i]; out[i] = acc; acc += v; }
   if (total) *total = acc;
}
void zb_inclusive_max(zb_stream_t, const uint32_t *in, uint32_t *out, long n, uint32_t *) {
   uint32_t acc = 0;
   for (long i = 0; i < n; i++) { acc = std::max(acc, in[i]); out[i] = acc; }
}
void zb_tile_filter(zb_stream_t, const uint32_t *srcw, const uint32_t *src_cnt, const ZbTileDesc *tiles, int ntiles, int first, uint32_t *out, size_t stride, uint32_t *cnt, int, uint32_t *) {
   for (int k = 0; k < ntiles; k++) {
      const ZbTileDesc t = tiles[first + k];
      const uint32_t *src = srcw + t.src_base;
      const uint32_t n = t.src_cnt_idx >= 0 ? src_cnt[t.src_cnt_idx] : t.src_n;
      const uint32_t lo = t.lo - t.src_lo, hi = t.hi - t.src_lo;
      uint32_t *dst = out + (size_t)k * stride;
      uint32_t carry = 0x1ff, c = 0;
      for (uint32_t r = 0; r < n; r++) {
         uint32_t w = src[r], pos = w & ZB_POS_MASK, l = (w >> ZB_POS_BITS) & 0x1ff;
         carry = std::min(carry, l);
         if (pos >= lo && pos < hi) { dst[c++] = (pos - lo) | (carry << ZB_POS_BITS); carry = 0x1ff; }
      }
      cnt[k] = c;
   }
}
