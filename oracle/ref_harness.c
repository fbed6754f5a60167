/*
 * ref_harness.c - stage dumper linked INTO the compiled reference (oracle/_ref/libzultra_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (zultra_b200/) may link, load or call this.
 * This file is ours; it contains no reference code.  It includes the reference's own headers from
 * where they lie (-I/root/reference/src) and calls the reference's extern stage functions so that
 * tests can diff our CUDA stages against the real thing:
 *
 *   refh_window_sa_lcp   -> divsufsort_build_array (divsufsort.c:377) + the packed "SA_and_LCP" word as it
 *                           stands after matchfinder.c:90 (the interval build at :98-155 overwrites it
 *                           in place, so the LCP/pack step is recomputed here by direct comparison).
 *   refh_window_matches  -> zultra_build_suffix_array / zultra_skip_matches / zultra_find_all_matches
 *                           (matchfinder.c:49,243,262) exactly as libzultra.c:287-293 drives them.
 *   refh_block_stages    -> zultra_block_split + per sub-block static/dynamic decision + zultra_block_deflate
 *                           as libzultra.c:303-343 drives them, each sub-block written at bit phase 0.
 */
#include <stdlib.h>
#include <string.h>
#include "libzultra.h"
#include "matchfinder.h"
#include "blockdeflate.h"
#include "format.h"
#include "private.h"

#define EXPORT __attribute__((visibility("default")))

EXPORT int refh_window_sa_lcp(const unsigned char *win, int n, unsigned int *out_words) {
   zultra_stream_t strm;
   int *sa, r;
   if (n <= 0) return 0;
   memset(&strm, 0, sizeof(strm));
   if (zultra_stream_init(&strm, 0, 2097152) != ZULTRA_OK) return -1;
   sa = (int *)malloc(sizeof(int) * (size_t)n);
   if (divsufsort_build_array(&strm.state->divsufsort_context, win, sa, n) != 0) { free(sa); zultra_stream_end(&strm); return -2; }
   out_words[0] = (unsigned int)sa[0];
   for (r = 1; r < n; r++) {
      int a = sa[r], b = sa[r - 1], lim = n - (a > b ? a : b), l = 0;
      while (l < lim && l < 259 && win[a + l] == win[b + l]) l++;
      if (l < MIN_MATCH_SIZE) l = 0;
      if (l > MAX_MATCH_SIZE) l = MAX_MATCH_SIZE;
      out_words[r] = (unsigned int)a | ((unsigned int)l << LCP_SHIFT);
   }
   free(sa);
   zultra_stream_end(&strm);
   return n;
}

/* out_match: (n_total - hist) * 8 entries of {u16 length, u16 offset} */
EXPORT int refh_window_matches(const unsigned char *win, int hist, int n_total, unsigned short *out_match) {
   zultra_stream_t strm;
   memset(&strm, 0, sizeof(strm));
   if (zultra_stream_init(&strm, 0, 2097152) != ZULTRA_OK) return -1;
   if (zultra_build_suffix_array(strm.state, win, n_total)) { zultra_stream_end(&strm); return -2; }
   if (hist) zultra_skip_matches(strm.state, 0, hist);
   zultra_find_all_matches(strm.state, hist, n_total);
   memcpy(out_match, strm.state->match + ((size_t)hist << MATCHES_PER_OFFSET_SHIFT),
          (size_t)(n_total - hist) * NMATCHES_PER_OFFSET * sizeof(zultra_match_t));
   zultra_stream_end(&strm);
   return 0;
}

/*
 * One max-block: window = [hist bytes of history | n_total-hist block bytes].
 * Outputs (arrays sized for 64 sub-blocks):
 *   split_end[k]   exclusive end offset (window coordinates) of sub-block k
 *   is_dynamic[k], static_cost[k], dynamic_cost[k]  the estimate-based decision (libzultra.c:317-324)
 *   lit_len[k*288..], off_len[k*32..]  final nCodeLength of both encoders after zultra_block_deflate
 *   body_bits[k]   bits zultra_block_deflate wrote after the 3 header bits, starting at bit phase 0 (-1 = failed)
 *   best_match     n_total entries {u16 length,u16 offset} as left by the last sub-block run over each range
 *   body           concatenated byte-padded sub-block bodies (3 header bits with BFINAL=0 included), body_off[k] byte offsets
 * Returns the number of sub-blocks or <0.
 */
EXPORT int refh_block_stages(const unsigned char *win, int hist, int n_total,
                             int *split_end, int *is_dynamic, int *static_cost, int *dynamic_cost,
                             int *lit_len, int *off_len, int *body_bits, unsigned short *best_match,
                             unsigned char *body, int body_cap, int *body_off) {
   zultra_stream_t strm;
   zultra_compressor_t *c;
   int nsplit, k, start, i, used = 0;
   memset(&strm, 0, sizeof(strm));
   if (zultra_stream_init(&strm, 0, 2097152) != ZULTRA_OK) return -1;
   c = strm.state;
   if (zultra_build_suffix_array(c, win, n_total)) { zultra_stream_end(&strm); return -2; }
   if (hist) zultra_skip_matches(c, 0, hist);
   zultra_find_all_matches(c, hist, n_total);
   nsplit = zultra_block_split(c, win, hist, n_total - hist, MAX_SPLITS, split_end);
   if (nsplit < 0) { zultra_stream_end(&strm); return -3; }
   start = hist;
   for (k = 0; k < nsplit; k++) {
      int size = split_end[k] - start, sc = 0, dc = 0, dyn = 1, res;
      zultra_bitwriter_t bw;
      zultra_block_prepare_cost_evaluation(c, win, start, size);
      zultra_block_evaluate_static_cost(&c->literalsEncoder, &c->offsetEncoder, &sc);
      zultra_huffman_encoder_estimate_dynamic_codelens(&c->literalsEncoder);
      zultra_huffman_encoder_estimate_dynamic_codelens(&c->offsetEncoder);
      zultra_block_evaluate_dynamic_cost(&c->literalsEncoder, &c->offsetEncoder, &dc);
      if (sc <= dc) dyn = 0;
      is_dynamic[k] = dyn; static_cost[k] = sc; dynamic_cost[k] = dc;
      body_off[k] = used;
      zultra_bitwriter_init(&bw, body + used, 0, body_cap - used);
      zultra_bitwriter_put_bits(&bw, 0, 1);
      zultra_bitwriter_put_bits(&bw, 1 + dyn, 2);
      res = zultra_block_deflate(c, &bw, win, start, size, dyn);
      if (res < 0 || zultra_bitwriter_get_offset(&bw) < 0) body_bits[k] = -1;
      else {
         body_bits[k] = bw.nOutOffset * 8 + bw.nEncBitCount - 3;
         zultra_bitwriter_flush_bits(&bw);
         used += bw.nOutOffset;
      }
      for (i = 0; i < NLITERALSYMS; i++) lit_len[k * NLITERALSYMS + i] = c->literalsEncoder.nCodeLength[i];
      for (i = 0; i < NOFFSETSYMS; i++) off_len[k * NOFFSETSYMS + i] = c->offsetEncoder.nCodeLength[i];
      memcpy(best_match + 2 * (size_t)start, c->best_match + start, (size_t)size * sizeof(zultra_match_t));
      start = split_end[k];
   }
   body_off[nsplit] = used;
   zultra_stream_end(&strm);
   return nsplit;
}

/* Streaming entry with dictionary, for the FDICT path (libzultra.c:177-190). */
EXPORT long refh_compress_with_dict(const unsigned char *in, long n, const unsigned char *dict, int dict_size,
                                    unsigned char *out, long out_cap, unsigned flags, unsigned block_size) {
   zultra_stream_t strm;
   zultra_status_t st;
   memset(&strm, 0, sizeof(strm));
   if (zultra_stream_init(&strm, flags, block_size) != ZULTRA_OK) return -1;
   if (dict && dict_size && zultra_stream_set_dictionary(&strm, dict, dict_size) != ZULTRA_OK) { zultra_stream_end(&strm); return -1; }
   strm.next_in = in; strm.avail_in = (size_t)n; strm.next_out = out; strm.avail_out = (size_t)out_cap;
   st = zultra_stream_compress(&strm, ZULTRA_FINALIZE);
   zultra_stream_end(&strm);
   if (st != ZULTRA_STREAM_END) return -1;
   return (long)(out_cap - (long)strm.avail_out);
}
