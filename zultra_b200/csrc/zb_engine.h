/*
 * zb_engine.h - host-side driver of one batch: builds the window list for a set of streams, runs the
 * pipeline stages and brings the stitched bitstreams back.
 */
#ifndef ZB_ENGINE_H
#define ZB_ENGINE_H
#include "zb_pipeline.h"

struct ZbStreamIn {
   const uint8_t *data; size_t n;          /* bytes to compress in this call (whole max-blocks except possibly the last) */
   const uint8_t *hist; uint32_t hist_len; /* <= 32768 bytes preceding data (previous block tail or preset dictionary) */
   int finalize;                           /* 1: the last window is the end of the stream (BFINAL) */
   uint32_t in_bits;                       /* bits already pending in the current output byte (0..7) */
};
struct ZbStreamRes { uint64_t total_bits; size_t out_off; int err; };

struct ZbDump {   /* optional stage dumps for tests (host copies) */
   std::vector<uint32_t> sa_lcp; std::vector<zb_match_t> match; std::vector<ZbSub> sub; std::vector<ZbSubTabs> tabs; std::vector<zb_match_t> best;
   std::vector<uint32_t> wbase;
};

static inline uint32_t zb_pick_tile(size_t total) {
   if (total <= ((size_t)8 << 20)) return 1024;
   if (total <= ((size_t)64 << 20)) return 2048;
   return 4096;
}

/* Returns 0 on success.  out receives, per stream, ceil(total_bits/8) bytes at res[i].out_off; the pending bits of the
   first byte are zero (the caller ORs its carried partial byte in). */
static inline int zb_run_batch(ZbPipe &p, const ZbStreamIn *s, int ns, uint32_t block_size, std::vector<uint8_t> &out,
                               std::vector<ZbStreamRes> &res, ZbDump *dump = 0, uint32_t tile_main = 0) {
   std::vector<ZbWinDesc> wins;
   std::vector<ZbStreamOut> so(ns);
   size_t in_bytes = 0;
   for (int i = 0; i < ns; i++) in_bytes += s[i].hist_len + s[i].n;
   std::vector<uint8_t> stage;   /* contiguous [hist|data] per stream; TODO pinned staging / direct copies */
   stage.resize(in_bytes);
   size_t off = 0, total_block = 0;
   for (int i = 0; i < ns; i++) {
      if (s[i].hist_len) memcpy(stage.data() + off, s[i].hist, s[i].hist_len);
      memcpy(stage.data() + off + s[i].hist_len, s[i].data, s[i].n);
      memset(&so[i], 0, sizeof(ZbStreamOut));
      so[i].first_win = (uint32_t)wins.size();
      so[i].in_bits = s[i].in_bits;
      size_t done = 0; uint32_t k = 0;
      while (done < s[i].n) {
         ZbWinDesc d; memset(&d, 0, sizeof(d));
         size_t blk = std::min((size_t)block_size, s[i].n - done);
         d.hist = k == 0 ? s[i].hist_len : ZB_HISTORY;
         d.in_off = (uint32_t)(off + s[i].hist_len + done - d.hist);
         d.len = d.hist + (uint32_t)blk;
         d.stream = (uint32_t)i;
         done += blk; k++;
         d.last = (done == s[i].n && s[i].finalize) ? 1 : 0;
         wins.push_back(d);
      }
      so[i].nwin = k;
      off += s[i].hist_len + s[i].n;
      total_block += s[i].n;
   }
   res.assign(ns, ZbStreamRes());
   if (wins.empty()) { out.clear(); return 0; }
   p.setup(wins, stage.data(), in_bytes, false);
   p.stage_sa();
   p.stage_match(tile_main ? tile_main : zb_pick_tile(total_block));
   p.stage_greedy();
   p.stage_split();
   p.stage_parse();
   p.stage_emit(so);
   /* bring the bitstreams back */
   size_t total_words = 0;
   for (int i = 0; i < ns; i++) total_words = std::max<size_t>(total_words, p.h_sout[i].out_word_off + (p.h_sout[i].total_bits + 31) / 32);
   std::vector<uint32_t> words(total_words + 1);
   zb_d2h(p.st, words.data(), p.out.p, total_words * 4);
   zb_sync(p.st);
   size_t ob = 0;
   for (int i = 0; i < ns; i++) { res[i].total_bits = p.h_sout[i].total_bits; res[i].out_off = ob; res[i].err = 0; ob += (size_t)((p.h_sout[i].total_bits + 7) / 8); }
   out.resize(ob);
   for (int i = 0; i < ns; i++) memcpy(out.data() + res[i].out_off, (const uint8_t *)(words.data() + p.h_sout[i].out_word_off), (size_t)((p.h_sout[i].total_bits + 7) / 8));
   if (dump) {
      dump->wbase = p.h_wbase;
      dump->sa_lcp.resize(p.P); zb_d2h(p.st, dump->sa_lcp.data(), p.sa_lcp.p, (size_t)p.P * 4);
      dump->match.resize((size_t)p.P * 8); zb_d2h(p.st, dump->match.data(), p.match.p, (size_t)p.P * 8 * sizeof(zb_match_t));
      dump->best.resize(p.P); zb_d2h(p.st, dump->best.data(), p.best.p, (size_t)p.P * sizeof(zb_match_t));
      dump->sub.resize(p.nsub); zb_d2h(p.st, dump->sub.data(), p.sub.p, sizeof(ZbSub) * p.nsub);
      dump->tabs.resize(p.nsub); zb_d2h(p.st, dump->tabs.data(), p.tabs.p, sizeof(ZbSubTabs) * p.nsub);
      zb_sync(p.st);
   }
   return 0;
}
#endif
