/* zultra_oracle.h - entry points of the plain-C restatement (TEST INFRASTRUCTURE ONLY, see zultra_oracle.c). */
#ifndef ZULTRA_ORACLE_H
#define ZULTRA_ORACLE_H
typedef struct {
   int max_sub, nsub;
   int *end, *is_dyn, *static_cost, *dynamic_cost, *body_bits;   /* per sub-block (max_sub entries) */
   int *lit_len, *off_len;                                       /* 288 / 32 ints per sub-block */
   void *best; int best_cap;                                     /* {u16 len,u16 off} per window position of the first max-block */
} zo_stage_dump_t;
/* packed SA|LCP words of one window (matchfinder.c:49-90) */
int zo_window_sa_lcp(const unsigned char *t, int n, unsigned int *words);
/* match[(i-hist)*8+m] = {u16 length,u16 offset} (matchfinder.c:262-286) */
int zo_window_matches(const unsigned char *t, int hist, int n, unsigned short *out);
/* whole stream, same result as zultra_memory_compress (libzultra.c:601) / with a preset dictionary; -1 on failure */
long zo_compress(const unsigned char *in, long n, const unsigned char *dict, int dict_size, unsigned char *out, long cap, unsigned flags, unsigned block, zo_stage_dump_t *dump);
#endif
