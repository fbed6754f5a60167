/*
 * libzultra.c - the libzultra.h API of zultra-b200: one-shot and zlib-like streaming compression in front
 * of the CUDA pipeline (C-ABI in zultra_cuda.h).
 *
 * Same observable contract as the reference state machine (libzultra.c:82-619): header first, deflate data
 * for whole max-blocks, footer after FINALIZE, ZULTRA_STREAM_END exactly once, identical output bytes.
 * What differs is pacing: input is staged on the host and handed to the GPU a BATCH of max-blocks at a time
 * (ZULTRA_CUDA_BATCH_BLOCKS, default 64; never more than that per engine call, whatever avail_in is), because one
 * block per launch would leave the device idle.  As in the reference a full block is only compressed once more
 * input is visible or on FINALIZE (libzultra.c:269), so block boundaries - and therefore the bytes - are the same.
 *
 * Overlap (reference libzultra.c:255-267,441-462 does input copy, compression and output
 * draining strictly one after the other): a batch is compressed by a worker thread (H2D, kernels, D2H) while
 * the calling thread keeps absorbing the caller's input into the other of two staging buffers and draining the
 * previous batch's output.  The bit phase and the running checksum travel from batch to batch inside the
 * worker's job record, so jobs are submitted in order and at most one is in flight.
 *
 * There is no CPU compressor in this library: no CUDA device -> ZULTRA_ERROR_COMPRESSION.
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "libzultra.h"
#include "zultra_cuda.h"

#define CSTATE_HAS_DICTIONARY 1
#define CSTATE_HEADER_EMITTED 2
#define CSTATE_FINALIZED 4
#define CSTATE_FOOTER_EMITTED 8
#define CSTATE_STREAM_ENDED 16
#define CSTATE_STARTED 32

/* one staged batch: [history (<= 32 KiB) | bytes]; n bytes of it (whole max-blocks unless finalizing) go to the GPU */
typedef struct {
   unsigned char *in; size_t in_cap, in_len; int hist_len;      /* staging buffer, bytes after the history, history length */
   unsigned char *out; size_t out_cap;                          /* compressed bytes of this batch */
   size_t n; int finalize;
   unsigned int bit_count_in; unsigned char bit_byte_in; unsigned int adler;   /* carried in, adler also carried out */
   unsigned long long bits; int rc;
   struct _zultra_compressor_s *owner;
   pthread_t th; int running;                                   /* 1: thread started and not yet joined, 2: ran inline (no thread could be made) */
   int done;                                                    /* set (release) by the worker when rc / bits / out are final */
} zb_job_t;

struct _zultra_compressor_s {
   unsigned int flags, block_size, state;
   const void *dict; int dict_size;
   zultra_cuda_ctx_t *ctx;
   unsigned int batch_blocks;
   zb_job_t job[2];
   int fill;                                         /* job[fill] is being filled; job[fill ^ 1] may be in flight */
   /* compressed bytes waiting to be drained (they live in a finished job's out buffer) */
   const unsigned char *out; size_t out_len, out_pos;
   unsigned int bit_count; unsigned char bit_byte;   /* partial byte carried between engine calls */
   unsigned int adler;                               /* checksum after the last collected batch */
   unsigned char frame[16]; size_t frame_pos, frame_len;
};

static void *def_alloc(void *o, unsigned int items, unsigned int size) { (void)o; return malloc((size_t)items * size); }
static void def_free(void *o, void *p) { (void)o; free(p); }

static unsigned int clamp_block(unsigned int b) {
   if (!b) b = ZULTRA_DEFAULT_MAX_BLOCK_SIZE;
   if (b < 32768) b = 32768;
   if (b > 2097152) b = 2097152;
   return b;
}

/* grow a host buffer through the stream's allocator (items*size is 32-bit in the reference's signature, so
   large buffers are requested as 64 KiB items) */
static int grow(zultra_stream_t *s, unsigned char **buf, size_t *cap, size_t keep, size_t need) {
   unsigned char *n;
   size_t c;
   if (need <= *cap) return 0;
   c = *cap ? *cap : 65536;
   while (c < need) c *= 2;
   n = (unsigned char *)s->zalloc(s->opaque, (unsigned int)((c + 65535) >> 16), 65536);
   if (!n) return -1;
   if (keep) memcpy(n, *buf, keep);
   if (*buf) s->zfree(s->opaque, *buf);
   *buf = n; *cap = c;
   return 0;
}

static int env_int(const char *name, int dflt) { const char *e = getenv(name); return (e && *e) ? atoi(e) : dflt; }

zultra_status_t zultra_stream_init(zultra_stream_t *pStream, const unsigned int nFlags, unsigned int nMaxBlockSize) {
   zultra_compressor_t *c;
   int nb, ndev;
   if (!pStream->zalloc) pStream->zalloc = def_alloc;
   if (!pStream->zfree) pStream->zfree = def_free;
   pStream->adler = 0;
   pStream->state = c = (zultra_compressor_t *)pStream->zalloc(pStream->opaque, 1, sizeof(zultra_compressor_t));
   if (!c) return ZULTRA_ERROR_MEMORY;
   memset(c, 0, sizeof(*c));
   c->flags = nFlags;
   c->block_size = clamp_block(nMaxBlockSize);
   ndev = env_int("ZULTRA_CUDA_DEVICES", 1);
   if (ndev < 1) ndev = 1;
   nb = env_int("ZULTRA_CUDA_BATCH_BLOCKS", 0);
   c->batch_blocks = nb > 0 ? (unsigned int)nb : 64u * (unsigned int)ndev;
   if ((size_t)c->batch_blocks * c->block_size > ((size_t)512 << 20) * (size_t)ndev) c->batch_blocks = (unsigned int)(((size_t)512 << 20) * (size_t)ndev / c->block_size);
   c->job[0].owner = c->job[1].owner = c;
   if (zultra_cuda_ctx_acquire(&c->ctx, env_int("ZULTRA_CUDA_DEVICE", -1)) != 0) {
      /* no usable GPU: this library has no other way to compress */
      zultra_stream_end(pStream);
      return ZULTRA_ERROR_COMPRESSION;
   }
   zultra_cuda_ctx_set_devices(c->ctx, ndev);
   return ZULTRA_OK;
}

zultra_status_t zultra_stream_set_dictionary(zultra_stream_t *pStream, const void *pDictionaryData, const int nDictionaryDataSize) {
   zultra_compressor_t *c = pStream->state;
   if (c && c->state == 0) {
      c->dict = pDictionaryData; c->dict_size = nDictionaryDataSize;
      c->state |= CSTATE_HAS_DICTIONARY;
      return ZULTRA_OK;
   }
   return ZULTRA_ERROR_COMPRESSION;
}

/* worker: one engine call for one staged batch */
static void *job_main(void *arg) {
   zb_job_t *j = (zb_job_t *)arg;
   zultra_compressor_t *c = j->owner;
   j->rc = zultra_cuda_compress_blocks(c->ctx, j->in, j->hist_len, j->in + j->hist_len, j->n, c->block_size, j->finalize, j->bit_count_in, c->flags,
                                       &j->adler, j->out, j->out_cap, &j->bits);
   __atomic_store_n(&j->done, 1, __ATOMIC_RELEASE);
   return NULL;
}

/* wait for the batch in flight (if any) and make its bytes the ones being drained */
static zultra_status_t collect(zultra_stream_t *s) {
   zultra_compressor_t *c = s->state;
   zb_job_t *j = &c->job[c->fill ^ 1];
   if (!j->running) return ZULTRA_OK;
   if (j->running == 1) pthread_join(j->th, NULL);
   j->running = 0;
   if (j->rc) return j->rc == ZULTRA_CUDA_ERR_DST ? ZULTRA_ERROR_DST : ZULTRA_ERROR_COMPRESSION;
   j->out[0] |= j->bit_byte_in;
   c->out = j->out; c->out_pos = 0;
   c->out_len = (size_t)(j->bits >> 3);
   c->bit_count = (unsigned int)(j->bits & 7);
   c->bit_byte = c->bit_count ? j->out[c->out_len] : 0;
   if (j->finalize && c->bit_count) { c->out_len++; c->bit_count = 0; c->bit_byte = 0; }   /* flush_bits, libzultra.c:416 */
   s->adler = c->adler = j->adler;
   if (j->finalize) c->state |= CSTATE_FINALIZED;
   return ZULTRA_OK;
}

/* Hand the first n staged bytes of the fill buffer to the worker; the rest, preceded by the last 32 KiB as history
   (libzultra.c:406-412), moves to the other buffer, which becomes the fill buffer.  The caller has collected AND drained
   the previous batch: its staging buffer is reused right now, its out buffer one batch later. */
static zultra_status_t submit(zultra_stream_t *s, size_t n, int finalize) {
   zultra_compressor_t *c = s->state;
   zb_job_t *j = &c->job[c->fill], *o = &c->job[c->fill ^ 1];
   const size_t total = (size_t)j->hist_len + n, rest = j->in_len - n;
   const size_t keep = total > HISTORY_SIZE ? HISTORY_SIZE : total;
   const size_t cap = n + n / 8 + 4096 + (n / c->block_size + 1) * 64 * 6;
   if (grow(s, &j->out, &j->out_cap, 0, cap)) return ZULTRA_ERROR_MEMORY;
   if (!finalize) {
      if (grow(s, &o->in, &o->in_cap, 0, (size_t)HISTORY_SIZE + (size_t)c->batch_blocks * c->block_size + 65536)) return ZULTRA_ERROR_MEMORY;   /* its final size at once */
      memcpy(o->in, j->in + total - keep, keep + rest);
      o->hist_len = (int)keep; o->in_len = rest;
   } else { o->hist_len = 0; o->in_len = 0; }
   j->n = n; j->finalize = finalize;
   j->bit_count_in = c->bit_count; j->bit_byte_in = c->bit_byte; j->adler = c->adler;
   j->rc = 0; j->bits = 0; j->done = 0;
   if (pthread_create(&j->th, NULL, job_main, j) == 0) j->running = 1;
   else { job_main(j); j->running = 2; }      /* no thread to be had: same result, no overlap */
   c->fill ^= 1;
   return ZULTRA_OK;
}

static void drain(zultra_stream_t *s, const unsigned char *src, size_t *pos, size_t len) {
   size_t k = len - *pos;
   if (k > s->avail_out) k = s->avail_out;
   if (!k) return;
   memcpy(s->next_out, src + *pos, k);
   *pos += k; s->next_out += k; s->avail_out -= k; s->total_out += k;
}

zultra_status_t zultra_stream_compress(zultra_stream_t *pStream, const int nDoFinalize) {
   zultra_compressor_t *c = pStream->state;
   zultra_status_t err = ZULTRA_OK;
   size_t batch;
   if (!c || (c->state & CSTATE_STREAM_ENDED)) return ZULTRA_ERROR_COMPRESSION;
   batch = (size_t)c->batch_blocks * c->block_size;

   if (!(c->state & CSTATE_HEADER_EMITTED)) {
      int n = zultra_frame_encode_header(c->frame, 16, c->flags, c->dict, c->dict_size);
      c->state |= CSTATE_HEADER_EMITTED | CSTATE_STARTED;
      if (n < 0) return ZULTRA_ERROR_COMPRESSION;
      c->frame_pos = 0; c->frame_len = (size_t)n;
      pStream->adler = c->adler = zultra_frame_init_checksum(c->flags);
      if (c->dict && c->dict_size > 0) {   /* dictionary becomes the first history (libzultra.c:250-253) */
         const unsigned char *d = (const unsigned char *)c->dict;
         zb_job_t *f = &c->job[c->fill];
         int k = c->dict_size > HISTORY_SIZE ? HISTORY_SIZE : c->dict_size;
         if (grow(pStream, &f->in, &f->in_cap, 0, (size_t)HISTORY_SIZE + c->block_size)) return ZULTRA_ERROR_MEMORY;
         memcpy(f->in, d + c->dict_size - k, (size_t)k);
         f->hist_len = k;
      }
   }

   for (;;) {
      zb_job_t *f, *o;
      /* 1. frame bytes (header or footer) */
      if (c->frame_pos < c->frame_len) {
         drain(pStream, c->frame, &c->frame_pos, c->frame_len);
         if (c->frame_pos < c->frame_len) break;   /* caller must provide more room */
      }
      /* 2. compressed bytes of the last collected batch */
      if (c->out_pos < c->out_len) {
         drain(pStream, c->out, &c->out_pos, c->out_len);
         if (c->out_pos < c->out_len) break;
      }
      if (c->state & CSTATE_FOOTER_EMITTED) break;
      if (c->state & CSTATE_FINALIZED) {
         int n = zultra_frame_encode_footer(c->frame, 16, pStream->adler, (long long)pStream->total_in, c->flags);
         if (n < 0) { err = ZULTRA_ERROR_COMPRESSION; break; }
         c->state = (c->state | CSTATE_FOOTER_EMITTED) & ~CSTATE_FINALIZED;
         c->frame_pos = 0; c->frame_len = (size_t)n;
         continue;
      }
      f = &c->job[c->fill]; o = &c->job[c->fill ^ 1];
      /* 3. a batch that finished in the meantime: its bytes can flow now */
      if (o->running && __atomic_load_n(&o->done, __ATOMIC_ACQUIRE)) { err = collect(pStream); if (err) break; continue; }
      /* 4. absorb caller input, never more than one batch plus the byte that proves the last block has a successor */
      if (pStream->avail_in && f->in_len <= batch) {
         size_t n = batch + 1 - f->in_len;
         if (n > pStream->avail_in) n = pStream->avail_in;
         /* the staging buffer is asked for at its final size the first time (untouched pages cost nothing; doubling up to
            it copied and faulted in as much again as the batch itself) */
         if (grow(pStream, &f->in, &f->in_cap, (size_t)f->hist_len + f->in_len, f->in_len ? (size_t)f->hist_len + f->in_len + n : (size_t)HISTORY_SIZE + batch + 65536)) { err = ZULTRA_ERROR_MEMORY; break; }
         memcpy(f->in + f->hist_len + f->in_len, pStream->next_in, n);
         f->in_len += n; pStream->next_in += n; pStream->avail_in -= n; pStream->total_in += n;
      }
      /* 5. compress: a whole batch once a further byte is staged; everything that is left on FINALIZE */
      if (f->in_len > batch) {
         if (o->running) { err = collect(pStream); if (err) break; continue; }      /* one batch in flight: wait, drain, come back */
         err = submit(pStream, batch, 0); if (err) break;
         continue;
      }
      if (pStream->avail_in) continue;
      if (!nDoFinalize) break;
      if (o->running) { err = collect(pStream); if (err) break; continue; }
      if (!f->in_len) break;   /* nothing was ever staged: the reference emits no block either (libzultra.c:275) */
      err = submit(pStream, f->in_len, 1); if (err) break;
      err = collect(pStream); if (err) break;
   }
   if (err) return err;
   if ((c->state & CSTATE_FOOTER_EMITTED) && c->frame_pos >= c->frame_len && c->out_pos >= c->out_len) {
      c->state |= CSTATE_STREAM_ENDED;
      return ZULTRA_STREAM_END;
   }
   return ZULTRA_OK;
}

void zultra_stream_end(zultra_stream_t *pStream) {
   if (pStream->state && pStream->zfree) {
      zultra_compressor_t *c = pStream->state;
      int k;
      zultra_cuda_trace("stream_end: begin");
      for (k = 0; k < 2; k++) if (c->job[k].running == 1) { pthread_join(c->job[k].th, NULL); c->job[k].running = 0; }
      if (c->ctx) zultra_cuda_ctx_release(c->ctx);      /* a context that saw a CUDA failure is destroyed there, not pooled */
      zultra_cuda_trace("stream_end: context released");
      for (k = 0; k < 2; k++) {
         if (c->job[k].in) pStream->zfree(pStream->opaque, c->job[k].in);
         if (c->job[k].out) pStream->zfree(pStream->opaque, c->job[k].out);
      }
      pStream->zfree(pStream->opaque, c);
      pStream->state = NULL;
      zultra_cuda_trace("stream_end: buffers freed");
   }
}

size_t zultra_memory_bound(size_t nInputSize, const unsigned int nFlags, unsigned int nMaxBlockSize) {
   nMaxBlockSize = clamp_block(nMaxBlockSize);   /* same formula as libzultra.c:586 */
   return (size_t)zultra_frame_get_header_size(nFlags, NULL, 0) + ((nInputSize + (nMaxBlockSize - 1)) / nMaxBlockSize) * 6 * 64 + nInputSize + 1 +
          (size_t)zultra_frame_get_footer_size(nFlags);
}

/*
 * One-shot: no staging copy - the caller's buffer goes to the GPU(s) in batches of up to 256 max-blocks per device with
 * the preceding 32 KiB as history and the bit phase carried across batches.  Returns (size_t)-1 on any failure,
 * including empty input and an output buffer that is too small (libzultra.c:608,617).
 */
size_t zultra_memory_compress(const unsigned char *pInputData, size_t nInputSize, unsigned char *pOutBuffer, size_t nMaxOutBufferSize,
                              const unsigned int nFlags, unsigned int nMaxBlockSize) {
   zultra_cuda_ctx_t *ctx = NULL;
   const unsigned int block = clamp_block(nMaxBlockSize);
   int ndev = env_int("ZULTRA_CUDA_DEVICES", 1);
   size_t batch;
   unsigned int ck = zultra_frame_init_checksum(nFlags), bit_count = 0;
   unsigned char bit_byte = 0;
   size_t w = 0, done = 0;
   int n;
   if (ndev < 1) ndev = 1;
   batch = (size_t)256 * block > ((size_t)256 << 20) ? ((size_t)256 << 20) / block * block : (size_t)256 * block;
   batch *= (size_t)ndev;
   if (!nInputSize || !pInputData || !pOutBuffer) return (size_t)-1;
   n = zultra_frame_encode_header(pOutBuffer, nMaxOutBufferSize > 16 ? 16 : (int)nMaxOutBufferSize, nFlags, NULL, 0);
   if (n < 0) return (size_t)-1;
   w = (size_t)n;
   if (zultra_cuda_ctx_acquire(&ctx, env_int("ZULTRA_CUDA_DEVICE", -1)) != 0) return (size_t)-1;
   zultra_cuda_ctx_set_devices(ctx, ndev);
   while (done < nInputSize) {
      size_t k = nInputSize - done > batch ? batch : nInputSize - done;
      const int fin = done + k == nInputSize;
      const int hist = done > HISTORY_SIZE ? HISTORY_SIZE : (int)done;
      unsigned long long bits = 0;
      unsigned char first = 0;
      /* the engine writes whole bytes starting at the byte that holds the pending bits */
      if (bit_count) { w--; first = bit_byte; }
      if (zultra_cuda_compress_blocks(ctx, pInputData + done - hist, hist, pInputData + done, k, block, fin, bit_count, nFlags, &ck,
                                      pOutBuffer + w, nMaxOutBufferSize - w, &bits) != 0) { zultra_cuda_ctx_release(ctx); return (size_t)-1; }
      pOutBuffer[w] |= first;
      w += (size_t)((bits + 7) >> 3);
      bit_count = (unsigned int)(bits & 7);
      bit_byte = bit_count ? pOutBuffer[w - 1] : 0;
      done += k;
   }
   zultra_cuda_ctx_release(ctx);
   if (nMaxOutBufferSize - w < (size_t)zultra_frame_get_footer_size(nFlags)) return (size_t)-1;
   n = zultra_frame_encode_footer(pOutBuffer + w, 16, ck, (long long)nInputSize, nFlags);
   if (n < 0) return (size_t)-1;
   return w + (size_t)n;
}
