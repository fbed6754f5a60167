/*
 * libzultra.c - the libzultra.h API of zultra-b200: one-shot and zlib-like streaming compression in front
 * of the CUDA pipeline (C-ABI in zultra_cuda.h).
 *
 * Same observable contract as the reference state machine (libzultra.c:82-619): header first, deflate data
 * for whole max-blocks, footer after FINALIZE, ZULTRA_STREAM_END exactly once, identical output bytes.
 * What differs is pacing: input is staged on the host and handed to the GPU many max-blocks at a time
 * (ZULTRA_CUDA_BATCH_BLOCKS, default 64), because one block per launch would leave the device idle.  As in the
 * reference a full block is only compressed once more input is visible or on FINALIZE (libzultra.c:269), so
 * block boundaries - and therefore the bytes - are the same.  There is no CPU compressor in this library.
 */
#include <stdlib.h>
#include <string.h>
#include "libzultra.h"
#include "zultra_cuda.h"

#define CSTATE_HAS_DICTIONARY 1
#define CSTATE_HEADER_EMITTED 2
#define CSTATE_FINALIZED 4
#define CSTATE_FOOTER_EMITTED 8
#define CSTATE_STREAM_ENDED 16
#define CSTATE_STARTED 32

struct _zultra_compressor_s {
   unsigned int flags, block_size, state;
   const void *dict; int dict_size;
   zultra_cuda_ctx_t *ctx;
   /* staged input: [history (<=32K) | pending bytes] */
   unsigned char *in; size_t in_cap, in_len; int hist_len;
   unsigned int batch_blocks;
   /* compressed bytes waiting to be drained */
   unsigned char *out; size_t out_cap, out_len, out_pos;
   unsigned int bit_count; unsigned char bit_byte;   /* partial byte carried between engine calls */
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

zultra_status_t zultra_stream_init(zultra_stream_t *pStream, const unsigned int nFlags, unsigned int nMaxBlockSize) {
   zultra_compressor_t *c;
   const char *e;
   if (!pStream->zalloc) pStream->zalloc = def_alloc;
   if (!pStream->zfree) pStream->zfree = def_free;
   pStream->adler = 0;
   pStream->state = c = (zultra_compressor_t *)pStream->zalloc(pStream->opaque, 1, sizeof(zultra_compressor_t));
   if (!c) return ZULTRA_ERROR_MEMORY;
   memset(c, 0, sizeof(*c));
   c->flags = nFlags;
   c->block_size = clamp_block(nMaxBlockSize);
   c->batch_blocks = 64;
   e = getenv("ZULTRA_CUDA_BATCH_BLOCKS");
   if (e && atoi(e) > 0) c->batch_blocks = (unsigned int)atoi(e);
   if ((size_t)c->batch_blocks * c->block_size > ((size_t)512 << 20)) c->batch_blocks = (unsigned int)(((size_t)512 << 20) / c->block_size);
   e = getenv("ZULTRA_CUDA_DEVICE");
   if (zultra_cuda_ctx_acquire(&c->ctx, e ? atoi(e) : -1) != 0) {
      /* no usable GPU: this library has no other way to compress */
      zultra_stream_end(pStream);
      return ZULTRA_ERROR_COMPRESSION;
   }
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

/* hand `n` staged bytes (whole max-blocks unless finalizing) to the GPU */
static zultra_status_t run_gpu(zultra_stream_t *s, size_t n, int finalize) {
   zultra_compressor_t *c = s->state;
   unsigned long long bits = 0;
   size_t cap = n + n / 8 + 4096 + (n / c->block_size + 1) * 64 * 6;
   int rc;
   if (grow(s, &c->out, &c->out_cap, 0, cap)) return ZULTRA_ERROR_MEMORY;
   rc = zultra_cuda_compress_blocks(c->ctx, c->in, c->hist_len, c->in + c->hist_len, n, c->block_size, finalize, c->bit_count, c->flags,
                                    &s->adler, c->out, c->out_cap, &bits);
   if (rc) return rc == ZULTRA_CUDA_ERR_DST ? ZULTRA_ERROR_DST : ZULTRA_ERROR_COMPRESSION;
   c->out[0] |= c->bit_byte;
   c->out_pos = 0;
   c->out_len = (size_t)(bits >> 3);
   c->bit_count = (unsigned int)(bits & 7);
   c->bit_byte = c->bit_count ? c->out[c->out_len] : 0;
   if (finalize && c->bit_count) { c->out_len++; c->bit_count = 0; c->bit_byte = 0; }   /* flush_bits, libzultra.c:416 */
   /* slide: keep the last 32 KiB as history (libzultra.c:406-412) */
   {
      size_t total = (size_t)c->hist_len + n, rest = c->in_len - n;
      size_t keep = total > HISTORY_SIZE ? HISTORY_SIZE : total;
      memmove(c->in, c->in + total - keep, keep + rest);
      c->hist_len = (int)keep;
      c->in_len = rest;
   }
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
   if (!c || (c->state & CSTATE_STREAM_ENDED)) return ZULTRA_ERROR_COMPRESSION;

   if (!(c->state & CSTATE_HEADER_EMITTED)) {
      int n = zultra_frame_encode_header(c->frame, 16, c->flags, c->dict, c->dict_size);
      c->state |= CSTATE_HEADER_EMITTED | CSTATE_STARTED;
      if (n < 0) return ZULTRA_ERROR_COMPRESSION;
      c->frame_pos = 0; c->frame_len = (size_t)n;
      pStream->adler = zultra_frame_init_checksum(c->flags);
      if (c->dict && c->dict_size > 0) {   /* dictionary becomes the first history (libzultra.c:250-253) */
         const unsigned char *d = (const unsigned char *)c->dict;
         int k = c->dict_size > HISTORY_SIZE ? HISTORY_SIZE : c->dict_size;
         if (grow(pStream, &c->in, &c->in_cap, 0, (size_t)HISTORY_SIZE + c->block_size)) return ZULTRA_ERROR_MEMORY;
         memcpy(c->in, d + c->dict_size - k, (size_t)k);
         c->hist_len = k;
      }
   }

   for (;;) {
      /* 1. frame bytes (header or footer) */
      if (c->frame_pos < c->frame_len) {
         drain(pStream, c->frame, &c->frame_pos, c->frame_len);
         if (c->frame_pos < c->frame_len) break;   /* caller must provide more room */
      }
      /* 2. compressed bytes */
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
      /* 3. absorb all caller input */
      if (pStream->avail_in) {
         size_t n = pStream->avail_in;
         if (grow(pStream, &c->in, &c->in_cap, (size_t)c->hist_len + c->in_len, (size_t)c->hist_len + c->in_len + n)) { err = ZULTRA_ERROR_MEMORY; break; }
         memcpy(c->in + c->hist_len + c->in_len, pStream->next_in, n);
         c->in_len += n; pStream->next_in += n; pStream->avail_in = 0; pStream->total_in += n;
      }
      /* 4. compress: everything on FINALIZE, else whole blocks while at least one further byte is staged */
      if (nDoFinalize) {
         if (c->in_len) { err = run_gpu(pStream, c->in_len, 1); if (err) break; c->state |= CSTATE_FINALIZED; continue; }
         break;   /* nothing was ever staged for this call: the reference emits no block either (libzultra.c:275) */
      } else {
         size_t full = (c->in_len - 1) / c->block_size;   /* blocks that have a successor byte */
         if (c->in_len && full >= c->batch_blocks) { err = run_gpu(pStream, full * c->block_size, 0); if (err) break; continue; }
         break;
      }
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
      if (c->ctx) zultra_cuda_ctx_release(c->ctx);
      if (c->in) pStream->zfree(pStream->opaque, c->in);
      if (c->out) pStream->zfree(pStream->opaque, c->out);
      pStream->zfree(pStream->opaque, c);
      pStream->state = NULL;
   }
}

size_t zultra_memory_bound(size_t nInputSize, const unsigned int nFlags, unsigned int nMaxBlockSize) {
   nMaxBlockSize = clamp_block(nMaxBlockSize);   /* same formula as libzultra.c:586 */
   return (size_t)zultra_frame_get_header_size(nFlags, NULL, 0) + ((nInputSize + (nMaxBlockSize - 1)) / nMaxBlockSize) * 6 * 64 + nInputSize + 1 +
          (size_t)zultra_frame_get_footer_size(nFlags);
}

/*
 * One-shot: no staging copy - the caller's buffer goes to the GPU in batches of up to 256 max-blocks with the
 * preceding 32 KiB as history and the bit phase carried across batches.  Returns (size_t)-1 on any failure,
 * including empty input and an output buffer that is too small (libzultra.c:608,617).
 */
size_t zultra_memory_compress(const unsigned char *pInputData, size_t nInputSize, unsigned char *pOutBuffer, size_t nMaxOutBufferSize,
                              const unsigned int nFlags, unsigned int nMaxBlockSize) {
   zultra_cuda_ctx_t *ctx = NULL;
   const unsigned int block = clamp_block(nMaxBlockSize);
   const size_t batch = (size_t)256 * block > ((size_t)256 << 20) ? ((size_t)256 << 20) / block * block : (size_t)256 * block;
   unsigned int ck = zultra_frame_init_checksum(nFlags), bit_count = 0;
   unsigned char bit_byte = 0;
   size_t w = 0, done = 0;
   int n;
   const char *e = getenv("ZULTRA_CUDA_DEVICE");
   if (!nInputSize || !pInputData || !pOutBuffer) return (size_t)-1;
   n = zultra_frame_encode_header(pOutBuffer, nMaxOutBufferSize > 16 ? 16 : (int)nMaxOutBufferSize, nFlags, NULL, 0);
   if (n < 0) return (size_t)-1;
   w = (size_t)n;
   if (zultra_cuda_ctx_acquire(&ctx, e ? atoi(e) : -1) != 0) return (size_t)-1;
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
