/*
 * libzultra.h - public API of zultra-b200: a drop-in for emmanuel-marty/zultra's src/libzultra.h
 * (reference: libzultra.h:54-157, frame.h, dictionary.h, format.h).  Same types, constants, struct layout and
 * function names; the compression work behind them runs on an NVIDIA B200 through the C-ABI in zultra_cuda.h.
 * There is no CPU fallback: if no CUDA device is usable zultra_stream_init returns ZULTRA_ERROR_COMPRESSION.
 */
#ifndef ZULTRA_B200_LIBZULTRA_H
#define ZULTRA_B200_LIBZULTRA_H
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* deflate format limits (reference format.h:37-50) */
#define MIN_MATCH_SIZE 3
#define MAX_MATCH_SIZE 258
#define MIN_OFFSET 1
#define MAX_OFFSET 32768
#define HISTORY_SIZE 0x8000

typedef struct _zultra_compressor_s zultra_compressor_t;

/* status codes (reference libzultra.h:54-60) */
typedef enum _zultra_stream_e {
   ZULTRA_OK = 0,
   ZULTRA_STREAM_END,
   ZULTRA_ERROR_SRC = -1,
   ZULTRA_ERROR_DST = -2,
   ZULTRA_ERROR_DICTIONARY = -3,
   ZULTRA_ERROR_MEMORY = -4,
   ZULTRA_ERROR_COMPRESSION = -5
} zultra_status_t;

/* framing flags (reference libzultra.h:64-66) */
#define ZULTRA_FLAG_DEFLATE_FRAMING 0
#define ZULTRA_FLAG_ZLIB_FRAMING    1
#define ZULTRA_FLAG_GZIP_FRAMING    2

#define ZULTRA_CONTINUE 0
#define ZULTRA_FINALIZE 1
#define ZULTRA_DEFAULT_MAX_BLOCK_SIZE 1048576

typedef unsigned int zultra_frame_checksum_t;

/* streaming context; field order and types as reference libzultra.h:78-93 */
typedef struct _zultra_stream_s {
   const unsigned char *next_in;
   size_t avail_in;
   unsigned long long total_in;

   unsigned char *next_out;
   size_t avail_out;
   unsigned long long total_out;

   void *(*zalloc)(void *opaque, unsigned int items, unsigned int size);
   void (*zfree)(void *opaque, void *address);
   void *opaque;

   zultra_compressor_t *state;
   zultra_frame_checksum_t adler;
} zultra_stream_t;

zultra_status_t zultra_stream_init(zultra_stream_t *pStream, const unsigned int nFlags, unsigned int nMaxBlockSize);
zultra_status_t zultra_stream_set_dictionary(zultra_stream_t *pStream, const void *pDictionaryData, const int nDictionaryDataSize);
zultra_status_t zultra_stream_compress(zultra_stream_t *pStream, const int nDoFinalize);
void zultra_stream_end(zultra_stream_t *pStream);
size_t zultra_memory_bound(size_t nInputSize, const unsigned int nFlags, unsigned int nMaxBlockSize);
size_t zultra_memory_compress(const unsigned char *pInputData, size_t nInputSize, unsigned char *pOutBuffer, size_t nMaxOutBufferSize,
                              const unsigned int nFlags, unsigned int nMaxBlockSize);

/* framing (reference frame.h) */
#define ZULTRA_ENCODE_ERR (-1)
int zultra_frame_get_header_size(const unsigned int nFlags, const void *pDictionaryData, const int nDictionarySize);
int zultra_frame_encode_header(unsigned char *pFrameData, const int nMaxFrameDataSize, const unsigned int nFlags, const void *pDictionaryData, const int nDictionarySize);
zultra_frame_checksum_t zultra_frame_init_checksum(const unsigned int nFlags);
zultra_frame_checksum_t zultra_frame_update_checksum(zultra_frame_checksum_t nChecksum, const void *pData, size_t nDataSize, const unsigned int nFlags);
int zultra_frame_get_footer_size(const unsigned int nFlags);
int zultra_frame_encode_footer(unsigned char *pFrameData, const int nMaxFrameDataSize, const zultra_frame_checksum_t nChecksum, long long nOriginalSize, const unsigned int nFlags);

/* dictionary file loader (reference dictionary.h) */
zultra_status_t zultra_dictionary_load(const char *pszDictionaryFilename, void **ppDictionaryData, int *pDictionaryDataSize);
void zultra_dictionary_free(void **ppDictionaryData);

#ifdef __cplusplus
}
#endif
#endif
