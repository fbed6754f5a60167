/*
 * frame.c - zlib / gzip framing and host-side checksums of zultra-b200 (API of reference frame.h).
 *
 * Header and trailer bytes follow RFC 1950 / RFC 1952 with the reference's fixed choices (frame.c:387-452:
 * zlib 0x78 + level 3 + FCHECK, optional FDICT + DICTID; gzip MTIME 0, XFL 2, OS 255; frame.c:509-547:
 * Adler-32 big-endian, CRC-32 + ISIZE little-endian).  The checksums of compressed input are computed on the
 * GPU (zb_pipeline.h stage_checksum); the host versions here serve the public zultra_frame_update_checksum
 * entry point and the dictionary id.
 */
#include <pthread.h>
#include "libzultra.h"

#define ADLER_MOD 65521u

static unsigned int adler32_update(unsigned int adler, const unsigned char *p, size_t n) {
   unsigned int a = adler & 0xffffu, b = (adler >> 16) & 0xffffu;
   while (n) {
      size_t k = n > 5552 ? 5552 : n;   /* largest run that cannot overflow 32 bits */
      n -= k;
      while (k--) { a += *p++; b += a; }
      a %= ADLER_MOD; b %= ADLER_MOD;
   }
   return a | (b << 16);
}

static unsigned int g_crc_tab[8][256];
static pthread_once_t g_crc_once = PTHREAD_ONCE_INIT;      /* streams of different host threads may meet here first */
static void crc_init(void) {
   unsigned int i, k;
   for (i = 0; i < 256; i++) {
      unsigned int c = i;
      for (k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? 0xedb88320u : 0u);
      g_crc_tab[0][i] = c;
   }
   for (i = 0; i < 256; i++)
      for (k = 1; k < 8; k++) g_crc_tab[k][i] = (g_crc_tab[k - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[k - 1][i] & 0xff];
}
static unsigned int crc32_update(unsigned int crc, const unsigned char *p, size_t n) {
   pthread_once(&g_crc_once, crc_init);
   crc = ~crc;
   while (n >= 8) {
      unsigned int lo = ((unsigned int)p[0] | ((unsigned int)p[1] << 8) | ((unsigned int)p[2] << 16) | ((unsigned int)p[3] << 24)) ^ crc;
      unsigned int hi = (unsigned int)p[4] | ((unsigned int)p[5] << 8) | ((unsigned int)p[6] << 16) | ((unsigned int)p[7] << 24);
      crc = g_crc_tab[7][lo & 0xff] ^ g_crc_tab[6][(lo >> 8) & 0xff] ^ g_crc_tab[5][(lo >> 16) & 0xff] ^ g_crc_tab[4][lo >> 24] ^
            g_crc_tab[3][hi & 0xff] ^ g_crc_tab[2][(hi >> 8) & 0xff] ^ g_crc_tab[1][(hi >> 16) & 0xff] ^ g_crc_tab[0][hi >> 24];
      p += 8; n -= 8;
   }
   while (n--) crc = g_crc_tab[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
   return ~crc;
}

int zultra_frame_get_header_size(const unsigned int nFlags, const void *pDictionaryData, const int nDictionarySize) {
   if (nFlags & ZULTRA_FLAG_GZIP_FRAMING) return 10;
   if (nFlags & ZULTRA_FLAG_ZLIB_FRAMING) return (pDictionaryData && nDictionarySize) ? 6 : 2;
   return 0;
}

int zultra_frame_encode_header(unsigned char *pFrameData, const int nMaxFrameDataSize, const unsigned int nFlags, const void *pDictionaryData, const int nDictionarySize) {
   if (nFlags & ZULTRA_FLAG_GZIP_FRAMING) {
      static const unsigned char hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 2, 255};
      int i;
      if (nMaxFrameDataSize < 10) return ZULTRA_ENCODE_ERR;
      for (i = 0; i < 10; i++) pFrameData[i] = hdr[i];
      return 10;
   }
   if (nFlags & ZULTRA_FLAG_ZLIB_FRAMING) {
      const int has_dict = (pDictionaryData && nDictionarySize) ? 1 : 0;
      unsigned int cmf = 0x78, flg = 0xc0 | (has_dict ? 0x20 : 0);
      if (nMaxFrameDataSize < 2) return ZULTRA_ENCODE_ERR;
      flg |= (31 - ((cmf << 8) | flg) % 31) & 0x1f;
      pFrameData[0] = (unsigned char)cmf;
      pFrameData[1] = (unsigned char)flg;
      if (has_dict) {
         unsigned int id = adler32_update(1, (const unsigned char *)pDictionaryData, (size_t)nDictionarySize);
         pFrameData[2] = (unsigned char)(id >> 24); pFrameData[3] = (unsigned char)(id >> 16);
         pFrameData[4] = (unsigned char)(id >> 8); pFrameData[5] = (unsigned char)id;
         return 6;
      }
      return 2;
   }
   return 0;
}

zultra_frame_checksum_t zultra_frame_init_checksum(const unsigned int nFlags) {
   if (nFlags & ZULTRA_FLAG_GZIP_FRAMING) return 0;
   if (nFlags & ZULTRA_FLAG_ZLIB_FRAMING) return 1;
   return 0;
}

zultra_frame_checksum_t zultra_frame_update_checksum(zultra_frame_checksum_t nChecksum, const void *pData, size_t nDataSize, const unsigned int nFlags) {
   if (nFlags & ZULTRA_FLAG_GZIP_FRAMING) return crc32_update(nChecksum, (const unsigned char *)pData, nDataSize);
   if (nFlags & ZULTRA_FLAG_ZLIB_FRAMING) return adler32_update(nChecksum, (const unsigned char *)pData, nDataSize);
   return 0;
}

int zultra_frame_get_footer_size(const unsigned int nFlags) {
   if (nFlags & ZULTRA_FLAG_GZIP_FRAMING) return 8;
   if (nFlags & ZULTRA_FLAG_ZLIB_FRAMING) return 4;
   return 0;
}

int zultra_frame_encode_footer(unsigned char *pFrameData, const int nMaxFrameDataSize, const zultra_frame_checksum_t nChecksum, long long nOriginalSize, const unsigned int nFlags) {
   int i;
   if (nFlags & ZULTRA_FLAG_GZIP_FRAMING) {
      if (nMaxFrameDataSize < 8) return ZULTRA_ENCODE_ERR;
      for (i = 0; i < 4; i++) {
         pFrameData[i] = (unsigned char)(nChecksum >> (8 * i));
         pFrameData[4 + i] = (unsigned char)((unsigned long long)nOriginalSize >> (8 * i));
      }
      return 8;
   }
   if (nFlags & ZULTRA_FLAG_ZLIB_FRAMING) {
      if (nMaxFrameDataSize < 4) return ZULTRA_ENCODE_ERR;
      for (i = 0; i < 4; i++) pFrameData[i] = (unsigned char)(nChecksum >> (8 * (3 - i)));
      return 4;
   }
   return 0;
}
