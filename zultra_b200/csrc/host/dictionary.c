/*
 * dictionary.c - preset dictionary file loader (API of reference dictionary.h; behaviour of dictionary.c:49-104:
 * a missing file is ZULTRA_ERROR_DICTIONARY, only the last 32 KiB of the file are kept).
 */
#include <stdio.h>
#include <stdlib.h>
#include "libzultra.h"

zultra_status_t zultra_dictionary_load(const char *pszDictionaryFilename, void **ppDictionaryData, int *pDictionaryDataSize) {
   unsigned char *data = NULL;
   int size = 0;
   if (pszDictionaryFilename) {
      FILE *f;
      long total;
      data = (unsigned char *)malloc(HISTORY_SIZE);
      if (!data) return ZULTRA_ERROR_MEMORY;
      f = fopen(pszDictionaryFilename, "rb");
      if (!f) { free(data); return ZULTRA_ERROR_DICTIONARY; }
      fseek(f, 0, SEEK_END);
      total = ftell(f);
      fseek(f, total > HISTORY_SIZE ? total - HISTORY_SIZE : 0, SEEK_SET);
      size = (int)fread(data, 1, HISTORY_SIZE, f);
      if (size < 0) size = 0;
      fclose(f);
   }
   *ppDictionaryData = data;
   *pDictionaryDataSize = size;
   return ZULTRA_OK;
}

void zultra_dictionary_free(void **ppDictionaryData) {
   if (ppDictionaryData && *ppDictionaryData) { free(*ppDictionaryData); *ppDictionaryData = NULL; }
}
