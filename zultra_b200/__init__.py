"""zultra-b200: B200-native implementation of zultra's compression hot path behind the libzultra API.

Python here is a thin ctypes mirror of include/libzultra.h + include/zultra_cuda.h for tests and benchmarks;
the product is zultra_b200/libzultra_b200.so (C host library + CUDA pipeline) and the zultra CLI.
"""
from .api import (ZULTRA_FLAG_DEFLATE_FRAMING, ZULTRA_FLAG_ZLIB_FRAMING, ZULTRA_FLAG_GZIP_FRAMING, ZULTRA_CONTINUE, ZULTRA_FINALIZE,  # noqa: F401
                  ZULTRA_OK, ZULTRA_STREAM_END, lib_path, load, memory_bound, memory_compress, memory_compress_batch, Stream, CudaCtx)
