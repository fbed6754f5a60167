"""Host logic of the multi-GPU path: block-range sharding and bit-offset stitching.

A shard's bitstream depends on the shards before it only through the entering bit phase (0..7): the stored-block test
of libzultra.c:345-347 is made on byte deltas of the bit writer.  Every shard therefore reports its size in bits for all
8 phases; the maps compose left to right into each shard's true phase and absolute bit offset.
"""
import numpy as np


def plan_shards(n, block, world):
    """Contiguous max-block ranges: [(lo, hi)] byte ranges per rank (empty ranges allowed)."""
    nblocks = (n + block - 1) // block
    per = (nblocks + world - 1) // world
    return [(min(n, r * per * block), min(n, (r + 1) * per * block)) for r in range(world)]


def chunk_blocks(nblocks, world):
    """Max-blocks per chunk (same rule as multi_chunk_blocks, zb_capi.cu): at least ~4 chunks per device, at most 8 blocks."""
    return max(1, min(8, nblocks // (world * 4)))


def plan_chunks(n, block, world, g=None):
    """Chunks of g max-blocks dealt round-robin to the ranks: [(lo, hi, rank)] in stream order.  Every rank gets a sample of
    the whole stream, so a stream whose cost per byte varies along its length still loads all ranks evenly."""
    nblocks = (n + block - 1) // block
    g = g or chunk_blocks(nblocks, world)
    nchunk = (nblocks + g - 1) // g
    return [(j * g * block, min(n, (j + 1) * g * block), j % world) for j in range(nchunk)]


def compose(maps):
    """maps[r][p] = bits of shard r when entered at phase p, INCLUDING the p pending bits.  Returns (offsets, nbits, total):
    absolute bit offset and produced bits of every shard."""
    offs, nbits, pos = [], [], 0
    for m in maps:
        ph = pos & 7
        offs.append(pos)
        produced = int(m[ph]) - ph
        nbits.append(produced)
        pos += produced
    return offs, nbits, pos


def merge(buffers, offs, nbits, total_bits):
    """OR the shard buffers (each starts at the byte holding its first bit) into one stream."""
    out = np.zeros((total_bits + 7) // 8 + 1, dtype=np.uint8)
    for buf, off, nb in zip(buffers, offs, nbits):
        if nb == 0:
            continue
        start = off >> 3
        ln = ((off & 7) + nb + 7) // 8
        out[start:start + ln] |= np.frombuffer(buf, dtype=np.uint8)[:ln]
    return out[: (total_bits + 7) // 8].tobytes()
