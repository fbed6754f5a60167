"""ctypes access to the checkers under oracle/ (test infrastructure only)."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libzultra_ref.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libzb_emu.so")

def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None

class Ref:
    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        self.lib.zultra_memory_compress.restype = C.c_size_t
        self.lib.zultra_memory_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint]
        self.lib.zultra_memory_bound.restype = C.c_size_t
        self.lib.zultra_memory_bound.argtypes = [C.c_size_t, C.c_uint, C.c_uint]
        self.lib.refh_compress_with_dict.restype = C.c_long

    def compress(self, data, flags=0, block=0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cap = self.lib.zultra_memory_bound(len(data), flags, block)
        out = np.zeros(cap, dtype=np.uint8)
        n = self.lib.zultra_memory_compress(_p(data), len(data), _p(out), cap, flags, block)
        if n == C.c_size_t(-1).value:
            return None
        return out[:n].tobytes()

    def compress_dict(self, data, dic, flags=1, block=0):
        data = np.ascontiguousarray(data, dtype=np.uint8); dic = np.ascontiguousarray(dic, dtype=np.uint8)
        cap = len(data) + len(data) // 8 + 1024
        out = np.zeros(cap, dtype=np.uint8)
        n = self.lib.refh_compress_with_dict(_p(data), C.c_long(len(data)), _p(dic), len(dic), _p(out), C.c_long(cap), flags, block)
        return None if n < 0 else out[:n].tobytes()

    def sa_lcp(self, win):
        win = np.ascontiguousarray(win, dtype=np.uint8)
        out = np.zeros(len(win), dtype=np.uint32)
        r = self.lib.refh_window_sa_lcp(_p(win), len(win), _p(out))
        assert r == len(win)
        return out

    def matches(self, win, hist):
        win = np.ascontiguousarray(win, dtype=np.uint8)
        out = np.zeros(((len(win) - hist) * 8, 2), dtype=np.uint16)
        r = self.lib.refh_window_matches(_p(win), hist, len(win), _p(out))
        assert r == 0
        return out

    def block_stages(self, win, hist):
        win = np.ascontiguousarray(win, dtype=np.uint8)
        n = len(win)
        split = np.zeros(64, dtype=np.int32); dyn = np.zeros(64, dtype=np.int32); sc = np.zeros(64, dtype=np.int32); dc = np.zeros(64, dtype=np.int32)
        ll = np.zeros((64, 288), dtype=np.int32); ol = np.zeros((64, 32), dtype=np.int32); bits = np.zeros(64, dtype=np.int32)
        best = np.zeros((n, 2), dtype=np.uint16); body = np.zeros(n + n // 2 + 65536, dtype=np.uint8); boff = np.zeros(65, dtype=np.int32)
        k = self.lib.refh_block_stages(_p(win), hist, n, _p(split), _p(dyn), _p(sc), _p(dc), _p(ll), _p(ol), _p(bits), _p(best), _p(body), len(body), _p(boff))
        assert k > 0, k
        return dict(n=k, split=split[:k], dyn=dyn[:k], sc=sc[:k], dc=dc[:k], ll=ll[:k], ol=ol[:k], bits=bits[:k], best=best, body=body, boff=boff[:k + 1])

class Emu:
    def __init__(self):
        self.lib = C.CDLL(EMU_SO)
        self.lib.emu_compress.restype = C.c_long

    def compress(self, data, hist=None, block=1 << 20, finalize=1, in_bits=0, tile=0, dump=False):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        hist = np.zeros(0, dtype=np.uint8) if hist is None else np.ascontiguousarray(hist, dtype=np.uint8)
        cap = len(data) + len(data) // 4 + 4096
        out = np.zeros(cap, dtype=np.uint8)
        bits = C.c_ulonglong(0); crc = C.c_uint(0)
        nwin = max(1, (len(data) + block - 1) // block)
        P = len(data) + len(hist) + (nwin - 1) * 32768
        d = {}
        if dump:
            d = dict(sa_lcp=np.zeros(P, dtype=np.uint32), match=np.zeros((P * 8, 2), dtype=np.uint16), nsub=C.c_int(0),
                     sub=np.zeros((64 * nwin, 8), dtype=np.int32), ll=np.zeros((64 * nwin, 288), dtype=np.int32), ol=np.zeros((64 * nwin, 32), dtype=np.int32),
                     best=np.zeros((P, 2), dtype=np.uint16))
        n = self.lib.emu_compress(_p(data), C.c_long(len(data)), _p(hist) if len(hist) else None, len(hist), C.c_uint(block), finalize, in_bits,
                                  _p(out), C.c_long(cap), C.byref(bits), C.c_uint(tile),
                                  _p(d.get("sa_lcp")), _p(d.get("match")), C.byref(d["nsub"]) if dump else None, _p(d.get("sub")),
                                  _p(d.get("ll")), _p(d.get("ol")), _p(d.get("best")), C.byref(crc))
        d["crc"] = crc.value
        assert n >= 0, n
        if dump:
            k = d["nsub"].value
            d["nsub"] = k; d["sub"] = d["sub"][:k]; d["ll"] = d["ll"][:k]; d["ol"] = d["ol"][:k]
        return out[:n].tobytes(), bits.value, d

    def shard_prepare(self, buf, hist, n, block=1 << 20, finalize=0):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self._keep = buf
        bits = (C.c_ulonglong * 8)()
        rc = self.lib.emu_shard_prepare(_p(buf), hist, C.c_long(n), C.c_uint(block), finalize, bits)
        assert rc == 0
        return [int(b) for b in bits]

    def shard_emit(self, in_bits, cap):
        out = np.zeros(cap, dtype=np.uint8)
        bits = C.c_ulonglong(0)
        self.lib.emu_shard_emit.restype = C.c_long
        n = self.lib.emu_shard_emit(in_bits, _p(out), C.c_long(cap), C.byref(bits))
        assert n >= 0
        return out[:n].tobytes(), bits.value

    def chunks_prepare(self, buf, chunks, block=1 << 20, flags=0):
        """chunks: [(offset of the chunk's history start in buf, history bytes, chunk bytes, finalize)] -> (maps [n][8], checksums)."""
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        self._keep = buf
        n = len(chunks)
        off = (C.c_size_t * n)(*[c[0] for c in chunks]); hist = (C.c_int * n)(*[c[1] for c in chunks])
        ln = (C.c_size_t * n)(*[c[2] for c in chunks]); fin = (C.c_int * n)(*[c[3] for c in chunks])
        cks = (C.c_uint * n)(); maps = (C.c_ulonglong * (8 * n))()
        rc = self.lib.emu_chunks_prepare(_p(buf), n, off, hist, ln, fin, C.c_uint(block), C.c_uint(flags), cks, maps)
        assert rc == 0
        return [[int(maps[8 * i + p]) for p in range(8)] for i in range(n)], [int(c) for c in cks]

    def chunks_emit(self, in_bits, cap):
        n = len(in_bits)
        ib = (C.c_uint * n)(*in_bits); off = (C.c_size_t * n)(); bits = (C.c_ulonglong * n)()
        out = np.zeros(cap, dtype=np.uint8)
        self.lib.emu_chunks_emit.restype = C.c_long
        r = self.lib.emu_chunks_emit(ib, _p(out), C.c_long(cap), off, bits)
        assert r >= 0, r
        return [out[off[i]:off[i] + (bits[i] + 7) // 8].tobytes() for i in range(n)], [int(b) for b in bits]
