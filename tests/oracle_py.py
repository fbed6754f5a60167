"""ctypes loader of the plain-C oracle (oracle/libzultra_oracle.so) - test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libzultra_oracle.so")


class _Dump(C.Structure):
    _fields_ = [("max_sub", C.c_int), ("nsub", C.c_int), ("end", C.c_void_p), ("is_dyn", C.c_void_p), ("static_cost", C.c_void_p),
                ("dynamic_cost", C.c_void_p), ("body_bits", C.c_void_p), ("lit_len", C.c_void_p), ("off_len", C.c_void_p), ("best", C.c_void_p), ("best_cap", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "zultra_oracle.c")):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        _lib = C.CDLL(SO)
        _lib.zo_compress.restype = C.c_long
        _lib.zo_compress.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_uint, C.c_uint, C.c_void_p]
    return _lib


def _u8(a):
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8)


def compress(data, flags=0, block=0, dic=None, dump=False):
    d = _u8(data)
    cap = len(d) + len(d) // 4 + 70000
    out = np.zeros(cap, dtype=np.uint8)
    dd = _u8(dic) if dic is not None else None
    st = None
    keep = {}
    if dump:
        ms = 64 * max(1, (len(d) + 32767) // 32768)
        keep = dict(end=np.zeros(ms, np.int32), dyn=np.zeros(ms, np.int32), sc=np.zeros(ms, np.int32), dc=np.zeros(ms, np.int32), bits=np.zeros(ms, np.int32),
                    ll=np.zeros((ms, 288), np.int32), ol=np.zeros((ms, 32), np.int32), best=np.zeros((len(d) + 32768, 2), np.uint16))
        st = _Dump(ms, 0, keep["end"].ctypes.data, keep["dyn"].ctypes.data, keep["sc"].ctypes.data, keep["dc"].ctypes.data, keep["bits"].ctypes.data,
                   keep["ll"].ctypes.data, keep["ol"].ctypes.data, keep["best"].ctypes.data, len(d) + 32768)
    n = lib().zo_compress(d.ctypes.data, len(d), dd.ctypes.data if dd is not None else None, len(dd) if dd is not None else 0, out.ctypes.data, cap, flags, block,
                          C.byref(st) if st is not None else None)
    res = None if n < 0 else out[:n].tobytes()
    if dump:
        k = st.nsub
        return res, {a: b[:k] if a != "best" else b for a, b in keep.items()}
    return res


def sa_lcp(win):
    w = _u8(win)
    out = np.zeros(len(w), dtype=np.uint32)
    lib().zo_window_sa_lcp(w.ctypes.data, len(w), out.ctypes.data)
    return out


def matches(win, hist):
    w = _u8(win)
    out = np.zeros(((len(w) - hist) * 8, 2), dtype=np.uint16)
    lib().zo_window_matches(w.ctypes.data, hist, len(w), out.ctypes.data)
    return out
