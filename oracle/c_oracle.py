"""ctypes wrapper of oracle/sam_oracle.c (TEST INFRASTRUCTURE ONLY - same rules as samd_oracle.py)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libsam_oracle.so")
_lib = None
i32p, i64p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(_PATH)
        L.so_new.restype = vp
        L.so_new.argtypes = [C.c_int64]
        L.so_free.argtypes = [vp]
        L.so_extend.argtypes = [vp, i32p, C.c_int64]
        L.so_advance.argtypes = [vp, i32p, C.c_int64]
        L.so_reset_cursor.argtypes = [vp]
        L.so_peek.argtypes = [vp, C.c_int32, i32p, i32p]
        L.so_draft_samd.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p]
        L.so_draft_so.restype = C.c_int32
        L.so_draft_so.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, i32p]
        L.so_select_samd.restype = C.c_int32
        L.so_select_samd.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i32p, i32p]
        L.so_build_docs.restype = vp
        L.so_build_docs.argtypes = [i32p, i64p, C.c_int64, C.c_int32]
        L.so_info.argtypes = [vp, i64p]
        L.so_export.argtypes = [vp, i32p, i32p, i32p]
        L.so_run_steps.restype = C.c_int64
        L.so_run_steps.argtypes = [C.POINTER(vp), C.c_int64, i32p, i32p, i32p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32]
        L.so_batch_advance.argtypes = [vp, i32p, i32p, C.c_int64, i32p, C.c_int64]
        L.so_batch_lookup.argtypes = [vp, i32p, i32p, C.c_int64, C.c_int32, i32p, i32p, i32p]
        L.so_min_endpos.restype = C.c_int32
        L.so_min_endpos.argtypes = [vp, C.c_int32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(i32p)


class CSam:
    def __init__(self, max_tokens=None, handle=None):
        self.h = handle if handle is not None else lib().so_new(int(max_tokens))

    @staticmethod
    def build(docs, eos):
        lens = np.array([len(d) for d in docs], dtype=np.int64)
        offs = np.zeros(len(docs) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        flat = np.concatenate([np.asarray(d, dtype=np.int32) for d in docs])
        return CSam(handle=lib().so_build_docs(_p(flat), offs.ctypes.data_as(i64p), len(docs), int(eos)))

    def extend(self, tokens):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        lib().so_extend(self.h, _p(t), len(t))

    def advance(self, tokens):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        lib().so_advance(self.h, _p(t), len(t))

    def reset_cursor(self):
        lib().so_reset_cursor(self.h)

    def peek(self, tok):
        a, b = C.c_int32(), C.c_int32()
        lib().so_peek(self.h, int(tok), C.byref(a), C.byref(b))
        return a.value, b.value

    def draft_samd(self, state, start, n, anchor=True):
        out = np.zeros(n, dtype=np.int32)
        lib().so_draft_samd(self.h, int(state), int(start), int(n), int(anchor), _p(out))
        return out.tolist()

    def draft_so(self, state, matched, start, max_predicts, alpha):
        out = np.zeros(max_predicts, dtype=np.int32)
        k = lib().so_draft_so(self.h, int(state), int(matched), int(start), int(max_predicts), float(alpha), _p(out))
        return out[:k].tolist()

    def select_samd(self, static, start, n, len_bias, len_threshold):
        out = np.zeros(n, dtype=np.int32)
        info = np.zeros(4, dtype=np.int32)
        kind = lib().so_select_samd(self.h, static.h if static is not None else None, int(start), int(n), int(len_bias),
                                    int(len_threshold), _p(out), _p(info))
        return kind, out.tolist(), info.tolist()

    def batch_advance(self, cursors, tokens, counts=None):
        """cursors [n, 2] int32 (state, matched), advanced in place by tokens [n, stride] (counts [n] or all)."""
        assert cursors.dtype == np.int32 and cursors.flags.c_contiguous
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        c = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
        lib().so_batch_advance(self.h, _p(cursors), _p(t), t.shape[1], None if c is None else _p(c), len(cursors))

    def batch_lookup(self, cursors, start, n_predicts):
        """-> (state [n], matched [n], draft [n, n_predicts]) of StaticSAM.lookup + gen_draft for every cursor."""
        n = len(cursors)
        st, ln = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        dr = np.zeros((n, n_predicts), dtype=np.int32)
        lib().so_batch_lookup(self.h, _p(cursors), _p(np.ascontiguousarray(start, dtype=np.int32)), n, int(n_predicts), _p(st), _p(ln), _p(dr))
        return st, ln, dr

    def info(self):
        o = np.zeros(8, dtype=np.int64)
        lib().so_info(self.h, o.ctypes.data_as(i64p))
        return dict(n_states=int(o[0]), n=int(o[1]), n_edges=int(o[2]), n_clones=int(o[3]), cur=int(o[4]), cur_len=int(o[5]))

    def export(self):
        n = self.info()["n_states"]
        a, b, c = (np.zeros(n, dtype=np.int32) for _ in range(3))
        lib().so_export(self.h, _p(a), _p(b), _p(c))
        return a, b, c

    def __del__(self):
        try:
            lib().so_free(self.h)
        except Exception:
            pass
