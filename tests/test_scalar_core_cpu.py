"""CPU test of the one-thread step logic the CUDA kernel is built from (sam-decoding_b200/csrc/sam_scalar.cuh),
compiled for the host by tests/host_emul (test infrastructure, never loaded by the product).  State for state
against the oracle: numbering, links, lengths, min_endpos, every state's edges in insertion order, the cursor, and
the lookups / drafts after every chunk.  Reference being restated: samd/sam/dyn_sam.py:41-113."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import samd_oracle as O
import adversarial as A

HERE = os.path.dirname(os.path.abspath(__file__))
i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", os.path.join(HERE, "host_emul")], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(HERE, "host_emul", "_build", "libsamd_emul.so"))
        L.emul_new.restype = C.c_void_p
        L.emul_new.argtypes = [C.c_int]
        L.emul_free.argtypes = [C.c_void_p]
        L.emul_boundary.argtypes = [C.c_void_p]
        L.emul_extend.argtypes = [C.c_void_p, i32p, C.c_int]
        L.emul_transfer.argtypes = [C.c_void_p, i32p, C.c_int]
        L.emul_lookup.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.emul_draft_samd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, i32p]
        L.emul_info.argtypes = [C.c_void_p, i64p]
        L.emul_export.argtypes = [C.c_void_p, i32p, i32p, i32p]
        L.emul_export_edges.restype = C.c_int64
        L.emul_export_edges.argtypes = [C.c_void_p, i32p, C.c_int64]
        _lib = L
    return _lib


class Emul:
    def __init__(self, cap):
        self.h = lib().emul_new(cap)

    def __del__(self):
        lib().emul_free(self.h)

    def extend(self, toks):
        a = np.ascontiguousarray(toks, dtype=np.int32)
        assert lib().emul_extend(self.h, a.ctypes.data_as(i32p), len(a)) == 0

    def transfer(self, toks):
        a = np.ascontiguousarray(toks, dtype=np.int32)
        lib().emul_transfer(self.h, a.ctypes.data_as(i32p), len(a))

    def boundary(self):
        lib().emul_boundary(self.h)

    def lookup(self, tok):
        i, l = C.c_int32(), C.c_int32()
        lib().emul_lookup(self.h, int(tok), C.byref(i), C.byref(l))
        return i.value, l.value

    def draft(self, index, start, n):
        out = np.zeros(n, dtype=np.int32)
        lib().emul_draft_samd(self.h, int(index), int(start), n, out.ctypes.data_as(i32p))
        return out.tolist()

    def info(self):
        o = np.zeros(8, dtype=np.int64)
        lib().emul_info(self.h, o.ctypes.data_as(i64p))
        return dict(zip(("n_states", "n", "n_edges", "n_clones", "cur", "cur_len", "last", "last_link"), map(int, o)))

    def export(self):
        n = self.info()["n_states"]
        a, b, c = (np.zeros(n, dtype=np.int32) for _ in range(3))
        lib().emul_export(self.h, a.ctypes.data_as(i32p), b.ctypes.data_as(i32p), c.ctypes.data_as(i32p))
        ne = self.info()["n_edges"]
        e = np.zeros((max(ne, 1), 3), dtype=np.int32)
        k = lib().emul_export_edges(self.h, e.ctypes.data_as(i32p), ne)
        return a, b, c, e[:ne], int(k)


def check_equal(em: Emul, ref: O.Automaton):
    inf = em.info()
    assert inf["n_states"] == ref.n_states and inf["n"] == ref.n and inf["n_clones"] == ref.n_clones
    assert (inf["cur"], inf["cur_len"]) == (ref.cur, ref.cur_len)
    assert inf["last"] == ref.tail
    link, length, end, edges, k = em.export()
    assert k == inf["n_edges"] == ref.n_edges
    assert link.tolist() == ref.link and length.tolist() == ref.length and end.tolist() == ref.first_end
    want = [(v, t, g) for v in range(ref.n_states) for t, g in ref.trans[v].items()]      # dict insertion order
    assert [tuple(r) for r in edges.tolist()] == want


def run_stream(stream, sizes, probes, boundary_every=1, n_predicts=16):
    em, ref = Emul(len(stream) + 8), O.Automaton()
    pos = 0
    for step, c in enumerate(sizes):
        chunk = stream[pos:pos + c]
        pos += c
        em.extend(chunk)
        ref.extend(chunk)
        if boundary_every and step % boundary_every == 0:
            em.boundary()
        for tok in probes:
            got = em.lookup(tok)
            assert got == ref.peek(tok), (step, tok)
            assert em.draft(got[0], tok, n_predicts) == O.dyn_draft_samd(ref, got[0], tok, n_predicts), (step, tok)
    check_equal(em, ref)


@pytest.mark.parametrize("name", sorted(A.streams()))
def test_adversarial_streams(name):
    stream = A.streams()[name]
    alphabet = sorted(set(stream))[:4] + [9999]
    for mode, sizes in A.chunkings(len(stream), 5).items():
        run_stream(stream, sizes, alphabet, boundary_every=(0 if mode == "whole" else 2))


def test_long_chains_are_reached():
    """a^200 b really produces a 200-stop chain (so the > SC_CHAIN_MAX fallback is what ran): b's edges = 201."""
    em, ref = Emul(256), O.Automaton()
    em.extend(A.a_k_b(200))
    ref.extend(A.a_k_b(200))
    check_equal(em, ref)
    _, _, _, edges, _ = em.export()
    assert int((edges[:, 1] == 4).sum()) == 201


@pytest.mark.parametrize("vocab,n,seed", [(2, 3000, 1), (3, 3000, 2), (4, 3000, 3), (6, 3000, 4), (40, 6000, 5)])
def test_random_small_alphabets(vocab, n, seed):
    rng = np.random.default_rng(seed)
    stream = rng.integers(3, 3 + vocab, size=n).tolist()
    run_stream(stream, A.chunkings(n, seed)["steps1-8"], list(range(3, 3 + min(vocab, 3))), boundary_every=3)


def test_copy_mix_streams_and_overflow_hubs():
    """The bench's own generator (Zipf head: the root and a few hubs carry long overflow lists) and the clone-light one."""
    from samd_b200 import synth
    for seed, uniform in ((11, False), (12, True)):
        stream = synth.copy_mix(6000, 32000, seed, uniform_fresh=uniform).tolist()
        em, ref = Emul(len(stream) + 8), O.Automaton()
        em.extend(stream[:5000])
        ref.extend(stream[:5000])
        pos = 5000
        rng = np.random.default_rng(seed)
        while pos < len(stream) - 9:
            c = int(rng.integers(1, 9))
            em.extend(stream[pos:pos + c])
            ref.extend(stream[pos:pos + c])
            pos += c
            em.boundary()
            tok = stream[pos]
            got = em.lookup(tok)
            assert got == ref.peek(tok)
            assert em.draft(got[0], tok, 16) == O.dyn_draft_samd(ref, got[0], tok, 16)
        check_equal(em, ref)


def test_cursor_moved_by_transfer_tokens_between_appends():
    """DynSAM.transfer_tokens (dyn_sam.py:90-92) moves the cursor off link(last): the append must not rely on the
    alignment of the two chains then."""
    rng = np.random.default_rng(8)
    stream = rng.integers(3, 6, size=600).tolist()
    em, ref = Emul(700), O.Automaton()
    pos = 0
    while pos < 590:
        c = int(rng.integers(1, 6))
        em.extend(stream[pos:pos + c])
        ref.extend(stream[pos:pos + c])
        pos += c
        if rng.random() < 0.5:
            t = rng.integers(3, 7, size=int(rng.integers(1, 5))).tolist()
            em.transfer(t)
            ref.advance(t)
        if rng.random() < 0.5:
            em.boundary()
        assert em.lookup(3) == ref.peek(3)
    check_equal(em, ref)


@settings(max_examples=150, deadline=None)
@given(st.integers(2, 4).flatmap(lambda v: st.tuples(st.lists(st.integers(3, 2 + v), min_size=1, max_size=300),
                                                     st.integers(0, 2 ** 31 - 1))))
def test_hypothesis_small_alphabets(case):
    stream, seed = case
    run_stream(stream, A.chunkings(len(stream), seed)["steps1-8"], [3, 4], boundary_every=2, n_predicts=7)
