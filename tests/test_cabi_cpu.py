"""CPU-only checks of the C ABI: the library loads, exports every symbol include/samd_b200.h declares,
refuses compute without a device, and its HOST halves (static builder, converter, file format) agree
with the reference's golden outputs.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, load, docs_of

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "samd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(samd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from samd_b200 import _cabi as K
    lib = K.lib()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/samd_b200.h but not exported"
    assert set(K.SYMBOLS) <= set(declared) | {"samd_verify_set_chunk"}
    assert set(declared) <= set(K.SYMBOLS), sorted(set(declared) - set(K.SYMBOLS))
    assert lib.samd_abi_version() == 3


def test_no_cpu_fallback():
    import torch
    from samd_b200 import _cabi as K, engine as E
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert K.lib().samd_device_count() == 0
    with pytest.raises(K.SamdError):
        E.DynSamBatch(1, 64)
    with pytest.raises(K.SamdError):
        E.Verifier(1, 8)
    import samd
    with pytest.raises(K.SamdError):
        samd.sam.DynSAM(16, device="cpu").add_tokens([1, 2, 3])


@pytest.mark.parametrize("name", ["tiny", "small", "mid"])
def test_host_builder_matches_reference(name, tmp_path):
    from samd_b200.engine import StaticSamDevice
    z = load("static_sam.npz")
    s = StaticSamDevice.build(docs_of(z, name), int(z[f"{name}/eos"]), with_counts=True, host_only=True)
    e = s.export()
    assert np.array_equal(e["link"], z[f"{name}/link"])
    assert np.array_equal(e["length"], z[f"{name}/length"])
    assert np.array_equal(e["min_endpos"], z[f"{name}/min_endpos"])
    assert np.array_equal(e["cnt_endpos"], z[f"{name}/cnt_endpos"])
    assert np.array_equal(e["topk"][:, :, 0], z[f"{name}/topk_tok"])
    assert np.array_equal(e["topk"][:, :, 1], z[f"{name}/topk_idx"])
    assert s.n_edges == int(z[f"{name}/n_edges"])
    path = str(tmp_path / "sam.bin")
    s.save(path)
    t = StaticSamDevice.load(path, host_only=True)
    e2 = t.export()
    assert (t.n_states, t.n_edges, t.n_tokens) == (s.n_states, s.n_edges, s.n_tokens)
    assert all(np.array_equal(e[k], e2[k]) for k in e)


def test_reference_pickle_converter_host_side():
    """samd_static_from_arrays on the object graph of a pickle written by the reference's dump_sam."""
    import pickle
    import samd.sam.static_sam  # noqa: F401  (the pickle resolves samd.sam.static_sam.StaticSAM to the drop-in class)
    from samd_b200 import _cabi as K
    from samd_b200.engine import StaticSamDevice
    z = load("static_sam.npz")
    with open(os.path.join(GOLDEN, "ref_static_samd.pkl"), "rb") as f:
        ref = pickle.load(f)
    states = ref.__dict__["states"]
    n = len(states)
    link = np.array([s.link for s in states], dtype=np.int32)
    length = np.array([s.length for s in states], dtype=np.int32)
    endpos = np.array([s.min_endpos for s in states], dtype=np.int32)
    edges = np.array([(i, t, g) for i, s in enumerate(states) for t, g in s.next.items()], dtype=np.int32)
    text = np.array(ref.__dict__["input_ids"], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(K.c_i32p)
    h = K.vp()
    K.check(K.lib().samd_static_from_arrays(n, p(link), p(length), p(endpos), None, len(edges), p(edges), len(text) - 1, p(text),
                                            C.byref(h)))
    sam = StaticSamDevice(h, None)
    e = sam.export()
    assert np.array_equal(e["link"], z["small/link"]) and np.array_equal(e["min_endpos"], z["small/min_endpos"])
    out = np.zeros((sam.n_edges, 3), dtype=np.int32)
    K.check(K.lib().samd_static_export_edges(sam.handle, p(out), None))
    assert np.array_equal(out, edges)                     # per-state insertion order survives the round trip


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/samd_b200.h must compile as C99 on its own (no C++-isms, no torch / CUDA
    types), and a C translation unit that references every declared entry point must link against the library."""
    import shutil
    import subprocess
    from samd_b200 import _cabi as K
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    header = os.path.join(REPO, "include", "samd_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", header], check=True)
    names = _declared_symbols()
    src = tmp_path / "link_all.c"
    src.write_text('#include "samd_b200.h"\n#include <stdio.h>\nint main(void) {\n    void *p[] = {%s};\n'
                   '    printf("%%d %%d\\n", (int)(sizeof(p) / sizeof(p[0])), samd_abi_version());\n    return 0;\n}\n'
                   % ", ".join("(void *)%s" % n for n in names))
    exe = tmp_path / "link_all"
    libdir = os.path.dirname(K.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lsamd_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == len(names) and int(out[1]) == K.lib().samd_abi_version()


def test_host_builder_overflow_table_growth():
    """The host builder's overflow table starts at 64 k slots and doubles when half full (sam_static.cu HostSam::grow):
    70 000 single-token documents give the root 70 000 out-edges, i.e. two doublings, with hub states in between; states,
    links, lengths, min_endpos and every state's edges in insertion order must still equal the oracle's."""
    import samd_oracle as O
    from samd_b200 import engine as E
    rng = np.random.default_rng(9)
    docs = [rng.integers(3, 50, size=int(rng.integers(5, 40))).tolist() for _ in range(300)]
    docs += [[t] for t in range(70000)]
    docs += [rng.integers(3, 70000, size=30).tolist() for _ in range(200)]
    st = E.StaticSamDevice.build(docs, 2, with_counts=True, host_only=True)
    ref = O.build_static(docs, 2, count_occurrences=True)
    ex = st.export()
    assert st.n_states == ref.n_states and st.n_edges == ref.n_edges
    assert np.array_equal(ex["link"], np.array(ref.link)) and np.array_equal(ex["length"], np.array(ref.length))
    assert np.array_equal(ex["min_endpos"], np.array(ref.first_end)) and np.array_equal(ex["cnt_endpos"], np.array(ref.occ))
    assert st.n_slots >= 2 * 65536                       # the table really grew
    edges = np.zeros((st.n_edges, 3), dtype=np.int32)
    from samd_b200 import _cabi as K
    K.check(K.lib().samd_static_export_edges(st.handle, edges.ctypes.data_as(K.c_i32p), None))
    want = [(v, t, g) for v in range(ref.n_states) for t, g in ref.trans[v].items()]
    assert [tuple(e) for e in edges.tolist()] == want
