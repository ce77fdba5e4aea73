"""Shared machinery of the drop-in `samd` / `samd_sam_only` packages: single-request views of the
device-resident automata that keep the reference's method names, argument meaning and results
 while every operation runs in libsamd_b200.so.

Nothing here falls back to the CPU: constructing an automaton needs the CUDA library, and the
reference's pure-Python state lists only exist as lazily exported *views* (`.states`, `.input_ids`).
"""
from __future__ import annotations

import ctypes as C
import pickle
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi as K
from . import engine as E

FLAT_MAGIC = b"SAMD2B00"


def _dev(device) -> torch.device:
    d = torch.device(device if device is not None else "cuda")
    if d.type != "cuda":
        raise K.SamdError(f"device {d}: the SAM-Decoding hot path runs on CUDA only (no CPU fallback)")
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def _as_i32_row(tokens, device) -> torch.Tensor:
    """[1, k] int32 CUDA tensor from a python list / numpy array / torch tensor."""
    if isinstance(tokens, torch.Tensor):
        t = tokens.reshape(1, -1)
        if t.numel() and not t.dtype.is_floating_point and (bool((t < 0).any()) or bool((t >= 2 ** 31 - 1).any())):
            raise ValueError("token ids must be in [0, 2^31 - 2] (the flat automaton reserves -1 for free edges)")
        return t.to(device=device, dtype=torch.int32).contiguous()
    a = np.asarray(tokens, dtype=np.int64).reshape(1, -1)
    if a.size and (a.min() < 0 or a.max() >= 2 ** 31 - 1):
        raise ValueError("token ids must be in [0, 2^31 - 2] (the flat automaton reserves -1 for free edges)")
    return torch.from_numpy(a.astype(np.int32)).to(device)


class DynSamView:
    """One request's dynamic suffix automaton (reference: samd/sam/dyn_sam.py:8-113)."""

    _FLAVOUR = K.FLAVOUR_SAMD
    INITIAL_CAPACITY = 8192

    def _core_init(self, device="cuda", capacity: Optional[int] = None):
        self._device_arg = device
        self._capacity = int(capacity or self.INITIAL_CAPACITY)
        self._batch: Optional[E.DynSamBatch] = None
        self._n_tokens = 0

    # -- device plumbing ---------------------------------------------------------------
    def _ensure(self, need: int = 0) -> E.DynSamBatch:
        if self._batch is None:
            self._dev = _dev(self._device_arg)
            while self._capacity < self._n_tokens + need:
                self._capacity *= 2
            self._batch = E.DynSamBatch(1, self._capacity, self._dev)
            mk = lambda *s: torch.zeros(*s, dtype=torch.int32, device=self._dev)
            self._tok1, self._idx1, self._len1 = mk(1), mk(1), mk(1)
            self._draft_buf = mk(1, 256)
            self._dlen1 = mk(1)
            self._args = K.StepArgs()
        elif self._n_tokens + need > self._capacity:
            self._grow(self._n_tokens + need)
        return self._batch

    def _grow(self, need: int):
        """Capacity doubling (samd_dyn_grow): same state numbering, cursor and history."""
        while self._capacity < need:
            self._capacity *= 2
        old = self._batch
        self._batch = old.grown(self._capacity)
        old.close()

    def _step(self, tokens: Optional[torch.Tensor], start: Optional[torch.Tensor]):
        a = self._args
        a.dyn, a.stat, a.static_cursor_dev = self._batch.handle, None, None
        if tokens is not None:
            a.tokens_dev, a.token_stride, a.counts_dev = tokens.data_ptr(), tokens.shape[1], None
        else:
            a.tokens_dev, a.token_stride, a.counts_dev = None, 0, None
        a.start_tok_dev = K.ptr(start)
        a.flavour, a.n_predicts, a.len_bias, a.len_threshold, a.alpha = self._FLAVOUR, 1, 0, 0, 1.0
        a.out_type_dev = a.out_match_static_dev = a.out_index_static_dev = a.out_draft_dev = a.out_draft_len_dev = None
        a.out_match_dyn_dev, a.out_index_dyn_dev, a.draft_stride = self._len1.data_ptr(), self._idx1.data_ptr(), 0
        with torch.cuda.device(self._dev):
            K.check(K.lib().samd_step(C.byref(a), K.stream_ptr()), "samd_step")

    # -- reference API -------------------------------------------------------------------
    def reset(self):
        """dyn_sam.py:27-34"""
        if self._batch is not None:
            self._batch.reset()
        self._n_tokens = 0

    def add_tokens(self, tokens):
        """dyn_sam.py:84-88 (match-then-append for every token)."""
        row = _as_i32_row(tokens, _dev(self._device_arg))
        k = row.shape[1]
        if k == 0:
            return
        self._ensure(k)
        self._step(row, None)
        self._n_tokens += k

    def transfer_tokens(self, tokens):
        """dyn_sam.py:90-92 (cursor only)."""
        row = _as_i32_row(tokens, _dev(self._device_arg))
        if row.shape[1] == 0:
            return
        self._ensure(0)
        with torch.cuda.device(self._dev):
            K.check(K.lib().samd_dyn_transfer(self._batch.handle, row.data_ptr(), row.shape[1], None, K.stream_ptr()),
                    "samd_dyn_transfer")

    def lookup(self, token: int) -> Tuple[int, int]:
        """dyn_sam.py:94-97: non-mutating peek -> (state index, match length)."""
        self._ensure(0)
        self._tok1.fill_(int(token))
        self._step(None, self._tok1)
        return int(self._idx1.item()), int(self._len1.item())

    def _gen_draft(self, index: int, match_length: int, start_token: int, n_predicts: int, alpha: float) -> List[int]:
        self._ensure(0)
        if n_predicts > self._draft_buf.shape[1]:
            self._draft_buf = torch.zeros(1, n_predicts, dtype=torch.int32, device=self._dev)
        self._idx1.fill_(int(index))
        self._len1.fill_(int(match_length))
        self._tok1.fill_(int(start_token))
        with torch.cuda.device(self._dev):
            K.check(K.lib().samd_dyn_gen_draft(self._batch.handle, self._idx1.data_ptr(), self._len1.data_ptr(),
                                               self._tok1.data_ptr(), self._FLAVOUR, int(n_predicts), float(alpha),
                                               self._draft_buf.data_ptr(), self._draft_buf.shape[1], self._dlen1.data_ptr(),
                                               K.stream_ptr()), "samd_dyn_gen_draft")
        n = int(self._dlen1.item())
        return self._draft_buf[0, :n].tolist()

    # -- exported views of the device state (synchronise) -------------------------------------
    def _meta(self) -> dict:
        if self._batch is None:
            return dict(n_states=1, last=0, max_length=0, cur_index=0, cur_length=0)
        return self._batch.export(0, with_text=False)

    cur_index = property(lambda self: self._meta()["cur_index"])
    cur_length = property(lambda self: self._meta()["cur_length"])
    max_length = property(lambda self: self._meta()["max_length"])
    last = property(lambda self: self._meta()["last"])

    @property
    def input_ids(self) -> List[int]:
        if self._batch is None:
            return [-1]
        return self._batch.export(0)["text"].tolist()

    def _export_states(self, make_state):
        if self._batch is None:
            return [make_state({}, -1, 0, 0)]
        ex = self._batch.export(0, with_text=False)
        edges = np.zeros((max(1, ex["n_edges"]), 3), dtype=np.int32)
        K.check(K.lib().samd_dyn_export_edges(self._batch.handle, 0, edges.ctypes.data_as(K.c_i32p), len(edges)),
                "samd_dyn_export_edges")
        nxt: List[Dict[int, int]] = [dict() for _ in range(ex["n_states"])]
        for s, t, g in edges[:ex["n_edges"]].tolist():
            nxt[s][t] = g
        return [make_state(nxt[i], int(ex["link"][i]), int(ex["length"][i]), int(ex["min_endpos"][i]))
                for i in range(ex["n_states"])]


class StaticSamView:
    """A static suffix automaton with the reference's query cursor (samd/sam/static_sam.py:8-125,
    samd_sam_only/sam/static_sam.py:22-215).  The automaton lives in HBM (engine.StaticSamDevice);
    an instance unpickled from a reference pickle carries the reference's object graph in
    `__dict__['states']` and is flattened on first use."""

    _WITH_COUNTS = False

    def _core_init(self, device="cuda"):
        self._device_arg = device
        self._sam: Optional[E.StaticSamDevice] = None
        self._pending: List[int] = []              # tokens fed through add_tokens, not yet built
        self._cursor: Optional[torch.Tensor] = None

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def _build(cls, batch_tokens: Sequence[Sequence[int]], eos_token: int, device="cuda", **kw):
        sam = cls(**kw)
        sam._device_arg = device
        sam._sam = E.StaticSamDevice.build(batch_tokens, eos_token, with_counts=cls._WITH_COUNTS, device=_dev(device))
        return sam

    def add_tokens(self, tokens):
        """Incremental construction (static_sam.py:96-100): tokens are buffered on the host and the
        automaton is (re)built lazily over everything fed so far."""
        self._pending.extend(int(t) for t in tokens)
        self.__dict__.pop("states", None)
        if getattr(self, "_sam", None) is not None:
            self._sam.close()
            self._sam = None

    def add_batch_tokens(self, batch_tokens, eos_token: int, verbose: bool = False):
        """static_sam.py:32-36"""
        for tokens in batch_tokens:
            self.add_tokens(tokens)
            if tokens[-1] != eos_token:
                self.add_tokens([eos_token])

    def _ensure(self) -> E.StaticSamDevice:
        d = self.__dict__
        if d.get("_sam") is None:
            dev = _dev(d.get("_device_arg", "cuda"))
            graph = d.get("states")
            if graph is not None and len(graph) > 1:
                d["_sam"] = _flatten_object_graph(self, dev)      # reference pickle
            elif d.get("_pending"):
                toks = d["_pending"]
                # one pseudo-document that already "ends with EOS": nothing gets appended
                d["_sam"] = E.StaticSamDevice.build([toks], toks[-1], with_counts=self._WITH_COUNTS, device=dev)
            else:
                raise K.SamdError("empty StaticSAM: build() it, load_sam() it or add_tokens() first")
        if d.get("_cursor") is None:
            sam = d["_sam"]
            if sam.device is None:
                sam.upload(_dev(d.get("_device_arg", "cuda")))
            d["_cursor"] = sam.new_cursors(1)
            dev = sam.device
            mk = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
            d["_tok1"], d["_idx1"], d["_len1"] = mk(1), mk(1), mk(1)
            d["_draft_buf"] = mk(1, 256)
        return d["_sam"]

    # -- reference API -------------------------------------------------------------------
    def reset(self):
        """static_sam.py:28-30"""
        if self.__dict__.get("_cursor") is not None:
            self._cursor.zero_()

    def transfer_tokens(self, tokens):
        """static_sam.py:102-104"""
        sam = self._ensure()
        row = _as_i32_row(tokens, sam.device)
        if row.shape[1] == 0:
            return
        with torch.cuda.device(sam.device):
            K.check(K.lib().samd_static_walk(sam.handle, self._cursor.data_ptr(), row.data_ptr(), row.shape[1], None, None, 1,
                                             None, None, K.stream_ptr()), "samd_static_walk")

    def lookup(self, token: int) -> Tuple[int, int]:
        """static_sam.py:106-109"""
        sam = self._ensure()
        self._tok1.fill_(int(token))
        with torch.cuda.device(sam.device):
            K.check(K.lib().samd_static_walk(sam.handle, self._cursor.data_ptr(), None, 0, None, self._tok1.data_ptr(), 1,
                                             self._idx1.data_ptr(), self._len1.data_ptr(), K.stream_ptr()), "samd_static_walk")
        return int(self._idx1.item()), int(self._len1.item())

    def _gen_seq_draft(self, index: int, start_token: int, n_predicts: int) -> List[int]:
        sam = self._ensure()
        if n_predicts > self._draft_buf.shape[1]:
            self._draft_buf = torch.zeros(1, n_predicts, dtype=torch.int32, device=sam.device)
        self._idx1.fill_(int(index))
        self._tok1.fill_(int(start_token))
        with torch.cuda.device(sam.device):
            K.check(K.lib().samd_static_gen_draft(sam.handle, self._idx1.data_ptr(), self._tok1.data_ptr(), 1, int(n_predicts),
                                                  self._draft_buf.data_ptr(), self._draft_buf.shape[1], K.stream_ptr()),
                    "samd_static_gen_draft")
        return self._draft_buf[0, :n_predicts].tolist()

    @property
    def cur_index(self):
        c = self.__dict__.get("_cursor")
        return 0 if c is None else int(c[0, 0].item())

    @cur_index.setter
    def cur_index(self, v):            # pickles / load_sam assign these
        self.__dict__["_cur_index_loaded"] = v

    @property
    def cur_length(self):
        c = self.__dict__.get("_cursor")
        return 0 if c is None else int(c[0, 1].item())

    @cur_length.setter
    def cur_length(self, v):
        self.__dict__["_cur_length_loaded"] = v

    def _export_states(self, make_state):
        sam = self._ensure()
        ex = sam.export()
        edges = np.zeros((max(1, sam.n_edges), 3), dtype=np.int32)
        K.check(K.lib().samd_static_export_edges(sam.handle, edges.ctypes.data_as(K.c_i32p), None), "samd_static_export_edges")
        nxt: List[Dict[int, int]] = [dict() for _ in range(sam.n_states)]
        for s, t, g in edges[:sam.n_edges].tolist():
            nxt[s][t] = g
        aux = ex["cnt_endpos"] if self._WITH_COUNTS else ex["min_endpos"]
        return [make_state(nxt[i], int(ex["link"][i]), int(ex["length"][i]), int(aux[i])) for i in range(sam.n_states)]

    def _export_text(self) -> List[int]:
        sam = self._ensure()
        text = np.zeros(sam.n_tokens + 1, dtype=np.int32)
        K.check(K.lib().samd_static_export_edges(sam.handle, None, text.ctypes.data_as(K.c_i32p)), "samd_static_export_edges")
        return text.tolist()


def _flatten_object_graph(view: StaticSamView, dev: torch.device) -> E.StaticSamDevice:
    """Reference pickle (list of SAMState objects with `next` dicts) -> flat arrays -> device."""
    states = view.__dict__["states"]
    n = len(states)
    link = np.fromiter((s.link for s in states), dtype=np.int32, count=n)
    length = np.fromiter((s.length for s in states), dtype=np.int32, count=n)
    has_end = hasattr(states[0], "min_endpos")
    endpos = np.fromiter((s.min_endpos for s in states), dtype=np.int32, count=n) if has_end else None
    count = None if has_end else np.fromiter((s.cnt_endpos for s in states), dtype=np.int32, count=n)
    n_edges = sum(len(s.next) for s in states)
    edges = np.empty((max(1, n_edges), 3), dtype=np.int32)
    k = 0
    for i, s in enumerate(states):
        for t, g in s.next.items():
            edges[k] = (i, t, g)
            k += 1
    ids = view.__dict__.get("input_ids")
    text = np.asarray(ids, dtype=np.int32) if ids is not None and len(ids) > 1 else None
    n_tokens = int(view.__dict__.get("max_length", 0) or (len(text) - 1 if text is not None else length.max()))
    as_p = lambda a: None if a is None else a.ctypes.data_as(K.c_i32p)
    h = K.vp()
    K.check(K.lib().samd_static_from_arrays(n, as_p(link), as_p(length), as_p(endpos), as_p(count), n_edges, as_p(edges),
                                            n_tokens, as_p(text), C.byref(h)), "samd_static_from_arrays")
    sam = E.StaticSamDevice(h, None)
    sam.upload(dev)
    view.__dict__.pop("states", None)            # the object graph is not needed any more
    view.__dict__.pop("input_ids", None)
    return sam


# --------------------------------------------------------------------------------------
# persistence (samd/sam/utils.py:20-37)
# --------------------------------------------------------------------------------------
def dump_sam(path: str, sam: StaticSamView):
    """Flat, mmap-able file (not a pickle of the object graph)."""
    sam._ensure().save(path)


def load_sam(path: str, cls, device="cuda", verbose: bool = True, **kw):
    """Loads either this framework's flat file or a pickle written by the reference's dump_sam."""
    t0 = time.perf_counter()
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic == FLAT_MAGIC:
        sam = cls(**kw)
        sam._device_arg = device
        sam._sam = E.StaticSamDevice.load(path, _dev(device))
    else:
        with open(path, "rb") as f:
            loaded = pickle.load(f)                # resolves <pkg>.sam.static_sam.StaticSAM to this package
        sam = cls(**kw)
        sam._device_arg = device
        for key in ("states", "input_ids", "max_length"):
            if key in vars(loaded):
                sam.__dict__[key] = vars(loaded)[key]
        sam._ensure()
    if verbose:
        print("load sam: {:.2f} s".format(time.perf_counter() - t0))
    return sam
