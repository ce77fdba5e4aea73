"""Multi-GPU partitioning of the path.

* Requests shard across GPUs with NO communication (each rank owns its requests' dynamic automata,
  verification batches and KV cache) - bench.py --gpus N runs exactly that.
* A static corpus too large for one GPU is split by DOCUMENT into `world` contiguous ranges; rank g
  holds the automaton of range g and a cursor per query.  Per step every rank computes a packed key
      key = (match_len << 32) | (0xFFFFFFFF - (shard_offset + min_endpos))      (0 = no match)
  for every query, one all-reduce-max over the 64-bit keys picks "longest match, then earliest global
  occurrence" - the single-automaton tie-break - and the draft is read from the replicated corpus
  token array at the winning position.  One collective of Q x 8 bytes per step.

The key arithmetic and the sharding plan are plain host code (tested with gloo on CPU); the lookups
run in libsamd_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

KEY_LOW = 0xFFFFFFFF


def shard_documents(doc_lengths: Sequence[int], ends_with_eos: Sequence[bool], world: int) -> List[Tuple[int, int, int]]:
    """Contiguous document ranges balanced by token count.  Returns per shard (first_doc, last_doc_excl,
    token_offset) where token_offset = number of corpus tokens (documents + appended EOS) before it."""
    tok = np.asarray(doc_lengths, dtype=np.int64) + (~np.asarray(ends_with_eos, dtype=bool)).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(tok)])
    total = int(csum[-1])
    cuts = [0]
    for g in range(1, world):
        target = total * g // world
        d = int(np.searchsorted(csum, target, side="left"))
        cuts.append(min(max(d, cuts[-1]), len(tok)))
    cuts.append(len(tok))
    return [(cuts[g], cuts[g + 1], int(csum[cuts[g]])) for g in range(world)]


def pack_key(match_len, global_endpos):
    """numpy / torch int64: (len << 32) | (0xFFFFFFFF - endpos); 0 where len == 0."""
    if isinstance(match_len, torch.Tensor):
        key = (match_len.to(torch.int64) << 32) | (KEY_LOW - global_endpos.to(torch.int64))
        return torch.where(match_len > 0, key, torch.zeros_like(key))
    match_len = np.asarray(match_len, dtype=np.int64)
    key = (match_len << 32) | (KEY_LOW - np.asarray(global_endpos, dtype=np.int64))
    return np.where(match_len > 0, key, 0)


def unpack_key(key):
    """-> (match_len, global_endpos); endpos 0 where there was no match."""
    if isinstance(key, torch.Tensor):
        length = key >> 32
        end = torch.where(length > 0, KEY_LOW - (key & KEY_LOW), torch.zeros_like(key))
        return length, end
    key = np.asarray(key, dtype=np.int64)
    length = key >> 32
    return length, np.where(length > 0, KEY_LOW - (key & KEY_LOW), 0)


def reduce_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce-max of the packed keys (NCCL over NVLink on GPUs, gloo in the CPU tests).
    Keys are non-negative int64, so the signed max equals the unsigned 64-bit max of the design."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys


def flatten_corpus(docs: Sequence[Sequence[int]], eos: int) -> np.ndarray:
    """The 1-based global token array the single automaton would index (text[0] = -1)."""
    parts = [np.array([-1], dtype=np.int32)]
    for d in docs:
        d = np.asarray(d, dtype=np.int32)
        parts.append(d)
        if d[-1] != eos:
            parts.append(np.array([eos], dtype=np.int32))
    return np.concatenate(parts)


class ShardedStaticSam:
    """Rank-local shard of a document-sharded static automaton + the replicated corpus tokens."""

    def __init__(self, docs: Sequence[Sequence[int]], eos: int, rank: int, world: int, n_queries: int,
                 device: Optional[torch.device] = None, corpus: Optional[np.ndarray] = None):
        from . import engine as E
        self.rank, self.world = rank, world
        lens = [len(d) for d in docs]
        ends = [int(d[-1]) == eos for d in docs]
        self.plan = shard_documents(lens, ends, world)
        lo, hi, self.offset = self.plan[rank]
        self.sam = E.StaticSamDevice.build(docs[lo:hi], eos, with_counts=False, device=device)
        self.device = self.sam.device
        flat = corpus if corpus is not None else flatten_corpus(docs, eos)
        self.n_corpus = int(len(flat) - 1)
        self.corpus = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.int32)).to(self.device)
        self.cursor = self.sam.new_cursors(n_queries)
        self.keys = torch.zeros(n_queries, dtype=torch.int64, device=self.device)
        self.n_queries = n_queries
        self._xchg = None

    @classmethod
    def from_parts(cls, sam, offset: int, corpus: torch.Tensor, rank: int, world: int, n_queries: int) -> "ShardedStaticSam":
        """A shard whose automaton was built by this rank from ITS documents only (no rank ever holds the whole corpus
        on the host): `sam` = the uploaded StaticSamDevice of the shard, `offset` = corpus tokens in front of it,
        `corpus` = the replicated 1-based token array on this rank's device (corpus[0] = -1)."""
        self = cls.__new__(cls)
        self.rank, self.world, self.plan = rank, world, None
        self.offset, self.sam, self.device = int(offset), sam, sam.device
        self.n_corpus = int(corpus.numel() - 1)
        self.corpus = corpus
        self.cursor = sam.new_cursors(n_queries)
        self.keys = torch.zeros(n_queries, dtype=torch.int64, device=self.device)
        self.n_queries = n_queries
        self._xchg = None
        return self

    def connect_peers(self, group=None) -> bool:
        """Set up the NVLink peer exchange (one process per GPU of one node): every rank allocates its exchange
        buffer, the CUDA IPC handles are all-gathered through torch.distributed, and every rank maps its peers'
        buffers.  Collective; returns True when EVERY rank succeeded (only then may lookup_draft(p2p=True) be used -
        a rank that went ahead alone would wait for peers that never signal)."""
        import ctypes as C
        import torch.distributed as dist
        from . import _cabi as K
        h, ok = K.vp(), 1
        mine = (C.c_ubyte * 64)()
        try:
            with torch.cuda.device(self.device):
                K.check(K.lib().samd_xchg_create(self.rank, self.world, self.n_queries, C.byref(h)), "samd_xchg_create")
                K.check(K.lib().samd_xchg_export(h, C.cast(mine, K.vp)), "samd_xchg_export")
        except K.SamdError:
            ok = 0
        mine_t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        all_t = [torch.empty_like(mine_t) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(all_t, mine_t, group=group)
        else:
            all_t = [mine_t]
        if ok:
            flat = torch.stack(all_t).cpu().numpy().tobytes()
            buf = (C.c_ubyte * len(flat)).from_buffer_copy(flat)
            try:
                with torch.cuda.device(self.device):
                    K.check(K.lib().samd_xchg_connect(h, C.cast(buf, K.vp)), "samd_xchg_connect")
            except K.SamdError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        if self.world > 1:                              # also the barrier: nobody writes into an unmapped buffer
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self._xchg = h if int(flag.item()) == 1 else None
        return self._xchg is not None

    def peers_ok(self) -> bool:
        from . import _cabi as K
        return self._xchg is not None and K.lib().samd_xchg_status(self._xchg) == 0

    def reset(self):
        self.cursor.zero_()

    def advance(self, tokens: torch.Tensor, counts: Optional[torch.Tensor] = None):
        """StaticSAM.transfer_tokens on this shard for every query (replicated G x across shards)."""
        from . import _cabi as K
        with torch.cuda.device(self.device):
            K.check(K.lib().samd_static_walk(self.sam.handle, self.cursor.data_ptr(), tokens.data_ptr(), tokens.shape[1],
                                             K.ptr(counts), None, self.n_queries, None, None, K.stream_ptr()), "samd_static_walk")

    def local_keys(self, start_tok: torch.Tensor) -> torch.Tensor:
        from . import _cabi as K
        with torch.cuda.device(self.device):
            K.check(K.lib().samd_static_lookup_keys(self.sam.handle, self.cursor.data_ptr(), start_tok.data_ptr(), self.n_queries,
                                                    self.offset, self.keys.data_ptr(), K.stream_ptr()), "samd_static_lookup_keys")
        return self.keys

    def draft(self, keys: torch.Tensor, start_tok: torch.Tensor, n_predicts: int):
        from . import _cabi as K
        match = torch.empty(self.n_queries, dtype=torch.int32, device=self.device)
        draft = torch.empty(self.n_queries, n_predicts, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            K.check(K.lib().samd_draft_from_keys(keys.data_ptr(), self.corpus.data_ptr(), self.n_corpus, start_tok.data_ptr(),
                                                 self.n_queries, n_predicts, match.data_ptr(), draft.data_ptr(), n_predicts,
                                                 K.stream_ptr()), "samd_draft_from_keys")
        return match, draft

    def lookup_draft(self, start_tok: torch.Tensor, n_predicts: int, group=None, p2p: bool = False, out=None,
                     tokens: Optional[torch.Tensor] = None, counts: Optional[torch.Tensor] = None):
        """local keys -> max over the shards -> draft from the replicated corpus.  p2p=False: one NCCL all-reduce-max
        between two kernels.  p2p=True (after connect_peers()): the look-up kernel max-reduces into every rank's
        buffer over NVLink itself and the draft kernel waits for the shards' keys - two launches, no collective.
        `tokens` / `counts`: advance the cursors first (StaticSAM.transfer_tokens; the same launch when p2p)."""
        if not p2p:
            if tokens is not None:
                self.advance(tokens, counts)
            keys = reduce_keys(self.local_keys(start_tok), group)
            return self.draft(keys, start_tok, n_predicts)
        from . import _cabi as K
        if self._xchg is None:
            raise K.SamdError("lookup_draft(p2p=True) before connect_peers()")
        if out is None:
            out = (torch.empty(self.n_queries, dtype=torch.int32, device=self.device),
                   torch.empty(self.n_queries, n_predicts, dtype=torch.int32, device=self.device))
        match, draft = out
        with torch.cuda.device(self.device):
            K.check(K.lib().samd_static_lookup_exchange(self.sam.handle, self.cursor.data_ptr(), K.ptr(tokens),
                                                        tokens.shape[1] if tokens is not None else 0, K.ptr(counts),
                                                        start_tok.data_ptr(), self.offset, self._xchg, K.stream_ptr()),
                    "samd_static_lookup_exchange")
            K.check(K.lib().samd_draft_from_exchange(self._xchg, self.corpus.data_ptr(), self.n_corpus, start_tok.data_ptr(),
                                                     n_predicts, match.data_ptr(), draft.data_ptr(), n_predicts, K.stream_ptr()),
                    "samd_draft_from_exchange")
        return match, draft
