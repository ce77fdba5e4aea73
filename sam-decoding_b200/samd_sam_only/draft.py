"""Drop-in `samd_sam_only.draft.DraftModel` (reference: samd_sam_only/draft.py:15-73): dynamic-SAM
sequence when match_dyn >= match_static - len_bias, otherwise the static SAM's best-first tree."""
from collections import namedtuple
from enum import Enum
from typing import Optional

import torch

from profile_utils import profile_decorator
from samd_b200 import _cabi as K
from samd_b200 import engine as E
from .samd_config import SamdConfig
from .sam import DynSAM, StaticSAM
from .sam.static_sam import tree_buffers_from_parents


class CandidateType(str, Enum):
    sequence = "sequence"
    tree = "tree"


Candidates = namedtuple('Candidates', ['type', 'tokens', 'candidate_tokens', 'buffers_kwargs'])

TOPK = 8


class DraftModel(torch.nn.Module):

    def __init__(self, config: SamdConfig, sam_dyn: DynSAM = None, sam_static: StaticSAM = None, lm=None,
                 dtype: torch.dtype = torch.float16, device: str = "cuda") -> None:
        super().__init__()
        self.config = config
        self.device = device
        self.sam_dyn = sam_dyn if sam_dyn is not None else DynSAM(config.max_predicts, config.alpha, device)
        self.sam_static = sam_static          # None = no corpus: the dynamic automaton always wins
        self.sam_dyn.max_predicts = config.max_predicts
        self.sam_dyn.alpha = config.alpha
        if self.sam_static is not None:
            self.sam_static.max_predicts = config.max_predicts
            self.sam_static.alpha = config.alpha
            self.sam_static.K = config.K
            self.sam_static.device = device
        self.len_bias = config.len_bias
        self._engine: Optional[E.DraftEngine] = None

    def _bind(self) -> E.DraftEngine:
        dyn = self.sam_dyn._ensure(0)
        static = self.sam_static._ensure() if self.sam_static is not None else None
        e = self._engine
        if e is None or e.dyn is not dyn or e.static is not static or e.n_predicts != self.sam_dyn.max_predicts:
            e = E.DraftEngine(dyn, static, K.FLAVOUR_SAM_ONLY, n_predicts=self.sam_dyn.max_predicts, len_bias=self.len_bias,
                              alpha=self.sam_dyn.alpha)
            if static is not None:
                e.static_cursor = self.sam_static._cursor
            e.start = torch.zeros(1, dtype=torch.int32, device=dyn.device)
            self._engine = e
        e.len_bias, e.alpha = self.len_bias, float(self.sam_dyn.alpha)
        return e

    @profile_decorator("DraftModel.reset")
    def reset(self):
        self.sam_dyn.reset()
        if self.sam_static is not None:
            self.sam_static.reset()

    def lookup_device(self, start_token: torch.Tensor) -> E.DraftEngine:
        e = self._bind()
        e.start.copy_(start_token.reshape(1))
        e.step(None, None, e.start)
        if e.static is not None and e.static.with_counts:
            e.tree_draft(e.start, K_top=self.config.K)
        return e

    def results(self, e: E.DraftEngine):
        """(CandidateType, tokens tensor [n] int32 on the device, buffers) of the last lookup; one D2H copy."""
        has_tree = hasattr(e, "tree_n")
        head = [e.out_type, e.draft_len] + ([e.tree_n, e.tree_shape[0]] if has_tree else [])
        vals = torch.cat(head).tolist()
        if vals[0] == K.DRAFT_DYN_SEQ:
            n = vals[1]
            pos = torch.arange(0, n, dtype=torch.long, device=e.draft.device).unsqueeze(0)
            return CandidateType.sequence, e.draft[0, :n], {"seq_position_ids": pos}
        nn, leaves, width = vals[2], vals[3], vals[4]
        buffers = tree_buffers_from_parents(e.tree_parents[0, :nn], e.tree_depth[0, :nn], e.tree_retrieve[0, :leaves, :width])
        return CandidateType.tree, e.tree_tokens[0, :nn], buffers

    @profile_decorator("DraftModel.lookup")
    def lookup(self, start_token: int):
        e = self._bind()
        if e.static is None or not e.static.with_counts:
            # host-facing fast path (no static tree to draft): DraftEngine.host_lookup - one launch that reads the start
            # token from and writes every output to mapped pinned host memory, one stream synchronise
            out_np = e.host_lookup(start_token)
            n = int(out_np[5])                                   # layout of DraftEngine.out_buf at one request
            pos = e.__dict__.get("_seq_pos")
            if pos is None:
                pos = e._seq_pos = torch.arange(0, e.n_predicts, dtype=torch.long, device=e.dyn.device).unsqueeze(0)
            return (CandidateType.sequence, out_np[6:6 + n].tolist(), {"seq_position_ids": pos[:, :n]})
        e.start.fill_(int(start_token))
        kind, toks, buffers = self.results(self.lookup_device(e.start))
        return (kind, toks.tolist(), buffers)

    @profile_decorator("DraftModel.update")
    def update(self, tokens: Optional[torch.Tensor] = None):
        k = int(tokens.numel())
        if not k:
            return
        self.sam_dyn._ensure(k)
        e = self._bind()
        row = tokens.reshape(1, -1)
        if row.dtype != torch.int32 or row.device != e.dyn.device or not row.is_contiguous():
            row = row.to(device=e.dyn.device, dtype=torch.int32).contiguous()
        e.quick_update(row)
        self.sam_dyn._n_tokens += k

    @profile_decorator("DraftModel.prefill_update")
    def prefill_update(self, tokens: Optional[torch.Tensor] = None):
        self.update(tokens)
