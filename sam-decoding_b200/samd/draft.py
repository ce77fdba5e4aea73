"""Drop-in `samd.draft.DraftModel` (reference: samd/draft.py:16-79).

lookup() and update() are each ONE launch of the fused per-step kernel (`samd_step`): the dynamic
and static automata are probed together, the draft source is selected on the device and the draft
tokens are read from HBM; accepted tokens never leave the device on the update path."""
from collections import namedtuple
from enum import Enum
from typing import Optional

import torch

from profile_utils import profile_decorator, profile_lookup_decorator  # noqa: F401
from samd_b200 import _cabi as K
from samd_b200 import engine as E
from .samd_config import SamdConfig
from .sam import DynSAM, StaticSAM, NullStaticSAM
from .tree_model import TreeModel, tree_model_cls


class CandidateType(str, Enum):
    sequence = "sequence"
    tree = "tree"


Candidates = namedtuple('Candidates', ['type', 'tokens', 'candidate_tokens', 'buffers_kwargs'])

TOPK = 8


class DraftModel(torch.nn.Module):

    def __init__(self, config: SamdConfig, sam_dyn: DynSAM = None, sam_static: StaticSAM = None,
                 tree_model: TreeModel = None, lm=None, dtype: torch.dtype = torch.float16, device: str = "cuda") -> None:
        super().__init__()
        self.config = config
        self.device = device
        self.sam_dyn = sam_dyn if sam_dyn is not None else DynSAM(config.n_predicts, device)
        self.sam_static = sam_static if sam_static is not None else NullStaticSAM(config.n_predicts, device)
        self.tree_model = tree_model if tree_model is not None else tree_model_cls[config.tree_method](config, lm, dtype, device)
        self.sam_dyn.n_predicts = config.n_predicts
        self.sam_static.n_predicts = config.n_predicts
        self.len_bias = config.len_bias
        self.len_threshold = config.len_threshold
        self._engine: Optional[E.DraftEngine] = None

    # the engine is (re)bound lazily because DynSAM may grow its arena (new handle)
    def _bind(self) -> E.DraftEngine:
        dyn = self.sam_dyn._ensure(0)
        static = None if isinstance(self.sam_static, NullStaticSAM) else self.sam_static._ensure()
        e = self._engine
        if e is None or e.dyn is not dyn or e.static is not static or e.n_predicts != self.sam_dyn.n_predicts:
            e = E.DraftEngine(dyn, static, K.FLAVOUR_SAMD, n_predicts=self.sam_dyn.n_predicts, len_bias=self.len_bias,
                              len_threshold=self.len_threshold)
            if static is not None:
                e.static_cursor = self.sam_static._cursor          # one cursor, shared with StaticSAM's own API
            e.start = torch.zeros(1, dtype=torch.int32, device=dyn.device)
            self._engine = e
        e.len_bias, e.len_threshold = self.len_bias, self.len_threshold
        return e

    def reset(self):
        self.sam_dyn.reset()
        self.sam_static.reset()
        self.tree_model.reset()

    def lookup_device(self, start_token: torch.Tensor):
        """Device-side lookup: `start_token` is a 1-element CUDA tensor; results stay in the engine."""
        e = self._bind()
        e.start.copy_(start_token.reshape(1))
        e.step(None, None, e.start)
        return e

    def lookup(self, start_token: int):
        e = self._bind()
        out_np = e.host_lookup(start_token)       # one launch, mapped pinned host I/O, one stream synchronise
        if int(out_np[0]) != K.DRAFT_TREE_MODEL:
            return (CandidateType.sequence, out_np[6:6 + e.n_predicts].tolist(), {})
        return (CandidateType.tree,) + tuple(self.tree_model.gen_draft(int(start_token)))

    def update(self, tokens: Optional[torch.Tensor] = None, last_hidden_states: Optional[torch.Tensor] = None,
               tree_tokens: Optional[torch.Tensor] = None, tree_logits: Optional[torch.Tensor] = None):
        e = self._bind()
        k = int(tokens.numel())
        if k:
            self.sam_dyn._ensure(k)
            e = self._bind()
            row = tokens.reshape(1, -1)
            if row.dtype != torch.int32 or row.device != e.dyn.device or not row.is_contiguous():
                row = row.to(device=e.dyn.device, dtype=torch.int32).contiguous()
            e.quick_update(row)                                                 # dyn.add_tokens + static.transfer_tokens
            self.sam_dyn._n_tokens += k
        self.tree_model.update(tokens=tokens, last_hidden_states=last_hidden_states, tree_tokens=tree_tokens,
                               tree_logits=tree_logits)
