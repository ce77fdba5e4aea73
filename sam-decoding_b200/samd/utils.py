"""gen_candidates / eval_posterior / SamdGenerationConfig (reference: samd/utils.py:19-184).

Greedy decoding only: the start token, the draft lookup and the posterior evaluation run on the
device (one `samd_step` launch and one `samd_verify_compact` launch); the stochastic
typical-acceptance branch (samd/utils.py:142-184) is outside the scope of this framework.
"""
from dataclasses import dataclass, field
from typing import Callable, Optional

import torch

from profile_utils import profile_decorator
from samd_b200 import _cabi as K
from samd_b200 import engine as E
from .samd_config import SamdConfig
from .draft import DraftModel, Candidates, CandidateType


class OptionalTensor:

    def __init__(self, data: Optional[torch.Tensor] = None):
        self.data = data

    def apply(self, fn: Callable) -> 'OptionalTensor':
        return OptionalTensor(None if self.data is None else fn(self.data))


@dataclass
class SamdGenerationConfig:
    max_steps: int = field(default=512)
    max_new_tokens: int = field(default=512)
    max_cache_len: int = field(default=2048)
    greedy: bool = field(default=True)
    temperature: float = field(default=0.0)
    top_p: float = field(default=0.0)
    top_k: int = field(default=0)
    logits_processor: object = field(default=None)

    def __post_init__(self):
        if not self.greedy:
            raise NotImplementedError("only greedy decoding is supported (the sampling branch of the reference, "
                                      "samd/utils.py:142-184, is out of scope)")


@profile_decorator("gen_candidates")
def gen_candidates(sample_p: torch.Tensor, tree_retrieve_indices: torch.Tensor, draft: DraftModel,
                   samd_config: SamdConfig, gen_config: SamdGenerationConfig, device: torch.device):
    """samd/utils.py:67-104.  `sample_p` [1, V] -> Candidates(type, tokens [1, n], candidate_tokens, buffers)."""
    start = torch.argmax(sample_p, dim=-1)                       # stays on the device
    eng = draft.lookup_device(start)
    kind = int(eng.out_type.item())
    if kind != K.DRAFT_TREE_MODEL:
        tokens = eng.draft.to(torch.long)                        # [1, n_predicts]
        return Candidates(CandidateType.sequence, tokens, tokens, {})
    tree, buffers_kwargs = draft.tree_model.gen_draft(int(start.item()))
    tree_retrieve_indices = buffers_kwargs.get("tree_retrieve_indices", tree_retrieve_indices)
    tokens_ext = torch.tensor(tree + [0], dtype=torch.long, device=device)
    candidate_tokens = tokens_ext[tree_retrieve_indices]          # -1 picks the appended 0
    tokens = tokens_ext[:-1].unsqueeze(0)
    return Candidates(CandidateType.tree, tokens, candidate_tokens, buffers_kwargs)


_verifiers = {}


def _verifier(device, batch, nodes) -> E.Verifier:
    key = (str(device), batch >= 1, 0)
    v = _verifiers.get(key)
    if v is None or v.max_nodes < nodes or v.max_batch < batch:
        v = E.Verifier(max(batch, 1), max(nodes, 128), device)
        v.max_nodes, v.max_batch = max(nodes, 128), max(batch, 1)
        _verifiers[key] = v
    return v


@profile_decorator("eval_posterior")
def eval_posterior(logits: torch.Tensor, candidates: torch.Tensor, config: SamdGenerationConfig):
    """samd/utils.py:108-141 (greedy).  `logits` [P, D, V] are the already gathered candidate logits,
    `candidates` [P, D].  Returns (best_candidate, accept_length (accepted + 1), logits[best, accepted]
    as [1, V]).  Runs as one fused launch: the P*D rows are verified with an identity path table."""
    if not config.greedy:
        raise NotImplementedError("only greedy decoding is supported")
    P, D, V = logits.shape
    lg = logits.reshape(1, P * D, V)
    toks = candidates.reshape(1, P * D).to(torch.int32).contiguous()
    ident = torch.arange(P * D, dtype=torch.int32, device=logits.device).view(P, D)
    out = _verifier(logits.device, 1, P * D).verify(lg, toks, ident, move_kv=False)
    best = out["best"][0].to(torch.long)
    accept_length = out["accept_len"][0].to(torch.long)
    return best, accept_length, logits[best, accept_length - 1].view(1, -1)
