"""gen_candidates / eval_posterior / SamdGenerationConfig (reference: samd/utils.py:19-184).

The start token, the draft lookup and the posterior evaluation run on the device: greedy = one `samd_step`
launch and one `samd_verify_compact` launch; greedy=False = the typical-acceptance branch (samd/utils.py:142-184)
as one `samd_verify_sample` launch with a Philox stream per request (include/samd_b200.h states the contract:
Python's random.random() of the reference cannot be reproduced, the statistics are).
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
    seed: int = field(default=0)              # (not in the reference) seed of the device Philox stream when greedy=False

    def __post_init__(self):
        if not self.greedy:
            assert self.temperature >= 1e-5   # samd/utils.py:41


@profile_decorator("gen_candidates")
def gen_candidates(sample_p: torch.Tensor, tree_retrieve_indices: torch.Tensor, draft: DraftModel,
                   samd_config: SamdConfig, gen_config: SamdGenerationConfig, device: torch.device,
                   start_token: Optional[torch.Tensor] = None):
    """samd/utils.py:67-104.  `sample_p` [1, V] -> Candidates(type, tokens [1, n], candidate_tokens, buffers).
    `start_token` (not in the reference): the start token as a 1-element device tensor when the caller already has it -
    the fused verify launch returns the argmax of the row it hands back as sample_p, so the greedy loop skips this pass
    over the vocabulary."""
    # samd/utils.py:85-88: greedy -> argmax of the logits row; sampling -> one draw from the distribution
    if start_token is not None and gen_config.greedy:
        start = start_token.view(-1)
    else:
        start = torch.argmax(sample_p, dim=-1) if gen_config.greedy else torch.multinomial(sample_p, 1).view(-1)
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
_sampling = {}


def _sampling_state(device, seed: int):
    """The process-wide Philox stream of eval_posterior's sampling branch: (seed, offset) as device tensors."""
    key = (str(device), int(seed))
    st = _sampling.get(key)
    if st is None:
        st = _sampling[key] = dict(seed=torch.tensor([int(seed)], dtype=torch.int64, device=device),
                                   offset=torch.zeros(1, dtype=torch.int64, device=device))
    return st


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
    P, D, V = logits.shape
    if not config.greedy:
        # samd/utils.py:142-184: one launch, the P*D gathered rows as nodes with an identity path table
        st = _sampling_state(logits.device, config.seed)
        out = _verifier(logits.device, 1, P * D).verify_sample(
            logits.reshape(1, P * D, V), candidates.reshape(1, P * D).to(torch.int32).contiguous(),
            torch.arange(P * D, dtype=torch.int32, device=logits.device).view(P, D), config.temperature, config.top_p,
            config.top_k, st["seed"], st["offset"], want_sample_p=True)
        return out["best"][0].to(torch.long), out["accept_len"][0].to(torch.long), out["sample_p"]
    lg = logits.reshape(1, P * D, V)
    toks = candidates.reshape(1, P * D).to(torch.int32).contiguous()
    ident = torch.arange(P * D, dtype=torch.int32, device=logits.device).view(P, D)
    out = _verifier(logits.device, 1, P * D).verify(lg, toks, ident, move_kv=False)
    best = out["best"][0].to(torch.long)
    accept_length = out["accept_len"][0].to(torch.long)
    return best, accept_length, logits[best, accept_length - 1].view(1, -1)
