"""gen_candidates / eval_posterior / SamdGenerationConfig of the SAM-only variant
(reference: samd_sam_only/utils.py, identical to samd/utils.py apart from the draft types)."""
import torch

from profile_utils import profile_decorator
from samd.utils import OptionalTensor, SamdGenerationConfig, eval_posterior  # noqa: F401
from .samd_config import SamdConfig
from .draft import DraftModel, Candidates, CandidateType


@profile_decorator("gen_candidates")
def gen_candidates(sample_p: torch.Tensor, tree_retrieve_indices: torch.Tensor, draft: DraftModel, samd_config: SamdConfig,
                   gen_config: SamdGenerationConfig, device: torch.device):
    """samd_sam_only/utils.py:67-104: start token, lookup, candidate tensors - no host round trip
    besides the one small copy that tells the sequence / tree shapes."""
    start = torch.argmax(sample_p, dim=-1)
    kind, toks, buffers_kwargs = draft.results(draft.lookup_device(start))
    tokens = toks.to(torch.long).unsqueeze(0)
    if kind == CandidateType.sequence:
        return Candidates(kind, tokens, tokens, buffers_kwargs)
    ri = buffers_kwargs.get("tree_retrieve_indices", tree_retrieve_indices)
    tokens_ext = torch.cat([tokens[0], tokens.new_zeros(1)])
    return Candidates(kind, tokens, tokens_ext[ri], buffers_kwargs)
