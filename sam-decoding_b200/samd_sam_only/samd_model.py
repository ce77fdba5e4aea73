"""Drop-in `samd_sam_only.SamdModel` (reference: samd_sam_only/samd_model.py:24-333): the same
decode loop as `samd.SamdModel`, with per-step draft shapes (variable-length dynamic sequences,
per-query static trees whose buffers come from the tree-draft kernel) and no tree model."""
from typing import Dict, Optional

import torch

from samd.samd_model import SamdModel as _Base, Outputs  # noqa: F401
from .samd_config import SamdConfig, ForwardType
from .utils import CandidateType, SamdGenerationConfig, gen_candidates
from .draft import DraftModel


class SamdModel(_Base):

    def __init__(self, samd_config: SamdConfig, lm, draft: DraftModel, eos_token_id: int, dtype: torch.dtype, device: str,
                 stop_token_id: Optional[int] = None) -> None:
        super().__init__(samd_config, lm, draft, eos_token_id, dtype, device, stop_token_id)

    def init_seq_position_ids(self):
        return torch.arange(self.samd_config.max_predicts, dtype=torch.long, device=self.device).unsqueeze(0)

    def init_buffers(self):
        """samd_sam_only/samd_model.py:83-94: only the sequence positions are static."""
        self.base_seq_position_ids = self.init_seq_position_ids()
        self.base_tree_attn_mask = self.base_tree_position_ids = self.base_tree_retrieve_indices = None
        self.seq_position_ids = self.base_seq_position_ids
        self.tree_attn_mask = self.tree_position_ids = self.tree_retrieve_indices = None
        self._retrieve_i32 = None

    def update_buffers(self, buffers_kwargs: Dict[str, Optional[torch.Tensor]]):
        self.seq_position_ids = buffers_kwargs.get("seq_position_ids", self.base_seq_position_ids)
        self.tree_attn_mask = buffers_kwargs.get("tree_attn_mask", self.base_tree_attn_mask)
        self.tree_position_ids = buffers_kwargs.get("tree_position_ids", self.base_tree_position_ids)
        self.tree_retrieve_indices = buffers_kwargs.get("tree_retrieve_indices", self.base_tree_retrieve_indices)
        self._retrieve_i32 = None if self.tree_retrieve_indices is None else self.tree_retrieve_indices.to(torch.int32).contiguous()
        self.mask_state.set_state(self.tree_attn_mask)

    def prefill(self, input_ids: torch.Tensor, attention_mask: torch.Tensor):
        """samd_sam_only/samd_model.py:96-115"""
        self.forward_state.forward_type = ForwardType.prefill
        outputs = self.lm(input_ids=input_ids, attention_mask=attention_mask, past_key_values=self.cache)
        self.draft.update(tokens=input_ids.squeeze(0))
        self.cache.set_length()
        return outputs.logits[:, -1]

    def decode(self, sample_p: torch.Tensor, length: int):
        """samd_sam_only/samd_model.py:117-157 with the verify / commit tail fused."""
        cands = gen_candidates(sample_p, self.base_tree_retrieve_indices, self.draft, self.samd_config, self.gen_config, self.device)
        self.update_buffers(cands.buffers_kwargs)
        is_seq = cands.type == CandidateType.sequence
        input_ids = cands.tokens
        if is_seq:
            self.forward_state.forward_type = ForwardType.seq_decode
            outputs = self.lm(input_ids=input_ids, position_ids=self.seq_position_ids + length, past_key_values=self.cache)
        else:
            self.forward_state.forward_type = ForwardType.tree_decode
            outputs = self.lm(input_ids=input_ids, position_ids=self.tree_position_ids + length, past_key_values=self.cache,
                              attention_mask=self._tree_mask_4d(length))
        return self.update_state(input_ids, outputs.logits, is_seq)

    def _draft_update(self, tokens, tree_tokens, tree_logits):
        self.draft.update(tokens=tokens)
