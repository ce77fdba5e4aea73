"""Drop-in `samd.SamdModel` (reference: samd/samd_model.py:23-322): the speculative decode loop.

prefill -> { gen_candidates -> LM forward -> fused verify + KV compaction -> draft update }*

Differences from the reference are confined to HOW each step runs:
  * no monkey patching of the LM: the tree attention pattern is passed as an explicit 4-D additive
    mask (the reference spliced it inside a patched `_update_causal_mask`,
    samd/model_patch/llama.py:94-96), so any transformers >= 5 causal LM works;
  * the logits gather + eval_posterior + update_state + select_indices chain
    (samd/samd_model.py:159-211) is ONE kernel launch on the [1, T, V] tree logits;
  * accepted tokens feed the automata device-to-device; one small device->host copy per step
    brings the accepted ids to the Python loop (EOS / stop handling, output list).
"""
from collections import namedtuple
from typing import Dict, Optional

import torch
import torch.nn as nn

from profile_utils import profile_decorator, profile_accept_length  # noqa: F401
from samd_b200 import engine as E
from .samd_config import SamdConfig, ForwardState, ForwardType, MaskState
from .utils import OptionalTensor, CandidateType, SamdGenerationConfig, gen_candidates, eval_posterior  # noqa: F401
from .cache import SamdStaticCache
from .draft import DraftModel

Outputs = namedtuple('Outputs', ['output_ids', 'decode_tokens', 'decode_steps', 'accepet_length_per_step'])


class SamdModel(nn.Module):

    def __init__(self, samd_config: SamdConfig, lm, draft: DraftModel, eos_token_id: int, dtype: torch.dtype, device: str,
                 stop_token_id: Optional[int] = None) -> None:
        super().__init__()
        self.samd_config = samd_config
        self.gen_config: SamdGenerationConfig = None
        self.eos_token = eos_token_id
        self.stop_token = stop_token_id
        self.lm = lm
        self.draft = draft
        self.dtype = dtype
        self.device = device
        self.cache: Optional[SamdStaticCache] = None
        self.forward_state = ForwardState(None)
        self.mask_state = MaskState(None)
        self._verifier: Optional[E.Verifier] = None
        self._verify_out = None
        self.init_buffers()

    # -- buffers ----------------------------------------------------------------------------
    def register_forward_patch(self):
        """Kept for API compatibility: nothing is patched (see module docstring)."""

    def init_seq_position_ids(self):
        return torch.arange(self.samd_config.n_predicts, dtype=torch.long, device=self.device).unsqueeze(0)

    def init_buffers(self):
        self.seq_position_ids = self.init_seq_position_ids()
        buffers = self.draft.tree_model.gen_buffers()
        self.base_tree_attn_mask = buffers["tree_attn_mask"]
        self.base_tree_position_ids = buffers["tree_position_ids"]
        self.base_tree_retrieve_indices = buffers["tree_retrieve_indices"]
        self.update_buffers({})

    def update_buffers(self, buffers_kwargs: Dict[str, Optional[torch.Tensor]]):
        self.tree_attn_mask = buffers_kwargs.get("tree_attn_mask", self.base_tree_attn_mask)
        self.tree_position_ids = buffers_kwargs.get("tree_position_ids", self.base_tree_position_ids)
        self.tree_retrieve_indices = buffers_kwargs.get("tree_retrieve_indices", self.base_tree_retrieve_indices)
        self._retrieve_i32 = self.tree_retrieve_indices.to(torch.int32).contiguous()
        self.mask_state.set_state(self.tree_attn_mask)

    def _tree_mask_4d(self, past: int) -> torch.Tensor:
        """[1, 1, T, past + T] additive mask: every tree node sees the whole past and its own ancestors."""
        tree = self.tree_attn_mask.to(torch.bool)[0, 0]
        t = tree.shape[0]
        allow = torch.cat([torch.ones(t, past, dtype=torch.bool, device=tree.device), tree], dim=1)
        mask = torch.zeros(t, past + t, dtype=self.dtype, device=tree.device)
        mask.masked_fill_(~allow, torch.finfo(self.dtype).min)
        return mask.view(1, 1, t, past + t)

    # -- steps ------------------------------------------------------------------------------
    def prefill(self, input_ids: torch.Tensor, attention_mask: torch.Tensor):
        """samd/samd_model.py:101-128"""
        self.forward_state.forward_type = ForwardType.prefill
        outputs = self.lm(input_ids=input_ids, attention_mask=attention_mask, past_key_values=self.cache)
        logits = outputs.logits
        self.draft.update(tokens=input_ids.squeeze(0), tree_tokens=input_ids.squeeze(0), tree_logits=logits.squeeze(0))
        self.cache.set_length()
        self._next_token = None
        # samd/samd_model.py:124-128: the logits row when greedy, its softmax when sampling
        return logits[:, -1] if self.gen_config.greedy else torch.softmax(logits[:, -1].float(), dim=-1)

    def decode(self, sample_p: torch.Tensor, length: int):
        """samd/samd_model.py:131-182 with the verify / commit tail fused."""
        cands = gen_candidates(sample_p, self.base_tree_retrieve_indices, self.draft, self.samd_config, self.gen_config, self.device,
                               start_token=self._next_token)
        self.update_buffers(cands.buffers_kwargs)
        is_seq = cands.type == CandidateType.sequence
        input_ids = cands.tokens
        if is_seq:
            self.forward_state.forward_type = ForwardType.seq_decode
            position_ids = self.seq_position_ids[:, :input_ids.shape[1]] + length
            outputs = self.lm(input_ids=input_ids, position_ids=position_ids, past_key_values=self.cache)
        else:
            self.forward_state.forward_type = ForwardType.tree_decode
            position_ids = self.tree_position_ids + length
            outputs = self.lm(input_ids=input_ids, position_ids=position_ids, past_key_values=self.cache,
                              attention_mask=self._tree_mask_4d(length))
        tree_logits = outputs.logits                                  # [1, T, V]
        return self.update_state(input_ids, tree_logits, is_seq)

    def update_state(self, tree_tokens: torch.Tensor, tree_logits: torch.Tensor, is_seq: bool):
        """eval_posterior + update_state + select_indices (samd/samd_model.py:159-211) in one launch."""
        if self._verifier is None:
            self._verifier = E.Verifier(1, 256, self.device)
            self._verifier.bind_kv(self.cache.kv_tensors())
        toks = tree_tokens.to(torch.int32).contiguous()
        if not self.gen_config.greedy:
            return self._update_state_sampling(toks, tree_logits, is_seq)
        # Token Recycle: the table update (TokenRecycle.update) rides on the verify launch's pass over the logits
        recycle = getattr(getattr(self.draft, "tree_model", None), "table", None)
        out = self._verifier.verify(tree_logits, toks, None if is_seq else self._retrieve_i32, cache_len=self.cache.cache_len,
                                    move_kv=not is_seq, out=self._verify_out, recycle=recycle)
        self._verify_out = out
        # next step's sample_p = logits row of the last accepted node (samd/utils.py:141)
        acc_idx = (out["accept_len"] - 1).to(torch.long)
        last_node = acc_idx if is_seq else out["indices"][0].to(torch.long).gather(0, acc_idx).clamp_(min=-1)
        sample_p = tree_logits[0].index_select(0, last_node % tree_logits.shape[1]).view(1, -1)
        packed = torch.cat([out["accept_len"], out["tokens"][0]]).tolist()        # the step's one D2H copy
        k = packed[0]
        new_tokens = packed[1:1 + k]
        self._draft_update(out["tokens"][0, :k], tree_tokens.squeeze(0), None if recycle is not None else tree_logits.squeeze(0))
        self.cache.cache_length += k
        self._next_token = out["next_token"]            # = argmax(sample_p): the next step's start token, already on the device
        return sample_p, new_tokens

    def _update_state_sampling(self, toks: torch.Tensor, tree_logits: torch.Tensor, is_seq: bool):
        """The sampling branch (samd/utils.py:142-184): samd_verify_sample (typical acceptance + the residual / plain
        distribution of the next token), then the row moves of select_indices as their own launch."""
        from samd_b200 import _cabi as K
        from .utils import _sampling_state
        g = self.gen_config
        st = _sampling_state(self.device, g.seed)
        out = self._verifier.verify_sample(tree_logits, toks, None if is_seq else self._retrieve_i32, g.temperature, g.top_p,
                                           g.top_k, st["seed"], st["offset"], want_sample_p=True)
        v = self._verifier
        m = v._kv_meta
        with torch.cuda.device(self.device):
            K.check(K.lib().samd_kv_compact(v._kv_ptrs.data_ptr(), m["n_kv"], m["n_heads"], m["row_bytes"], m["batch_stride"],
                                            m["head_stride"], m["pos_stride"], None if is_seq else out["indices"].data_ptr(),
                                            out["indices"].shape[1], out["accept_len"].data_ptr(), self.cache.cache_len.data_ptr(),
                                            1, K.stream_ptr()), "samd_kv_compact")
        packed = torch.cat([out["accept_len"], out["tokens"][0]]).tolist()
        k = packed[0]
        new_tokens = packed[1:1 + k]
        self._draft_update(out["tokens"][0, :k], toks.squeeze(0), tree_logits.squeeze(0))
        self.cache.cache_length += k
        return out["sample_p"], new_tokens

    def _draft_update(self, tokens, tree_tokens, tree_logits):
        self.draft.update(tokens=tokens, tree_tokens=tree_tokens, tree_logits=tree_logits)

    def set_cache(self, generation_config: SamdGenerationConfig):
        if self.samd_config.cache_type == "dynamic":
            raise NotImplementedError("cache_type='dynamic' is not supported; use 'static'")
        if self.cache is None or self.cache.max_cache_len < generation_config.max_cache_len:
            self.cache = SamdStaticCache(self.lm.config, batch_size=1, max_cache_len=generation_config.max_cache_len,
                                         device=self.device, dtype=self.dtype,
                                         hf_device_map=getattr(self.lm, "hf_device_map", None) or {})
            self._verifier = None
        else:
            self.cache.reset()

    def _truncate(self, new_ids):
        """EOS has priority over the stop token (samd/samd_model.py:257-263)."""
        for stop in (self.eos_token, self.stop_token):
            if stop is not None and stop in new_ids:
                return new_ids[:new_ids.index(stop) + 1], True
        return new_ids, False

    @torch.inference_mode()
    def generate(self, input_ids: torch.Tensor, attention_mask: torch.Tensor = None,
                 generation_config: SamdGenerationConfig = None) -> Outputs:
        """samd/samd_model.py:231-274"""
        self.gen_config = generation_config = generation_config or SamdGenerationConfig()
        assert input_ids.shape[0] == 1, "Only support batch_size == 1"
        self.set_cache(generation_config)
        self.draft.reset()
        ids = input_ids.squeeze(0).tolist()
        sample_p = self.prefill(input_ids, attention_mask)
        n_in = input_ids.shape[-1]
        decode_tokens = decode_steps = 0
        accepts = []
        for _ in range(generation_config.max_new_tokens):
            if n_in + decode_tokens + self.samd_config.max_predicts >= generation_config.max_cache_len:
                break
            sample_p, new_ids = self.decode(sample_p, n_in + decode_tokens)
            new_ids, hit_stop = self._truncate(new_ids)
            ids.extend(new_ids)
            decode_steps += 1
            decode_tokens += len(new_ids)
            accepts.append(len(new_ids))
            if hit_stop or decode_tokens >= generation_config.max_new_tokens:
                break
        return Outputs([ids[:n_in + generation_config.max_new_tokens]], decode_tokens, decode_steps, accepts)

    @torch.inference_mode()
    def stream_generate(self, input_ids: torch.Tensor, tokenizer, generation_config: SamdGenerationConfig = None):
        """samd/samd_model.py:277-322"""
        self.gen_config = generation_config = generation_config or SamdGenerationConfig()
        assert input_ids.shape[0] == 1, "Only support batch_size == 1"
        self.set_cache(generation_config)
        self.draft.reset()
        ids = input_ids.squeeze(0).tolist()
        sample_p = self.prefill(input_ids, None)
        n_in = input_ids.shape[-1]
        decode_tokens = 0
        for _ in range(generation_config.max_steps):
            if n_in + decode_tokens + self.samd_config.max_predicts >= generation_config.max_cache_len:
                break
            sample_p, new_ids = self.decode(sample_p, n_in + decode_tokens)
            new_ids, hit_stop = self._truncate(new_ids)
            ids.extend(new_ids)
            yield {"text": tokenizer.decode(ids[n_in:], skip_special_tokens=True, spaces_between_special_tokens=False,
                                            clean_up_tokenization_spaces=True)}
            decode_tokens += len(new_ids)
            if hit_stop or decode_tokens >= generation_config.max_new_tokens:
                break
