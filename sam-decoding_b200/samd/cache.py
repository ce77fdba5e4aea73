"""Drop-in `samd.cache.SamdStaticCache` (reference: samd/cache.py:37-133) on transformers >= 5.

Same constructor, same attributes (`key_cache`, `value_cache`, `cache_length`, `last_length`) and
same methods.  All 2L tensors are views of ONE allocation [2L, B, H_kv, max_len, D_h] so that the
compaction kernel addresses them through a device pointer table; `select_indices` is one launch
(`samd_kv_compact`) instead of 4L index_select/copy_ launches, and `SamdModel` goes further and
fuses it into the verification kernel.  Lengths are per request on the device (`cache_len`), with
`cache_length` kept as the reference's host integer for batch size 1.
"""
from typing import Any, Dict, List, Optional, Tuple

import torch
from transformers.cache_utils import Cache, CacheLayerMixin

from profile_utils import profile_decorator  # noqa: F401
from samd_b200 import _cabi as K


class _SamdLayer(CacheLayerMixin):
    """Per-layer adaptor transformers' Cache container expects; state lives in the parent cache."""

    is_compileable = False
    is_sliding = False

    def __init__(self, parent: "SamdStaticCache", layer_idx: int):
        super().__init__()
        self.parent, self.layer_idx = parent, layer_idx
        self.is_initialized = True

    def lazy_initialization(self, key_states, value_states) -> None:
        pass

    def update(self, key_states, value_states, *args, **kwargs):
        return self.parent._update_layer(key_states, value_states, self.layer_idx)

    def get_mask_sizes(self, query_length: int) -> Tuple[int, int]:
        return self.parent.cache_length + int(query_length), 0

    def get_seq_length(self) -> int:
        return self.parent.cache_length

    def get_max_cache_shape(self) -> int:
        return self.parent.max_cache_len

    def reset(self) -> None:
        pass


class SamdStaticCache(Cache):

    def __init__(self, config, batch_size=None, max_cache_len=None, device=None, dtype=torch.float32,
                 max_batch_size=None, hf_device_map=None):
        n_layers = config.num_hidden_layers
        if hf_device_map is not None and len(hf_device_map) > 1:
            raise NotImplementedError("layer-wise device_map placement is not supported: shard requests across GPUs instead")
        super().__init__(layers=[_SamdLayer(self, i) for i in range(n_layers)])
        self.batch_size = batch_size or max_batch_size
        self._max_cache_len = config.max_position_embeddings if max_cache_len is None else max_cache_len
        self.head_dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
        self.dtype = dtype
        self.num_key_value_heads = getattr(config, "num_key_value_heads", None) or config.num_attention_heads
        self.device = torch.device(device if device is not None else "cuda")
        # one allocation, 2L views: key_cache[l] = kv[l], value_cache[l] = kv[L + l]   (cache.py:72-85)
        self.kv = torch.zeros((2 * n_layers, self.batch_size, self.num_key_value_heads, self._max_cache_len, self.head_dim),
                              dtype=dtype, device=self.device)
        self.key_cache: List[torch.Tensor] = [self.kv[i] for i in range(n_layers)]
        self.value_cache: List[torch.Tensor] = [self.kv[n_layers + i] for i in range(n_layers)]
        self.last_length = 0
        self.cache_length = 0
        if self.device.type == "cuda":
            self.cache_len = torch.zeros(self.batch_size, dtype=torch.int32, device=self.device)
            self.kv_ptrs = torch.tensor([t.data_ptr() for t in self.key_cache + self.value_cache], dtype=torch.int64,
                                        device=self.device)
        else:
            self.cache_len = self.kv_ptrs = None

    @property
    def max_cache_len(self) -> int:
        return self._max_cache_len

    def kv_tensors(self) -> List[torch.Tensor]:
        return self.key_cache + self.value_cache

    def reset(self):
        self.cache_length = 0
        self.last_length = 0
        if self.cache_len is not None:
            self.cache_len.zero_()

    def set_length(self):
        self.cache_length = self.last_length
        if self.cache_len is not None:
            self.cache_len.fill_(self.cache_length)

    def get_seq_length(self, layer_idx=0):
        return self.cache_length

    def get_max_cache_shape(self, layer_idx=0) -> Optional[int]:
        return self._max_cache_len

    def _update_layer(self, key_states, value_states, layer_idx):
        """cache.py:103-115: write the new rows at [cache_length, +T), return views of [0, last_length)."""
        k_out, v_out = self.key_cache[layer_idx], self.value_cache[layer_idx]
        t = key_states.shape[2]
        k_out.narrow(2, self.cache_length, t).copy_(key_states)
        v_out.narrow(2, self.cache_length, t).copy_(value_states)
        if layer_idx == 0:
            self.last_length = self.cache_length + t
        return k_out.narrow(2, 0, self.last_length), v_out.narrow(2, 0, self.last_length)

    def update(self, key_states, value_states, layer_idx, cache_kwargs: Optional[Dict[str, Any]] = None, *args, **kwargs):
        return self._update_layer(key_states, value_states, layer_idx)

    def select_indices(self, indices: Optional[torch.Tensor] = None, accept_length: int = 1):
        """cache.py:118-133: rows cache_length + indices[j] -> cache_length + j in every K and V tensor
        (one launch), then cache_length += accept_length.  `indices is None` (sequence draft) only bumps."""
        accept_length = int(accept_length)
        if indices is not None:
            if self.kv_ptrs is None:
                raise K.SamdError("SamdStaticCache.select_indices needs a CUDA cache (no CPU fallback)")
            idx = indices.reshape(1, -1).to(device=self.device, dtype=torch.int32).contiguous()
            acc = torch.full((self.batch_size,), accept_length, dtype=torch.int32, device=self.device)
            t0 = self.key_cache[0]
            es = t0.element_size()
            self.cache_len.fill_(self.cache_length)
            with torch.cuda.device(self.device):
                K.check(K.lib().samd_kv_compact(self.kv_ptrs.data_ptr(), self.kv_ptrs.numel(), t0.shape[1], t0.shape[3] * es,
                                                t0.stride(0) * es, t0.stride(1) * es, t0.stride(2) * es,
                                                idx.expand(self.batch_size, -1).contiguous().data_ptr(), idx.shape[1],
                                                acc.data_ptr(), self.cache_len.data_ptr(), self.batch_size, K.stream_ptr()),
                        "samd_kv_compact")
        elif self.cache_len is not None:
            self.cache_len.add_(accept_length)
        self.cache_length += accept_length

    def sync_length(self):
        """After a fused verify+compact launch (which bumped cache_len on the device) bring the
        reference's host-side integer up to date (batch size 1)."""
        self.cache_length = int(self.cache_len[0].item())
        return self.cache_length


class SamdCache:
    """The reference's DynamicCache variant (samd/cache.py:8-34) is not provided: `cache_type`
    defaults to "static" (samd/samd_config.py:16-18) and the dynamic variant re-allocates every step."""

    def __init__(self, *a, **k):
        raise NotImplementedError("cache_type='dynamic' is not supported; use the default cache_type='static'")
