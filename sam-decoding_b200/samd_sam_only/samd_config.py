"""SamdConfig of the SAM-only variant (reference: samd_sam_only/samd_config.py:9-69)."""
from dataclasses import dataclass, field
from typing import Literal

from samd.samd_config import ForwardType, ForwardState, MaskState, load_token_recycle, load_eagle, load_eagle2  # noqa: F401


@dataclass
class SamdConfig:
    max_predicts: int = field(default=60)
    alpha: float = field(default=4.0)
    K: int = field(default=8)
    len_bias: int = field(default=5)
    cache_type: Literal["dynamic", "static"] = field(default="static")
