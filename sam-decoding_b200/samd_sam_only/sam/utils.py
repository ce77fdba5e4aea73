"""build_sam / dump_sam / load_sam (reference: samd_sam_only/sam/utils.py:10-39)."""
from typing import List

from samd_b200 import dropin
from .static_sam import StaticSAM


def build_sam(batch_tokens: List[List[int]], eos_token: int, device: str = "cuda"):
    return StaticSAM.build(batch_tokens, eos_token, device=device)


def dump_sam(path: str, sam: StaticSAM):
    dropin.dump_sam(path, sam)


def load_sam(path: str, device: str = "cuda"):
    sam = dropin.load_sam(path, StaticSAM, device=device)
    assert type(sam) is StaticSAM
    sam.device = device
    return sam
