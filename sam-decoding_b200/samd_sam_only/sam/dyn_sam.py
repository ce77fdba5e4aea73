"""Drop-in `samd_sam_only.sam.DynSAM` (reference: samd_sam_only/sam/dyn_sam.py:11-121)."""
from dataclasses import dataclass
from typing import Dict

import torch

from samd_b200 import _cabi as K
from samd_b200.dropin import DynSamView


class DynSAM(DynSamView):
    _FLAVOUR = K.FLAVOUR_SAM_ONLY

    @dataclass
    class SAMState:
        next: Dict[int, int]
        link: int
        length: int
        min_endpos: int

    def __init__(self, max_predicts: int = 40, alpha: float = 4.0, device: str = "cuda"):
        self.max_predicts = max_predicts
        self.alpha = alpha
        self.device = device
        self._core_init(device)

    def gen_draft(self, index: int, match_length: int, start_token: int):
        """dyn_sam.py:116-121: n = min(max_predicts, 1 + int(match * alpha)) tokens, unpadded."""
        seq = self._gen_draft(index, match_length, start_token, self.max_predicts, self.alpha)
        pos = torch.arange(0, len(seq), dtype=torch.long, device=self.device).unsqueeze(0)
        return seq, {"seq_position_ids": pos}

    @property
    def states(self):
        return self._export_states(lambda nxt, link, length, end: DynSAM.SAMState(nxt, link, length, end))
