"""Drop-in `samd.sam.DynSAM` (reference: samd/sam/dyn_sam.py:8-113), device resident."""
from dataclasses import dataclass
from typing import Dict, List

from samd_b200 import _cabi as K
from samd_b200.dropin import DynSamView


class DynSAM(DynSamView):
    _FLAVOUR = K.FLAVOUR_SAMD

    @dataclass
    class SAMState:
        next: Dict[int, int]
        link: int
        length: int
        min_endpos: int

    def __init__(self, n_predicts: int = 40, device: str = "cuda"):
        self.n_predicts = n_predicts
        self._core_init(device)

    def gen_draft(self, index: int, start_token: int) -> List[int]:
        """dyn_sam.py:107-113: ancestor walk, then n_predicts tokens, zero padded."""
        return self._gen_draft(index, 0, start_token, self.n_predicts, 0.0)

    @property
    def states(self):
        return self._export_states(lambda nxt, link, length, end: DynSAM.SAMState(nxt, link, length, end))
