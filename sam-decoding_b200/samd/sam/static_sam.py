"""Drop-in `samd.sam.StaticSAM` / `NullStaticSAM` (reference: samd/sam/static_sam.py:8-137).

Module path, class name and the nested `SAMState` dataclass match the reference so that its
pickles (`dump_sam`, samd/sam/utils.py:20-22) unpickle into these classes.
"""
from dataclasses import dataclass
from typing import Dict, List

from samd_b200.dropin import StaticSamView


class StaticSAM(StaticSamView):
    _WITH_COUNTS = False

    @dataclass
    class SAMState:
        next: Dict[int, int]
        link: int
        length: int
        min_endpos: int

    def __init__(self, n_predicts: int = 40, device: str = "cuda"):
        self.n_predicts = n_predicts
        self._core_init(device)

    @staticmethod
    def build(batch_tokens: List[List[int]], eos_token: int, verbose: bool = True, device: str = "cuda"):
        """static_sam.py:38-46"""
        return StaticSAM._build(batch_tokens, eos_token, device=device)

    def gen_draft(self, index: int, start_token: int) -> List[int]:
        """static_sam.py:119-125 (no ancestor walk)."""
        return self._gen_seq_draft(index, start_token, self.n_predicts)

    @property
    def states(self):
        graph = self.__dict__.get("states")
        if graph is not None:
            return graph
        return self._export_states(lambda nxt, link, length, end: StaticSAM.SAMState(nxt, link, length, end))

    @property
    def input_ids(self):
        ids = self.__dict__.get("input_ids")
        return ids if ids is not None else self._export_text()

    @property
    def max_length(self):
        return self._ensure().n_tokens


class NullStaticSAM(StaticSAM):
    """static_sam.py:128-137: root-only automaton; every lookup answers (0, 0)."""

    def __init__(self, n_predicts: int = 40, device: str = "cuda"):
        super().__init__(n_predicts, device)

    def reset(self):
        pass

    def transfer_tokens(self, tokens):
        pass

    def lookup(self, token: int):
        return 0, 0

    def gen_draft(self, index, start_token):
        return -1, -1
