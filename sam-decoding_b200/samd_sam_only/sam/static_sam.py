"""Drop-in `samd_sam_only.sam.StaticSAM` (reference: samd_sam_only/sam/static_sam.py:22-215):
occurrence counts, stable top-k successors and the best-first tree drafter, all on the device.
Module path / class names match the reference so that its pickles load."""
from dataclasses import dataclass
from typing import Dict, List

import torch

from samd_b200 import _cabi as K
from samd_b200.dropin import StaticSamView


def tree_buffers_from_parents(parents: torch.Tensor, depth: torch.Tensor, retrieve: torch.Tensor) -> Dict[str, torch.Tensor]:
    """gen_buffers (static_sam.py:148-180) from the kernel's per-node parents / depths / paths."""
    n = parents.numel()
    mask = torch.eye(n, dtype=torch.bool, device=parents.device)
    cur = parents.to(torch.long)
    rows = torch.arange(n, device=parents.device)
    for _ in range(int(depth.max().item()) if n > 1 else 0):
        ok = cur >= 0
        mask[rows[ok], cur[ok]] = True
        cur = torch.where(ok, parents.to(torch.long)[cur.clamp(min=0)], cur)
    return {"tree_attn_mask": mask.view(1, 1, n, n), "tree_position_ids": depth.to(torch.long).unsqueeze(0),
            "tree_retrieve_indices": retrieve.to(torch.long)}


class StaticSAM(StaticSamView):
    _WITH_COUNTS = True

    @dataclass
    class SAMState:
        next: Dict[int, int]
        link: int
        length: int
        cnt_endpos: int

    def __init__(self, max_predicts: int = 40, alpha: float = 4.0, K: int = 8, device: str = "cuda"):
        self.max_predicts = max_predicts
        self.alpha = alpha
        self.device = device
        self.K = K
        self.states_topk_next = None
        self._core_init(device)

    @staticmethod
    def build(batch_tokens: List[List[int]], eos_token: int, verbose: bool = True, device: str = "cuda"):
        """static_sam.py:31-39 (the top-k table is built with the automaton)."""
        sam = StaticSAM._build(batch_tokens, eos_token, device=device)
        sam.device = device
        return sam

    def init_topk_next(self, k: int = 8):
        """static_sam.py:137-146: done by the builder; kept for API compatibility."""
        self._ensure()

    def gen_draft(self, index: int, match_length: int, start_token: int):
        """static_sam.py:182-215 + gen_buffers :148-180.  `match_length` is already biased by the caller."""
        sam = self._ensure()
        dev = sam.device
        n = int(self.max_predicts)
        st = self.__dict__.setdefault("_tree_buf", {})
        if st.get("n") != n:
            mk = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
            st.update(n=n, tok=mk(1, n), par=mk(1, n), dep=mk(1, n), cnt=mk(1), shape=mk(1, 2), ret=mk(1, n, n), m=mk(1))
        self._idx1.fill_(int(index))
        self._tok1.fill_(int(start_token))
        st["m"].fill_(int(match_length))
        with torch.cuda.device(dev):
            K.check(K.lib().samd_static_tree_draft(sam.handle, 1, None, self._idx1.data_ptr(), st["m"].data_ptr(),
                                                   self._tok1.data_ptr(), n, float(self.alpha), int(self.K), 0,
                                                   st["tok"].data_ptr(), st["par"].data_ptr(), st["dep"].data_ptr(),
                                                   st["cnt"].data_ptr(), st["ret"].data_ptr(), n, n, st["shape"].data_ptr(),
                                                   K.stream_ptr()), "samd_static_tree_draft")
        nn, leaves, width = torch.cat([st["cnt"], st["shape"][0]]).tolist()
        tree = st["tok"][0, :nn].tolist()
        return tree, tree_buffers_from_parents(st["par"][0, :nn], st["dep"][0, :nn], st["ret"][0, :leaves, :width])

    @property
    def states(self):
        graph = self.__dict__.get("states")
        if graph is not None:
            return graph
        return self._export_states(lambda nxt, link, length, cnt: StaticSAM.SAMState(nxt, link, length, cnt))

    @property
    def max_length(self):
        return self._ensure().n_tokens
