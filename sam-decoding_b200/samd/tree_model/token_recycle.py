"""Token-Recycle fallback drafter (reference: samd/tree_model/token_recycle/token_recycle.py:18-63,
utils.py:37-99): remembers, per token, the top-8 successors the LM last predicted after it and fills
a static 61-node tree from that table.  The table is a device array; its update is fused into the
verification pass (samd_verify_compact, out_topk_dev / recycle_table_dev)."""
from typing import Dict, List

import torch

from ..samd_config import SamdConfig
from .tree import TreeModel

TOPK = 8


def gen_buffers(tree: List[List[int]], device) -> Dict[str, torch.Tensor]:
    """Ancestor mask [1,1,n,n], depth per node [1,n], root-to-leaf paths [leaves, depth] (-1 padded,
    leaves in reversed node order, utils.py:77-90)."""
    n = len(tree)
    parent = [-1] * n
    for node, childs in enumerate(tree):
        for c in childs:
            parent[c] = node
    level = [0] * n
    for i in range(1, n):
        level[i] = level[parent[i]] + 1
    mask = torch.zeros(n, n)
    for i in range(n):
        j = i
        while j != -1:
            mask[i, j] = 1.0
            j = parent[j]
    paths = []
    for node, childs in enumerate(tree):
        if childs:
            continue
        p = [node]
        while p[-1] != 0:
            p.append(parent[p[-1]])
        paths.append(p[::-1])
    paths.reverse()
    width = max(level) + 1
    retrieve = torch.tensor([p + [-1] * (width - len(p)) for p in paths], dtype=torch.long)
    out = {"tree_attn_mask": mask.view(1, 1, n, n), "tree_position_ids": torch.tensor([level], dtype=torch.long),
           "tree_retrieve_indices": retrieve}
    return {k: v.to(device) for k, v in out.items()}


class TokenRecycle(TreeModel):
    """Same surface as the reference class; the token -> top-8 table lives on the device (engine.RecycleTable).
    `update` is one launch over the [T, V] logits (or nothing at all when SamdModel already updated the table inside
    its verify launch), `gen_draft` one small launch + one copy of the 61 tree tokens."""

    def __init__(self, config: SamdConfig, lm=None, dtype: torch.dtype = None, device: str = "cuda") -> None:
        super().__init__()
        from samd_b200 import engine as E
        self.samd_config = config
        self.dtype = dtype
        self.device = device
        self.tree = config.tree
        self.table = E.RecycleTable(self.tree)

    @property
    def cache(self) -> Dict[int, List[int]]:
        return self.table.as_dict()

    def reset(self):
        pass                                   # the table survives across requests, as in the reference

    def update(self, tree_tokens: torch.Tensor = None, tree_logits: torch.Tensor = None, **kwargs):
        if tree_tokens is None or tree_logits is None:
            return
        self.table.update(tree_tokens.reshape(-1), tree_logits.reshape(-1, tree_logits.shape[-1]))

    def gen_draft(self, start_token: int):
        if self.table.table is None:           # nothing learnt yet: the reference's empty-cache tree
            return [int(start_token)] + [0] * (len(self.tree) - 1), {}
        start = torch.tensor([int(start_token)], dtype=torch.int32, device=self.table.device)
        return self.table.gen_tree(start)[0].tolist(), {}

    def gen_buffers(self) -> Dict[str, torch.Tensor]:
        return gen_buffers(self.samd_config.tree, self.device)
