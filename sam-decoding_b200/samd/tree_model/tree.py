"""Contract of the fallback tree drafter DraftModel falls back to when no suffix match is long enough
(reference: samd/tree_model/tree.py:9-30).  Implementations: token_recycle.TokenRecycle."""
import torch


class TreeModel(torch.nn.Module):
    """reset() per conversation, update(...) after every verified step, gen_draft(start) -> (tree tokens, buffer
    overrides), gen_buffers() -> the static tree's mask / position / path tensors."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__()

    def _abstract(self, name):
        raise NotImplementedError(f"{type(self).__name__}.{name}")

    def reset(self):
        self._abstract("reset")

    def update(self, **kwargs):
        self._abstract("update")

    def gen_draft(self, start_token: int):
        self._abstract("gen_draft")

    def gen_buffers(self):
        self._abstract("gen_buffers")
