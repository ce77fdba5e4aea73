"""Tree drafters used when the suffix match is short (reference: samd/tree_model/__init__.py).

Only Token-Recycle is provided: it defines the static 61-node tree whose buffers the fused
verification kernel consumes.  EAGLE / EAGLE-2 are separate draft *models* (dense transformer
work) and are outside the retrieval + verification hot path this framework covers."""
from .tree import TreeModel
from .token_recycle import TokenRecycle


def _unsupported(name):
    class _Unsupported(TreeModel):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"tree_method={name!r} is a separate draft model, out of scope of the "
                                      "SAM draft-retrieval hot path; use tree_method='token_recycle'")
    _Unsupported.__name__ = name.capitalize()
    return _Unsupported


Eagle = _unsupported("eagle")
Eagle2 = _unsupported("eagle2")

tree_model_cls = {"token_recycle": TokenRecycle, "eagle": Eagle, "eagle2": Eagle2}
