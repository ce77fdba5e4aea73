"""Drop-in for the reference's `samd_sam_only` package (samd_sam_only/__init__.py:1-5): the retrieval-only flavour
(dynamic sequence drafts, static best-first tree drafts) with the same public names as `samd`."""
from . import draft as _draft, sam as _sam, samd_config as _config, samd_model as _model, utils as _utils

SamdConfig, SamdGenerationConfig = _config.SamdConfig, _utils.SamdGenerationConfig
build_sam, dump_sam, load_sam = _sam.build_sam, _sam.dump_sam, _sam.load_sam
DraftModel, SamdModel = _draft.DraftModel, _model.SamdModel

__all__ = ["SamdConfig", "SamdGenerationConfig", "build_sam", "dump_sam", "load_sam", "DraftModel", "SamdModel"]
