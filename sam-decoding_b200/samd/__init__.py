"""Drop-in for the reference's `samd` package (samd/__init__.py:1-5): same public names, CUDA behind a C ABI.

    SamdConfig, SamdGenerationConfig   configuration objects
    build_sam / dump_sam / load_sam    static suffix automaton: build on the host, flat file or reference pickle
    DraftModel                         dynamic + static automaton + tree fallback behind lookup / update / reset
    SamdModel                          generate / stream_generate
"""
from . import draft as _draft, sam as _sam, samd_config as _config, samd_model as _model, utils as _utils

SamdConfig, SamdGenerationConfig = _config.SamdConfig, _utils.SamdGenerationConfig
build_sam, dump_sam, load_sam = _sam.build_sam, _sam.dump_sam, _sam.load_sam
DraftModel, SamdModel = _draft.DraftModel, _model.SamdModel

__all__ = ["SamdConfig", "SamdGenerationConfig", "build_sam", "dump_sam", "load_sam", "DraftModel", "SamdModel"]
