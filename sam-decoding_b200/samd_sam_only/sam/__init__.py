"""Suffix automata of the `samd_sam_only` flavour (count-based static tree drafter): device-backed views with the
reference's class names (samd_sam_only/sam/__init__.py), plus build / dump / load."""
from . import dyn_sam as _dyn, static_sam as _static, utils as _io

DynSAM, StaticSAM = _dyn.DynSAM, _static.StaticSAM
build_sam, dump_sam, load_sam = _io.build_sam, _io.dump_sam, _io.load_sam

__all__ = ["DynSAM", "StaticSAM", "build_sam", "dump_sam", "load_sam"]
