"""Suffix automata of the `samd` flavour: device-backed views with the reference's class names
(samd/sam/__init__.py) - the dynamic automaton, the static one and its null stand-in, plus build / dump / load."""
from . import dyn_sam as _dyn, static_sam as _static, utils as _io

DynSAM = _dyn.DynSAM
StaticSAM, NullStaticSAM = _static.StaticSAM, _static.NullStaticSAM
build_sam, dump_sam, load_sam = _io.build_sam, _io.dump_sam, _io.load_sam

__all__ = ["DynSAM", "StaticSAM", "NullStaticSAM", "build_sam", "dump_sam", "load_sam"]
