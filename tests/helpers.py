"""Shared helpers for the parity tests: golden-fixture access and replay drivers."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unragged(flat, offs):
    return [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def docs_of(z, prefix=""):
    key = (prefix + "/") if prefix else ""
    return unragged(z[key + "docs_flat"], z[key + "docs_offs"])
