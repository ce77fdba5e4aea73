"""Drop-in for the reference's top-level `profile_utils` module (profile_utils.py:1-88), which
samd/draft.py:11, samd/utils.py:15 and samd/cache.py:6 import.

Same switches and report functions, but timings are taken with CUDA events on the current
stream when a GPU is present (the reference used perf_counter without a synchronize, which
mis-attributes asynchronous GPU time).  Disabled by default: the wrappers then cost one branch.
"""
from __future__ import annotations

import json
import time
from collections import defaultdict
from functools import wraps
from typing import Dict, List

fn_dict: Dict[str, List[float]] = defaultdict(list)
lookup_dict: Dict[str, List[str]] = defaultdict(list)
accept_dict: Dict[str, List[int]] = defaultdict(list)
decorator_flag: bool = False


def enable_decorator(mode: bool):
    global decorator_flag
    decorator_flag = bool(mode)


def clear_dict():
    fn_dict.clear()
    lookup_dict.clear()


def _timed(fn, args, kwargs):
    try:
        import torch
        on_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        on_gpu = False
    if on_gpu:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        e1.synchronize()
        return out, e0.elapsed_time(e1) * 1e-3
    t0 = time.perf_counter()
    out = fn(*args, **kwargs)
    return out, time.perf_counter() - t0


def profile_decorator(fn_name: str):
    def deco(fn):
        @wraps(fn)
        def wrapper(*args, **kwargs):
            if not decorator_flag:
                return fn(*args, **kwargs)
            out, dt = _timed(fn, args, kwargs)
            fn_dict[fn_name].append(dt)
            return out
        return wrapper
    return deco


def profile_lookup_decorator(fn_name: str):
    def deco(fn):
        @wraps(fn)
        def wrapper(*args, **kwargs):
            out = fn(*args, **kwargs)
            if decorator_flag:
                lookup_dict[fn_name].append(out[0])
            return out
        return wrapper
    return deco


def profile_accept_length(name: str, length: int):
    if decorator_flag:
        accept_dict[name].append(length)


def export_result(root_name: str = "forward"):
    if not fn_dict:
        return None
    totals = {name: sum(v) for name, v in fn_dict.items()}
    root = totals.get(root_name) or max(totals.values())
    width = max(len(n) for n in totals)
    lines = [f"{'name':<{width}}  {'time':>12}  {'ratio':>8}"]
    for name, t in totals.items():
        lines.append(f"{name:<{width}}  {t:12.6f}  {t / root:8.4f}")
    return "\n".join(lines)


def export_lookup_result():
    counts, means = {}, {}
    for name, kinds in lookup_dict.items():
        counts[name], sums = {}, {}
        for kind in kinds:
            counts[name][str(kind)] = counts[name].get(str(kind), 0) + 1
        for kind, length in zip(kinds, accept_dict[name]):
            sums[str(kind)] = sums.get(str(kind), 0) + length
        means[name] = {k: sums.get(k, 0) / c for k, c in counts[name].items()}
    return json.dumps({"result-1": counts, "result-2": means}, indent=4, ensure_ascii=False)
