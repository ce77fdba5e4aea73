#!/usr/bin/env python
"""Stage the UNMODIFIED reference packages next to the repo so that bench.py's reference arm can time the
reference's own classes on the GPU box's host cores.

  python baseline/stage_reference.py [--src /root/reference]

The reference (hyx1999/SAM-Decoding) is pure Python without packaging metadata, so there is nothing to pip-install:
this copies samd/, samd_sam_only/ and profile_utils.py verbatim into baseline/_ref/ (git-ignored - reference sources
never enter this repository's history - but not gpurun-ignored, so the directory travels to the GPU box with the
snapshot).  __graft_entry__.build() calls it whenever the reference checkout is present.  The files are loaded with
oracle/ref_loader.py's stub-package trick (the package __init__ pulls in model_patch/llama.py, which does not import
under transformers 5), i.e. the on-path modules themselves run as they are.
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def stage(src: str = "/root/reference", quiet: bool = False) -> bool:
    if not os.path.isdir(os.path.join(src, "samd", "sam")):
        if not quiet:
            print(f"reference checkout not found at {src}; nothing staged", file=sys.stderr)
        return False
    ignore = shutil.ignore_patterns("*.bak*", "__pycache__", "*.pyc", "inference")
    for pkg in ("samd", "samd_sam_only"):
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(src, pkg), dst, ignore=ignore)
    shutil.copy2(os.path.join(src, "profile_utils.py"), os.path.join(DST, "profile_utils.py"))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(src + "\n")
    if not quiet:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"staged {n} files under {DST}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    sys.exit(0 if stage(a.src) else 1)
