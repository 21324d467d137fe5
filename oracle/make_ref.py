#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, staged so they travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is two plain Python files
(src/ProbabilisticLayers.py, src/ProbabilisticLayers_Loss.py; MIT licence) with no build step, so "building" the
reference arm is a byte-for-byte copy from the read-only checkout at /root/reference into ``oracle/_ref/``.
That directory is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so
``bench.py --impl reference`` and the ``gpu_eager_reference`` leg can time the reference's own code on the GPU box.

    python oracle/make_ref.py          # in the builder container; __graft_entry__.build() calls it too

``load()`` imports the staged files by path after stubbing the import-time-only dependencies that are absent in this
image (matplotlib, future -- SURVEY.md appendix A); nothing in them is edited.
"""
import hashlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ("src/ProbabilisticLayers.py", "src/ProbabilisticLayers_Loss.py", "LICENSE")


def stage(quiet=False):
    """Copy the reference files into oracle/_ref/ (no-op when /root/reference is absent).  Returns True when
    oracle/_ref/ holds them afterwards."""
    if os.path.isdir(REF_SRC):
        os.makedirs(DST, exist_ok=True)
        sums = []
        for rel in FILES:
            src = os.path.join(REF_SRC, rel)
            if not os.path.exists(src):
                continue
            dst = os.path.join(DST, os.path.basename(rel))
            shutil.copyfile(src, dst)
            with open(dst, "rb") as f:
                sums.append(f"{hashlib.sha256(f.read()).hexdigest()}  {os.path.basename(rel)}")
        with open(os.path.join(DST, "SHA256SUMS"), "w") as f:
            f.write("\n".join(sums) + "\n")
        if not quiet:
            print("\n".join(sums))
    return available()


def available():
    return all(os.path.exists(os.path.join(DST, n)) for n in ("ProbabilisticLayers.py", "ProbabilisticLayers_Loss.py"))


_LOADED = None


def load():
    """(ProbabilisticLayers module, ProbabilisticLayers_Loss module) of the staged, unmodified reference."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise FileNotFoundError("oracle/_ref/ is empty: run `python oracle/make_ref.py` in the builder container")
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "future"):
        sys.modules.setdefault(n, types.ModuleType(n))
    if not hasattr(sys.modules["matplotlib"], "rcParams"):
        sys.modules["matplotlib"].rcParams = {}
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["matplotlib.lines"], "Line2D"):
        sys.modules["matplotlib.lines"].Line2D = object

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pl = _load("ref_PL_staged", os.path.join(DST, "ProbabilisticLayers.py"))
        loss = _load("ref_Loss_staged", os.path.join(DST, "ProbabilisticLayers_Loss.py"))
    _LOADED = (pl, loss)
    return _LOADED


if __name__ == "__main__":
    ok = stage()
    print("oracle/_ref ready" if ok else "reference checkout not found; oracle/_ref not staged")
