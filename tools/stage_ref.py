#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the hot path for the GPU box.

``/root/reference`` exists only in the build container.  This recipe copies the three pure-Python sub-packages the path
needs (``paos/classes``, ``paos/core``, ``paos/util``: ~170 KB, byte for byte, nothing edited) into the git-ignored
``oracle/_ref/paos/`` so that they travel with the snapshot like the built ``.so`` does, and records the sha1 of every file
in ``oracle/_ref/MANIFEST.json``.  ``oracle/refload.py`` then loads them through its stub loader (the third-party modules
that are absent from this image -- astropy.units, photutils, skimage, matplotlib -- are stubbed exactly as in the
container) and ``bench.py --impl reference`` times them: ``cpu_baseline.kind = "reference"``.

Nothing here is product code and nothing under ``oracle/_ref`` is ever committed (``.gitignore``); the product never imports it.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/paos"
DST = os.path.join(ROOT, "oracle", "_ref")
PARTS = ("classes", "core", "util")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "classes")):
        return None
    manifest = {}
    for part in PARTS:
        dst = os.path.join(DST, "paos", part)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, part), dst, ignore=shutil.ignore_patterns("__pycache__", "retired", "*.pyc"))
        for base, _, files in os.walk(dst):
            for f in sorted(files):
                p = os.path.join(base, f)
                with open(p, "rb") as fh:
                    manifest[os.path.relpath(p, DST)] = hashlib.sha1(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
