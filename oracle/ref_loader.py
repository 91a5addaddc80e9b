"""Locate and import the UNMODIFIED reference (TEST / BENCHMARK INFRASTRUCTURE, never imported by the product).

Search order: ``$VQA_REFERENCE`` / ``/root/reference`` (the build container), then ``oracle/_ref/reference.zip`` (what
``tools/stage_reference.sh`` stages; git-ignored, shipped to the GPU box by gpurun).  The reference's ``utils/config.py``
creates directories relative to the current directory at import time, so the import runs in a scratch
directory.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "reference.zip")


def reference_path():
    """Directory or zip archive holding the reference's ``models`` package, or None."""
    for cand in (os.environ.get("VQA_REFERENCE"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "models")):
            return cand
    return ARCHIVE if os.path.isfile(ARCHIVE) else None


def load_reference():
    """-> (models.vqa_model module, where it came from) or (None, reason)."""
    path = reference_path()
    if path is None:
        return None, "reference not staged (run tools/stage_reference.sh in the build container)"
    cwd = os.getcwd()
    try:
        os.chdir(tempfile.mkdtemp(prefix="vqa_ref_cwd_"))
        if path not in sys.path:
            sys.path.insert(0, path)
        mod = importlib.import_module("models.vqa_model")
    finally:
        os.chdir(cwd)
    return mod, path
