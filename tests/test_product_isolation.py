"""The product package must never import, call or link the oracle (or any CPU fallback)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product_sources():
    pkg = os.path.join(ROOT, "starfish_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(dirpath, f)


def test_no_oracle_reference_in_product():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|ref_loader|/root/reference", re.M)
    hits = [p for p in _product_sources() if pat.search(open(p).read())]
    assert not hits, hits


def test_only_allowed_files_touch_oracle():
    allowed = {"bench.py", "__graft_entry__.py"}
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for f in os.listdir(ROOT):
        if f.endswith(".py") and f not in allowed:
            assert not pat.search(open(os.path.join(ROOT, f)).read()), f


def test_importing_package_does_not_need_cuda():
    import importlib

    m = importlib.import_module("starfish_b200")
    assert m.c_kms == 2.99792458e5
