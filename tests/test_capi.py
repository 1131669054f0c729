"""The C-ABI library loads and exports every symbol include/sfb200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from starfish_b200 import build

    return build.build()


def _declared():
    text = open(os.path.join(ROOT, "include", "sfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfb_[A-Za-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(built_lib):
    from starfish_b200 import _lib

    assert sorted(_lib.EXPORTS) == _declared()


def test_library_exports_every_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.sfb_abi_version.restype = ctypes.c_int
    assert lib.sfb_abi_version() == 5


def test_library_is_sm100a_only(built_lib):
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_create_fails_cleanly_without_gpu(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from starfish_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.sfb_create(0, 256, 6, 2, 4, 0, ctypes.byref(h)) < 0
    assert not h.value
    from starfish_b200.engine import LikelihoodEngine

    with pytest.raises(RuntimeError):
        LikelihoodEngine(256, 6, 2, 4)


def test_sass_uses_dmma_and_tma(built_lib):
    """The fp64 tensor path (DMMA.8x8x4), TMA tensor loads (UTMALDG), bulk async copies (UBLKCP) and the
    mbarrier pipeline (SYNCS) must be in the shipped SASS."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("DMMA.8x8x4", "UTMALDG.3D", "UBLKCP", "SYNCS.ARRIVE.TRANS64", "LDS.128"):
        assert mnemonic in sass, mnemonic


def test_header_is_plain_c():
    """include/sfb200.h must be consumable from C (cgo / JNI / FFI generators read it): C99, no warnings."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    res = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror",
                          os.path.join(ROOT, "include", "sfb200.h")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
