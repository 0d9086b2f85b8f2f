"""CPU: liby11_b200.so builds for sm_100a here (nvcc cross-compiles), loads, and exports every function that
include/y11.h declares; the ctypes signature table covers the header one to one.  No compute calls."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "y11.h").read_text()


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(y11_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from yolo_infer_b200 import _cabi
    return _cabi.load()


def test_header_declares_functions():
    fns = declared_functions()
    assert "y11_create" in fns and "y11_plan_run" in fns and "y11_detect_postprocess" in fns and len(fns) >= 20


def test_every_declared_symbol_is_exported(lib):
    for fn in declared_functions():
        assert hasattr(lib, fn), f"{fn} declared in include/y11.h but not exported"


def test_signature_table_matches_header(lib):
    from yolo_infer_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_functions()
    assert lib.y11_abi_version() == _cabi.ABI_VERSION == int(re.search(r"#define Y11_ABI_VERSION (\d+)", HEADER).group(1))


def test_library_contains_blackwell_sass():
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    from yolo_infer_b200 import _cabi
    _cabi.load()
    out = subprocess.run(["cuobjdump", "-sass", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, mnemonic
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout


def test_no_cpu_fallback_without_gpu(lib):
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.y11_create(C.byref(h), 0)
    assert rc != 0 and b"no CUDA device" in lib.y11_last_error()
    from yolo_infer_b200 import YOLO11Model
    with pytest.raises(RuntimeError):
        YOLO11Model("yolo11n.yaml")
    with pytest.raises(RuntimeError):
        YOLO11Model("yolo11n.yaml", device="cpu")
