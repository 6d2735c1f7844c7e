"""The C-ABI library loads and exports every symbol include/figdraw_cuda.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT
from figdraw_b200 import abi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "figdraw_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in figdraw_cuda.h but not exported"
    assert sorted(abi.EXPORTS) == syms
    assert lib.fdc_abi_version() == abi.ABI_VERSION


def test_call_record_layout():
    assert abi.CALL_DTYPE.itemsize == 128
    assert abi.CALL_DTYPE.fields["u"][1] == 4 and abi.CALL_DTYPE.fields["f"][1] == 40
    assert ctypes.sizeof(abi.FdcFill) == 28


def test_sdf_mode_values_match_reference():
    # figbackend.nim:36-52
    m = abi.SdfMode
    assert (m.sdfModeAtlas, m.sdfModeClipAA, m.sdfModeDropShadow, m.sdfModeInsetShadow) == (0, 3, 7, 9)
    assert (m.sdfModeAnnularAA, m.sdfModeMsdf, m.sdfModeMtsdfAnnular, m.sdfModeBackdropBlur) == (12, 13, 16, 17)
    assert m.sdfModeBezierStrokeSquareAA == 20


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from figdraw_b200.cuda_context import CudaContext, FigDrawError

    with pytest.raises(FigDrawError) as e:
        CudaContext()
    assert e.value.code == abi.Status.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "figdraw_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "figdraw_oracle" not in text, f
