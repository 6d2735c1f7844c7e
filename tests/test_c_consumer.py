"""The drop-in boundary used from plain C (tests/c/abi_smoke.c): header + shared library only, no Python in the loop."""
import os
import subprocess

import pytest

from conftest import ROOT
from figdraw_b200 import abi


def _build(tmp_path):
    exe = tmp_path / "abi_smoke"
    lib_dir = os.path.dirname(abi.library_path())
    subprocess.run(["/usr/bin/gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                    "-L", lib_dir, "-lfigdraw_cuda", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    return str(exe)


def _run(exe):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout.splitlines()


def test_c_program_links_and_flattens(tmp_path):
    out = _run(_build(tmp_path))
    assert out[0] == f"abi {abi.ABI_VERSION}"
    # saveTransform, scale, drop shadow, fill, stroke, restoreTransform
    assert out[1] == "flatten rc 0 records 6 ops 1 5 32 32 32 2"
    assert out[2] == "flatten small buffer rc 4 needed 6"  # FDC_ERR_CAPACITY, size reported


@pytest.mark.gpu
def test_c_program_renders(tmp_path):
    out = _run(_build(tmp_path))
    px = [l for l in out if l.startswith("pixel(12,12)")]
    assert px, out
    v = [int(t) for t in px[0].replace("pixel(12,12)", "").replace("pixel(64,48)", "").split()]
    assert v[:4] == [255, 255, 255, 255] and v[4:] == [220, 40, 40, 255]
