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


def test_scene_pod_layouts_match_the_header(tmp_path):
    """The numpy/ctypes mirrors of the scene PODs (abi.py) against the C compiler's view of include/figdraw_cuda.h."""
    import subprocess

    from figdraw_b200 import abi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "figdraw_cuda.h"
#define S(t) printf(#t " %zu\n", sizeof(t))
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(fdc_call); S(fdc_node_fill); S(fdc_node_shadow); S(fdc_node_stroke); S(fdc_fig); S(fdc_glyph); S(fdc_draw_op);
  S(fdc_render_list); S(fdc_flatten_env); S(fdc_text_rect); S(fdc_scene);
  O(fdc_scene, glyphs); O(fdc_scene, text_rects); O(fdc_scene, ops); O(fdc_scene, points); O(fdc_fig, u.text.n_decoration);
  O(fdc_fig, flags); O(fdc_fig, parent); O(fdc_fig, child_count); O(fdc_fig, screen_box); O(fdc_fig, rotation); O(fdc_fig, fill);
  O(fdc_fig, corners); O(fdc_fig, corner_radii_y); O(fdc_fig, u);
  O(fdc_fig, u.rect.stroke); O(fdc_fig, u.drawable.steps); O(fdc_fig, u.drawable.first_op); O(fdc_fig, u.msdf.px_range);
  O(fdc_fig, u.transform.matrix); O(fdc_fig, u.transform.use_matrix);
  O(fdc_draw_op, center); O(fdc_draw_op, box); O(fdc_draw_op, start_angle); O(fdc_draw_op, first_point); O(fdc_draw_op, steps);
  O(fdc_render_list, root_ids); O(fdc_flatten_env, image_keys);
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    c = {k: int(v) for k, v in out.items()}
    assert c["fdc_call"] == abi.CALL_DTYPE.itemsize == 128
    assert c["fdc_node_fill"] == abi.NODE_FILL_DTYPE.itemsize
    assert c["fdc_node_shadow"] == abi.NODE_SHADOW_DTYPE.itemsize
    assert c["fdc_node_stroke"] == abi.NODE_STROKE_DTYPE.itemsize
    assert c["fdc_fig"] == abi.FIG_DTYPE.itemsize
    assert c["fdc_glyph"] == abi.GLYPH_DTYPE.itemsize
    assert c["fdc_draw_op"] == abi.DRAW_OP_DTYPE.itemsize
    import ctypes

    assert c["fdc_render_list"] == ctypes.sizeof(abi.FdcRenderList)
    assert c["fdc_flatten_env"] == ctypes.sizeof(abi.FdcFlattenEnv)
    f = abi.FIG_DTYPE.fields
    for name in ("flags", "parent", "child_count", "screen_box", "rotation", "fill", "corners", "corner_radii_y"):
        assert c[f"fdc_fig.{name}"] == f[name][1], name
    pay = f["payload"][1]
    assert c["fdc_fig.u"] == pay
    assert c["fdc_fig.u.rect.stroke"] == pay + abi.FIG_RECT_DTYPE.fields["stroke"][1]
    assert c["fdc_fig.u.drawable.steps"] == pay + abi.FIG_DRAWABLE_DTYPE.fields["steps"][1]
    assert c["fdc_fig.u.drawable.first_op"] == pay + abi.FIG_DRAWABLE_DTYPE.fields["first_op"][1]
    assert c["fdc_fig.u.msdf.px_range"] == pay + abi.FIG_MSDF_DTYPE.fields["px_range"][1]
    assert c["fdc_fig.u.transform.matrix"] == pay + abi.FIG_TRANSFORM_DTYPE.fields["matrix"][1]
    assert c["fdc_fig.u.transform.use_matrix"] == pay + abi.FIG_TRANSFORM_DTYPE.fields["use_matrix"][1]
    d = abi.DRAW_OP_DTYPE.fields
    for name in ("center", "box", "start_angle", "first_point", "steps"):
        assert c[f"fdc_draw_op.{name}"] == d[name][1], name
    assert c["fdc_text_rect"] == abi.TEXT_RECT_DTYPE.itemsize
    assert c["fdc_scene"] == ctypes.sizeof(abi.FdcScene)
    for name in ("glyphs", "text_rects", "ops", "points"):
        assert c[f"fdc_scene.{name}"] == getattr(abi.FdcScene, name).offset, name
    assert c["fdc_fig.u.text.n_decoration"] == pay + abi.FIG_TEXT_DTYPE.fields["n_decoration"][1]
    assert c["fdc_render_list.root_ids"] == abi.FdcRenderList.root_ids.offset
    assert c["fdc_flatten_env.image_keys"] == abi.FdcFlattenEnv.image_keys.offset


def test_nim_shim_declares_every_entry_point():
    """bindings/nim/cuda_context.nim is the reference-side binding a maintainer would add (not compilable here: no Nim):
    it must at least declare an importc proc for every function of the header, and mention no function the header lacks."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "figdraw_cuda.h")).read()
    nim = open(os.path.join(root, "bindings", "nim", "cuda_context.nim")).read()
    declared = set(re.findall(r"^(?:int|void|void\*|float|const char\*)\s+\*?(fdc_[a-z0-9_]+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 80
    in_nim = set(re.findall(r"^proc\s+(fdc_[a-z0-9_]+)\b", nim, flags=re.M))
    assert declared - in_nim == set(), f"no importc for {sorted(declared - in_nim)}"
    assert in_nim - declared == set(), f"the shim binds functions the header does not declare: {sorted(in_nim - declared)}"
