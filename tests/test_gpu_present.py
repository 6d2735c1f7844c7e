"""SURVEY 8f rank 4 -- present without a read-back: the framebuffer exported as a shareable handle is imported by a
SECOND PROCESS, whose bytes must equal what readPixels returns in the first."""
import os
import subprocess
import sys

import numpy as np
import pytest

from figdraw_b200 import scenes_synth as ss
from figdraw_b200.cuda_context import CudaContext, render_trace

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_exported_framebuffer_is_importable_by_another_process(tmp_path):
    tr = ss.config_trace(2, 1280, 720)
    want = render_trace(tr)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    fd, nbytes = ctx.exportFramebuffer(tr.width, tr.height)
    assert fd >= 0 and nbytes >= tr.width * tr.height * 4
    got = render_trace(tr, ctx)  # renders into the exported allocation
    assert np.array_equal(got, want)
    out = str(tmp_path / "imported.npy")
    r = subprocess.run([sys.executable, os.path.join(HERE, "import_fb_child.py"), str(fd), str(nbytes), str(tr.width), str(tr.height), out],
                       pass_fds=[fd], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(np.load(out), want)
    # a second frame (different content) is visible through the same import without any new export
    tr2 = ss.config_trace(5, 1280, 720, n_rects=2000, n_glyphs=400)
    ctx2 = CudaContext(atlasSize=tr2.atlas_size)
    fd2, nbytes2 = ctx2.exportFramebuffer(1280, 720)
    for k, t in enumerate((tr2, tr2)):
        img = render_trace(t, ctx2)
        out2 = str(tmp_path / f"imported2_{k}.npy")
        r = subprocess.run([sys.executable, os.path.join(HERE, "import_fb_child.py"), str(fd2), str(nbytes2), "1280", "720", out2],
                           pass_fds=[fd2], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        assert np.array_equal(np.load(out2), img)
    ctx.close()
    ctx2.close()


def test_exported_framebuffer_refuses_larger_frames():
    from figdraw_b200.cuda_context import FigDrawError

    ctx = CudaContext()
    ctx.exportFramebuffer(64, 64)  # rounded up to the allocation granularity (2 MiB)
    ctx.beginFrame((2048, 2048), clearMain=True)
    with pytest.raises(FigDrawError):
        ctx.endFrame()
    ctx.close()
