"""Analytic properties of the oracle on the parts of the path no reference golden reaches (backdrop blur, MSDF):
what the GLSL text implies independently of any particular image (blur.frag:1-32, atlas.frag:294-318)."""
import numpy as np
import pytest

from figdraw_b200.figbackend import TraceBackend, circularRadii, solid
from figdraw_b200.fignodes import rgba
from oracle import oracle

ZERO = circularRadii((0, 0, 0, 0))


def _frame(w, h, body, clear=(0.2, 0.5, 0.8, 1.0)):
    tb = TraceBackend()
    tb.beginFrame((w, h), clearMain=True, clearMainColor=clear)
    body(tb)
    tb.endFrame()
    return tb.trace()


@pytest.mark.parametrize("radius", [0.3, 2.0, 9.0, 40.0, 64.0])
def test_blur_of_a_uniform_backdrop_is_the_backdrop(radius):
    """Normalised weights: every tap sees the same colour, so the composite leaves a flat frame flat (to 1 LSB of the
    two RGBA8 passes)."""
    tr = _frame(200, 120, lambda tb: tb.drawBackdropBlur((30.0, 20.0, 120.0, 70.0), circularRadii((12, 12, 12, 12)), radius))
    img = oracle.render_trace(tr)
    flat = np.array([51, 128, 204, 255], dtype=np.int16)  # round(0.2, 0.5, 0.8, 1.0 * 255)
    assert int(np.abs(img.astype(np.int16) - flat).max()) <= 1


def test_blur_below_half_a_pixel_is_a_copy_and_blur_is_mirror_symmetric():
    def scene(tb, radius):
        tb.drawRoundedRectSdf((60.0, 30.0, 80.0, 60.0), solid(rgba(240, 30, 30, 255)), ZERO)   # centred in 200 x 120
        tb.drawRoundedRectSdf((90.0, 10.0, 20.0, 100.0), solid(rgba(20, 200, 60, 255)), ZERO)
        if radius > 0:
            tb.drawBackdropBlur((20.0, 10.0, 160.0, 100.0), ZERO, radius)

    sharp = oracle.render_trace(_frame(200, 120, lambda tb: scene(tb, 0.0)))
    copy = oracle.render_trace(_frame(200, 120, lambda tb: scene(tb, 0.4)))     # blur.frag:12-16: r <= 0.5 -> texture()
    assert np.array_equal(sharp, copy)
    blurred = oracle.render_trace(_frame(200, 120, lambda tb: scene(tb, 12.0)))
    assert not np.array_equal(sharp, blurred)
    assert int(np.abs(blurred.astype(np.int16) - blurred[:, ::-1].astype(np.int16)).max()) <= 1   # left-right symmetric scene
    assert int(np.abs(blurred.astype(np.int16) - blurred[::-1].astype(np.int16)).max()) <= 1       # and top-bottom
    # blurring spreads the red box: a pixel just outside it gets redder, its centre stays red
    assert blurred[60, 55, 0] > sharp[60, 55, 0] and blurred[60, 100, 1] < sharp[60, 100, 1]


def test_msdf_field_extremes():
    """median(rgb) = 1 everywhere -> coverage 1 inside the quad (the fill colour); = 0 -> nothing drawn; the annular
    variant draws nothing at distance >> stroke width."""
    inside = np.full((16, 16, 4), 255, np.uint8)
    outside = np.zeros((16, 16, 4), np.uint8)
    outside[..., 3] = 255

    def scene(tb, img, stroke=0.0):
        tb.putImage(77, img)
        tb.drawMsdfImage(77, (40.0, 30.0), rgba(10, 200, 90, 255), (64.0, 48.0), 4.0, 0.5, stroke, False)

    a = oracle.render_trace(_frame(160, 120, lambda tb: scene(tb, inside)))
    assert tuple(a[54, 72]) == (10, 200, 90, 255) and tuple(a[5, 5]) == (51, 128, 204, 255)
    assert (a[31:77, 41:103, :3] == (10, 200, 90)).all()
    b = oracle.render_trace(_frame(160, 120, lambda tb: scene(tb, outside)))
    assert (b[..., :3] == (51, 128, 204)).all()
    c = oracle.render_trace(_frame(160, 120, lambda tb: scene(tb, inside, stroke=2.0)))
    assert (c[35:73, 45:99, :3] == (51, 128, 204)).all()   # deep inside the shape: far from the contour, no stroke


@pytest.mark.gpu
def test_cuda_backend_on_the_property_scenes():
    """The same scenes through the CUDA backend: within the parity tolerance of the oracle (covers the blur pass-through
    path r <= 0.5 and the largest tap reach r = 64)."""
    from figdraw_b200.cuda_context import render_trace

    traces = [_frame(200, 120, lambda tb, r=r: tb.drawBackdropBlur((30.0, 20.0, 120.0, 70.0), circularRadii((12, 12, 12, 12)), r))
              for r in (0.3, 2.0, 9.0, 40.0, 64.0)]

    def scene(tb, radius):
        tb.drawRoundedRectSdf((60.0, 30.0, 80.0, 60.0), solid(rgba(240, 30, 30, 255)), ZERO)
        tb.drawRoundedRectSdf((90.0, 10.0, 20.0, 100.0), solid(rgba(20, 200, 60, 255)), ZERO)
        tb.drawBackdropBlur((20.0, 10.0, 160.0, 100.0), ZERO, radius)

    traces += [_frame(200, 120, lambda tb, r=r: scene(tb, r)) for r in (0.4, 12.0, 64.0)]
    for tr in traces:
        got, want = render_trace(tr), oracle.render_trace(tr)
        assert int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) <= 2
