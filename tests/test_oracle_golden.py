"""The oracle against every golden vector the reference's tests hold for this path.

tests/golden/*.png are the reference's own tests/expected/*.png (see tests/golden/import_goldens.py).
Measured when the oracle was written: max |diff| 1 LSB on all six, three of them bit-exact.
"""
import numpy as np
import pytest

from figdraw_b200 import scenes
from oracle import oracle

# name -> (max abs diff allowed, max number of pixels that may differ at all)
EXPECT = {
    "rgb_boxes_sdf": (1, 4000),
    "linear_gradient": (1, 5500),
    "layers_clip": (0, 0),
    "circle_rect": (0, 0),
    "line_rect": (0, 0),
    "image": (1, 2500),
}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_oracle_matches_reference_golden(name):
    img = oracle.render_trace(scenes.golden_trace(name))
    gold = scenes.load_golden(name)
    assert img.shape == gold.shape
    d = np.abs(img.astype(np.int16) - gold.astype(np.int16))
    max_diff, max_px = EXPECT[name]
    assert int(d.max()) <= max_diff
    assert int((d.max(axis=2) > 0).sum()) <= max_px


def test_oracle_is_thread_count_independent():
    tr = scenes.golden_trace("rgb_boxes_sdf")
    a = oracle.render_trace(tr, n_threads=1)
    b = oracle.render_trace(tr, n_threads=7)
    assert np.array_equal(a, b)


def test_reference_spot_pixels():
    """Spot checks the reference tests assert (trender_layers_clip.nim:272-289, trender_linear_gradient.nim:124-138)."""
    img = oracle.render_trace(scenes.golden_trace("linear_gradient"))

    def close(px, rgb, tol):
        return max(abs(int(px[i]) - rgb[i]) for i in range(3)) <= tol

    assert close(img[140, 120], (220, 40, 40), 40)
    assert close(img[140, 300], (40, 200, 90), 40)
    assert close(img[140, 480], (50, 90, 225), 40)
    assert close(img[270, 190], (240, 210, 40), 40)
    assert close(img[430, 190], (110, 60, 210), 40)
    assert int(img[252, 365][0]) > int(img[252, 365][2]) + 40
    assert int(img[252, 555][2]) > int(img[252, 555][0]) + 40
    # rect-mask variant renders the same picture as the clip variant within 1 % (trender_layers_clip.nim:322-325)
    clip = oracle.render_trace(scenes.golden_trace("layers_clip")).astype(np.int16)
    rm = oracle.render_trace(scenes.trace_scene(scenes.layers_rect_mask, 800, 375)).astype(np.int16)
    assert np.abs(clip - rm).sum() / (255.0 * clip.size) < 0.01
    # mixed rect-mask batch (trender_layers_clip.nim:350-353)
    mix = oracle.render_trace(scenes.trace_scene(scenes.mixed_rect_mask_batch, 480, 180))
    assert close(mix[88, 74], (230, 70, 52), 12)
    assert close(mix[88, 160], (255, 255, 255), 12)
    assert close(mix[88, 204], (56, 168, 88), 12)
    assert close(mix[88, 336], (54, 118, 230), 12)


def test_fuzz_streams_are_deterministic_on_the_oracle():
    from figdraw_b200 import scenes_fuzz

    for seed in (0, 5):
        tr = scenes_fuzz.random_trace(seed)
        a = oracle.render_trace(tr, n_threads=1)
        b = oracle.render_trace(tr, n_threads=5)
        assert np.array_equal(a, b)
        c1 = oracle.count_fragments(tr)
        _img, c2 = oracle.render_trace(tr, want_counts=True)
        assert np.array_equal(c1, c2)


def _max_channel_delta(px, r, g, b):
    return max(abs(int(px[0]) - r), abs(int(px[1]) - g), abs(int(px[2]) - b))


def oneframe_scene():
    """tests/tfigrender_oneframe_screenshot.nim:20-42: white root + one red rect, 240x160."""
    from figdraw_b200.fignodes import Fig, FigKind, RenderList, Renders, fill, rect, rgba

    lst = RenderList()
    lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(0, 0, 240, 160), fill=fill(rgba(255, 255, 255, 255))))
    lst.addRoot(Fig(kind=FigKind.nkRectangle, screenBox=rect(32, 24, 120, 80), fill=fill(rgba(220, 40, 40, 255))))
    r = Renders()
    r.setLayer(0, lst)
    return r


def check_reference_spot_pixels(render):
    """The spot pixels the reference's render tests assert (img[x, y], tolerances as in the Nim tests)."""
    from figdraw_b200.scenes_synth import trace_renders

    img = render(trace_renders(oneframe_scene(), 240, 160))
    assert img.shape[:2] == (160, 240)
    assert _max_channel_delta(img[12, 12], 255, 255, 255) <= 12   # tfigrender_oneframe_screenshot.nim:90-92
    assert _max_channel_delta(img[48, 64], 220, 40, 40) <= 12
    img = render(scenes.golden_trace("linear_gradient"))            # trender_linear_gradient.nim:124-138
    for (x, y), rgb in {(120, 140): (220, 40, 40), (300, 140): (40, 200, 90), (480, 140): (50, 90, 225),
                        (190, 270): (240, 210, 40), (190, 430): (110, 60, 210)}.items():
        assert _max_channel_delta(img[y, x], *rgb) <= 40, (x, y)
    left, right = img[252, 365], img[252, 555]
    assert int(left[0]) > int(left[2]) + 40 and int(right[2]) > int(right[0]) + 40
    left, right = img[400, 602], img[400, 768]
    assert int(left[0]) > int(left[2]) + 20 and int(right[2]) > int(right[0]) + 20


def test_oracle_reference_spot_pixels():
    check_reference_spot_pixels(oracle.render_trace)
