"""Glyph coverage rasterisation: the CPU oracle (oracle/glyph_oracle.c) against brute-force
supersampling and known areas (CPU tests), and the CUDA rasteriser against the oracle (GPU tests).

PARITY UNPINNED against pixie's anti-aliasing (pixie is not vendored): what is pinned here is that the oracle computes
exact area coverage, and that the CUDA path equals the oracle within 1 LSB of alpha."""
import numpy as np
import pytest

from figdraw_b200 import glyph_outlines as go
from oracle import oracle


def _flatten(segs, n=48):
    """Polylines of the contours (quadratics finely subdivided) for the brute-force reference."""
    lines = []
    for x0, y0, x1, y1, cx, cy, kind in segs:
        if kind == 0:
            lines.append((x0, y0, x1, y1))
        else:
            px, py = x0, y0
            for k in range(1, n + 1):
                t = k / n
                qx = (1 - t) ** 2 * x0 + 2 * (1 - t) * t * cx + t * t * x1
                qy = (1 - t) ** 2 * y0 + 2 * (1 - t) * t * cy + t * t * y1
                lines.append((px, py, qx, qy))
                px, py = qx, qy
    return np.array(lines, dtype=np.float64)


def _supersample(segs, w, h, ss=24):
    """Coverage by ss x ss point samples per pixel, non-zero winding of the flattened contours."""
    L = _flatten(segs)
    xs = (np.arange(w * ss) + 0.5) / ss
    ys = (np.arange(h * ss) + 0.5) / ss
    wind = np.zeros((h * ss, w * ss), dtype=np.int32)
    for x0, y0, x1, y1 in L:
        if y0 == y1:
            continue
        lo, hi = (y0, y1) if y0 < y1 else (y1, y0)
        rows = np.nonzero((ys >= lo) & (ys < hi))[0]
        if rows.size == 0:
            continue
        xint = x0 + (ys[rows] - y0) * (x1 - x0) / (y1 - y0)
        sgn = 1 if y1 > y0 else -1
        wind[rows[:, None], np.arange(w * ss)[None, :]] += sgn * (xs[None, :] > xint[:, None])
    inside = (wind != 0).astype(np.float64)
    return inside.reshape(h, ss, w, ss).mean(axis=(1, 3))


def test_oracle_known_areas():
    # an axis-aligned box with fractional edges: every pixel's coverage is known in closed form
    segs = go.to_array(go.polygon([(1.25, 2.5), (6.75, 2.5), (6.75, 7.25), (1.25, 7.25)]))
    img = oracle.rasterize_glyph(segs, 9, 10)
    a = img[..., 3].astype(np.float64) / 255.0
    assert abs(a.sum() - 5.5 * 4.75) < 0.02
    assert img[5, 4].tolist() == [255, 255, 255, 255] and img[0, 0].tolist() == [0, 0, 0, 0]
    assert abs(a[2, 1] - 0.75 * 0.5) < 0.003 and abs(a[7, 6] - 0.75 * 0.25) < 0.003
    # winding direction does not matter for a single contour, a hole wound the other way is empty
    rev = oracle.rasterize_glyph(go.to_array(go.polygon([(1.25, 2.5), (6.75, 2.5), (6.75, 7.25), (1.25, 7.25)], reverse=True)), 9, 10)
    assert np.array_equal(rev, img)
    ring = go.polygon([(1, 1), (11, 1), (11, 11), (1, 11)]) + go.polygon([(4, 4), (8, 4), (8, 8), (4, 8)], reverse=True)
    r = oracle.rasterize_glyph(go.to_array(ring), 12, 12)[..., 3]
    assert r[6, 6] == 0 and r[2, 2] == 255 and abs(r.astype(np.float64).sum() / 255.0 - (100 - 16)) < 0.05
    # a circle's area
    c = oracle.rasterize_glyph(go.to_array(go.ellipse(16.3, 15.6, 11.0, 11.0, n_arcs=16)), 32, 32)[..., 3]
    # (chords of the flattened arcs, 0.025 px tolerance, cut ~0.5 px^2 off; the rest is the 8-bit rounding of ~100 edge texels)
    assert abs(c.astype(np.float64).sum() / 255.0 - np.pi * 121.0) < 1.5


def test_oracle_matches_supersampling():
    for w, h, segs in go.sample_glyphs(seed=3, count=6):
        got = oracle.rasterize_glyph(go.to_array(segs), w, h)[..., 3].astype(np.float64) / 255.0
        want = _supersample(segs, w, h)
        d = np.abs(got - want)
        # point sampling is the coarser of the two: a near-vertical edge moves its estimate in steps of 1/24
        # (and where two overlapping contours' edges cross inside one pixel, accumulation adds their coverages where the
        # non-zero rule takes the union -- a known property of this family of rasterisers, a few percent in single texels)
        assert d.max() <= 0.08 and d.mean() <= 0.004, (w, h, d.max(), d.mean())


def test_oracle_lcd_filter_is_the_reference_integer_filter():
    w, h, segs = go.sample_glyphs(seed=5, count=1)[0]
    plain = oracle.rasterize_glyph(go.to_array(segs), w, h)[..., 3].astype(np.int32)
    lcd = oracle.rasterize_glyph(go.to_array(segs), w, h, lcd_filter=True)
    wts = [8, 77, 86, 77, 8]
    want = np.zeros_like(plain)
    for x in range(w):
        acc = np.zeros(h, dtype=np.int32)
        for i, wt in enumerate(wts):
            acc += plain[:, min(max(x + i - 2, 0), w - 1)] * wt
        want[:, x] = (acc + 128) >> 8
    assert np.array_equal(lcd[..., 3].astype(np.int32), want)
    assert np.array_equal(lcd[..., 0] == 255, want > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("lcd", [False, True])
def test_cuda_glyph_rasteriser_equals_the_oracle(lcd):
    from figdraw_b200.cuda_context import CudaContext
    from figdraw_b200.figbackend import TraceBackend

    glyphs = go.sample_glyphs(seed=7, count=24) + go.sample_glyphs(seed=8, count=4, size=(150, 210))  # incl. multi-strip bitmaps
    jobs, segs, keys = go.jobs_for(glyphs)
    # same atlas size and insertion order on both sides: the packer then places every glyph identically, which matters for
    # the minified draws (the reference builds mip levels from the slot's own origin, so their phase follows its parity)
    ctx = CudaContext(atlasSize=2048)
    ctx.rasterizeGlyphs(jobs, segs, lcdFilter=lcd)
    assert all(ctx.hasImage(k) for k in keys) and ctx.atlasUsage().glyph_count == len(keys)
    # draw every glyph 1:1 on black with a white tint: the frame then shows the atlas texels themselves
    W, H = 1024, 768
    tb = TraceBackend(atlasSize=2048)
    bitmaps = {}
    for (w, h, s), key in zip(glyphs, keys):
        bitmaps[key] = oracle.rasterize_glyph(go.to_array(s), w, h, lcd_filter=lcd)
        tb.putImage(key, bitmaps[key])
    tb.beginFrame((W, H), clearMain=True, clearMainColor=(0.0, 0.0, 0.0, 1.0))
    x = y = 4
    row_h = 0
    for (w, h, _s), key in zip(glyphs, keys):
        if x + w + 4 > W:
            x, y, row_h = 4, y + row_h + 4, 0
        tb.drawImage(key, (float(x), float(y)), [0xFFFFFFFF] * 4)
        tb.drawImage(key, (float(x), float(y) + 380.0), [0xFF40C0FF] * 4, (w * 0.5, h * 0.5))  # minified: uses the mip chain
        x += w + 4
        row_h = max(row_h, h)
    tb.endFrame()
    tr = tb.trace()
    ctx.beginFrame((W, H), clearMain=True, clearMainColor=(0.0, 0.0, 0.0, 1.0))
    ctx.submitCalls(tr.calls)
    ctx.endFrame()
    got = ctx.readPixels()
    want = oracle.render_trace(tr)
    d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
    assert int(d.max()) <= 2, f"max {int(d.max())} LSB at {np.argwhere(d == d.max())[0]}"
    assert got[..., :3].max() > 200 and ctx.missing_images == 0
    ctx.close()


@pytest.mark.gpu
def test_cuda_glyph_batch_survives_an_atlas_regrow():
    """The atlas doubles in the middle of a batch (native replay on): every glyph of the batch, placed before or after the
    regrow, must still come out right (1:1 draws: texel-exact positions do not depend on where the packer put them)."""
    from figdraw_b200.cuda_context import CudaContext
    from figdraw_b200.figbackend import TraceBackend

    glyphs = go.sample_glyphs(seed=11, count=40, size=(40, 52))
    jobs, segs, keys = go.jobs_for(glyphs)
    ctx = CudaContext(atlasSize=256)
    ctx.setAtlasReplay(True)
    assert ctx.rasterizeGlyphs(jobs, segs) is True and ctx.atlasSize() > 256
    assert all(ctx.hasImage(k) for k in keys)
    W, H = 1024, 400
    tb = TraceBackend(atlasSize=2048)
    for (w, h, s), key in zip(glyphs, keys):
        tb.putImage(key, oracle.rasterize_glyph(go.to_array(s), w, h))
    tb.beginFrame((W, H), clearMain=True, clearMainColor=(0.0, 0.0, 0.0, 1.0))
    x = y = 3
    for (w, h, _s), key in zip(glyphs, keys):
        if x + w + 3 > W:
            x, y = 3, y + 62
        tb.drawImage(key, (float(x), float(y)), [0xFFFFFFFF] * 4)
        x += w + 3
    tb.endFrame()
    tr = tb.trace()
    ctx.beginFrame((W, H), clearMain=True, clearMainColor=(0.0, 0.0, 0.0, 1.0))
    ctx.submitCalls(tr.calls)
    ctx.endFrame()
    d = np.abs(ctx.readPixels().astype(np.int16) - oracle.render_trace(tr).astype(np.int16)).max(axis=2)
    assert int(d.max()) <= 2 and ctx.missing_images == 0
    ctx.close()
