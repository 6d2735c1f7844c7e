"""Failure paths of the CUDA backend: bin-list overflow in any segment, re-runs of frames that blend over the previous
pixels, the cross-rank retry protocol of a tile-band partition, aborted frames, deep mask nesting and mask levels
holding several draws.  Everything is compared with the CPU oracle or with an undisturbed context on the same input."""
import numpy as np
import pytest

from figdraw_b200 import scenes_synth as ss
from figdraw_b200.abi import Op, SdfMode
from figdraw_b200.cuda_context import CudaContext, FigDrawError, render_trace
from figdraw_b200.figbackend import TraceBackend, ZeroRadii, circularRadii, solid
from figdraw_b200.fignodes import rgba
from oracle import oracle

pytestmark = pytest.mark.gpu


def _submit(c, tr, calls=None, clear=True):
    c.beginFrame((tr.width, tr.height), clearMain=clear and tr.clear is not None, clearMainColor=tr.clear or (1.0, 1.0, 1.0, 1.0))
    c.submitCalls(tr.calls if calls is None else calls)
    c.endFrame()


def _blur_sandwich(width=640, height=400, n_before=40, n_after=900, seed=11):
    """A few rects, a backdrop blur, then many translucent rects: the segment AFTER the blur needs the longer lists."""
    rng = np.random.default_rng(seed)
    tb = TraceBackend()
    tb.beginFrame((width, height), clearMain=True, clearMainColor=(0.95, 0.93, 0.9, 1.0))

    def rects(n):
        for _ in range(n):
            x, y = rng.uniform(-20, width - 30), rng.uniform(-20, height - 30)
            w, h = rng.uniform(20, 160), rng.uniform(14, 120)
            col = rgba(int(rng.integers(256)), int(rng.integers(256)), int(rng.integers(256)), int(rng.integers(60, 200)))
            tb.drawRoundedRectSdf((x, y, w, h), solid(col), circularRadii((6, 3, 9, 0)))

    rects(n_before)
    tb.drawBackdropBlur((120.0, 80.0, 300.0, 200.0), circularRadii((12, 12, 12, 12)), 14.0)
    rects(n_after)
    tb.endFrame()
    return tb.trace()


def test_overflow_in_the_first_segment_of_a_blur_frame():
    """an overflow in a segment that is not the last used to be forgotten when the next segment
    reset the counters -- the segment's content went missing for good.  The flags are sticky per frame now."""
    tr = ss.config_trace(2, 1280, 720)
    want = render_trace(tr)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    for _i, key, img in tr.images:
        ctx.putImage(key, img)
    _submit(ctx, tr)
    ctx.sync()
    n_seg = ctx.frameStats().n_segments
    assert n_seg >= 2
    for coarse, tile in ((0, 2000), (64, 0), (64, 2000)):
        ctx.debugLimitLists(coarse, tile)  # segment 0 (~700 primitives) overflows, the short last segment does not
        _submit(ctx, tr)
        got = ctx.readPixels()
        assert np.array_equal(got, want), f"limits {coarse}/{tile}"
    # n_tile_entries is the whole frame's, not the last segment's
    total = sum(len(ent) for _off, ent in (ctx.debugBins(s) for s in range(n_seg)))
    _submit(ctx, tr)
    assert ctx.frameStats().n_tile_entries == total
    ctx.close()


def test_overflow_replay_does_not_composite_twice_without_clear():
    """clearMain=false + overflow in a LATER segment: the earlier segments were already blended when
    the frame is re-run; the pre-frame pixels are restored first."""
    tr = _blur_sandwich()
    base = ss.config_trace(5, tr.width, tr.height, n_rects=300, n_glyphs=0)

    def run(limit):
        ctx = CudaContext(atlasSize=base.atlas_size)
        for _i, key, img in base.images:
            ctx.putImage(key, img)
        _submit(ctx, base)  # frame 1: something to blend over
        _submit(ctx, tr, clear=False)  # sizes the lists
        ctx.sync()
        _submit(ctx, base)
        if limit:
            ctx.debugLimitLists(0, limit)
        _submit(ctx, tr, clear=False)
        out = ctx.readPixels().copy()
        ctx.close()
        return out

    want = run(0)
    seg0 = oracle.reference_bins(tr)[0][1].size
    seg1 = oracle.reference_bins(tr)[1][1].size
    assert seg1 > 2 * seg0  # a limit between the two overflows only the segment after the blur
    got = run((seg0 + seg1) // 2)
    assert np.array_equal(got, want)


def test_banded_overflow_asks_every_rank_to_retry():
    """under a tile-band partition a private re-run would show the neighbours' later state in the blur
    halo.  The rank reports FDC_ERR_RETRY instead; every rank re-runs the frame (fdc_retry_frame)."""
    from figdraw_b200.bands import padded_rows

    tr = ss.config_trace(2, 1280, 720)
    full = render_trace(tr)
    n = 2
    ctxs = [CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=n) for r in range(n)]
    try:
        for c in ctxs:
            c.reserveFramebuffer(tr.width, padded_rows(tr.height, n))
        ptrs = [c.framebufferPtr() for c in ctxs]
        for c in ctxs:
            c.setPeerFramebuffers(ptrs)
            for _idx, key, img in tr.images:
                c.putImage(key, img)
        for c in ctxs:  # size the buffers one rank at a time, without cross-rank waits
            _submit(c, tr, tr.calls[tr.calls["op"] != Op.BACKDROP_BLUR])
            c.sync()
        for c in ctxs:
            _submit(c, tr)
        for c in ctxs:
            c.sync()
        ctxs[1].debugLimitLists(0, 3000)
        for c in ctxs:
            _submit(c, tr)
        status = [c.syncStatus() for c in ctxs]
        assert status == [0, 6]  # FDC_ERR_RETRY on the rank whose lists overflowed
        for c in ctxs:
            c.retryFrame()
        assert [c.syncStatus() for c in ctxs] == [0, 0]
        for r, c in enumerate(ctxs):
            assert np.array_equal(c.readPixels(), full), f"rank {r}"
    finally:
        for c in ctxs:
            c.close()


def test_a_failed_frame_does_not_wedge_the_context():
    """an error between beginFrame and endFrame left frame_begun set for good."""
    tr = ss.config_trace(5, 640, 360, n_rects=500, n_glyphs=0)
    want = render_trace(tr)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    for _i, key, img in tr.images:
        ctx.putImage(key, img)
    ctx.beginFrame((tr.width, tr.height), clearMain=True)
    bad = tr.calls[:40].copy()
    bad[20]["op"] = 99  # unknown record
    with pytest.raises(FigDrawError):
        ctx.submitCalls(bad)
    with pytest.raises(FigDrawError, match="already"):
        ctx.beginFrame((tr.width, tr.height), clearMain=True)
    ctx.abortFrame()
    assert np.array_equal(render_trace(tr, ctx), want)
    ctx.close()


def test_native_render_frame_aborts_itself_on_error():
    """fdc_render_frame with clip nesting beyond the limit fails -- and the next frame renders normally."""
    from figdraw_b200 import fignodes as fn, native_scene

    def nested(depth):
        r = fn.newRenders()
        parent = None
        for d in range(depth):
            node = fn.Fig(kind=fn.FigKind.nkRectangle, screenBox=fn.rect(4.0 + 3 * d, 4.0 + 2 * d, 240.0 - 6 * d, 200.0 - 4 * d),
                          fill=fn.fill(rgba(20 + 12 * d, 200 - 9 * d, 90, 255)), flags=fn.FigFlags.NfClipContent)
            parent = r.addRoot(0, node) if parent is None else r.addChild(0, parent, node)
        return r

    ctx = CudaContext()
    with pytest.raises(FigDrawError) as e:
        ctx.renderFrameNative(native_scene.pack_renders(nested(17)), (256, 208))
    assert e.value.code == 4
    ok = nested(15)
    ctx.renderFrameNative(native_scene.pack_renders(ok), (256, 208))
    got = ctx.readPixels()
    from figdraw_b200 import figrender
    from figdraw_b200.figbackend import TraceBackend as TB

    tb = TB()
    figrender.setFigUiScale(1.0)
    figrender.renderFrame(tb, ok, (256.0, 208.0))
    want = oracle.render_trace(tb.trace())
    assert int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) <= 2
    ctx.close()


def test_fifteen_nested_mask_levels():
    """GL nests mask textures without a limit (glcontext.nim:171-201); levels 9..15 live in shared memory here."""
    tb = TraceBackend()
    tb.beginFrame((320, 300), clearMain=True)
    for d in range(15):
        tb.beginMask((6.0 + 7 * d, 5.0 + 6 * d, 300.0 - 13 * d, 288.0 - 11 * d), circularRadii((20, 6, 12, 0)))
        tb.endMask()
        tb.drawRoundedRectSdf((0.0, 0.0, 320.0, 300.0), solid(rgba(16 * d, 255 - 15 * d, 90, 130)), ZeroRadii)
    for d in range(15):
        tb.popMask()
        tb.drawRoundedRectSdf((10.0 * d, 0.0, 9.0, 300.0), solid(rgba(200, 10 * d, 30, 90)), ZeroRadii)
    tb.endFrame()
    tr = tb.trace()
    got, want = render_trace(tr), oracle.render_trace(tr)
    assert int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) <= 2
    ctx = CudaContext()
    render_trace(tr, ctx)
    for seg, (off_ref, ent_ref) in enumerate(oracle.reference_bins(tr)):
        off, ent = ctx.debugBins(seg)
        assert np.array_equal(off, off_ref) and np.array_equal(ent, ent_ref)
    ctx.beginFrame((64, 64), clearMain=True)
    for _ in range(15):
        ctx.beginMask((0, 0, 64, 64), ZeroRadii)
        ctx.endMask()
    with pytest.raises(FigDrawError) as e:
        ctx.beginMask((0, 0, 64, 64), ZeroRadii)
    assert e.value.code == 4  # FDC_ERR_CAPACITY
    ctx.abortFrame()
    ctx.close()


def test_mask_level_with_several_draws_is_cleared_everywhere():
    """GL clears the whole mask texture at beginMask.  A level holding two shapes (content under it is
    not clipped to one bbox) drawn after another mask of the same depth must not see that one's stale values; a dropped
    (zero-sized) clip rect followed by a real shape clears too."""
    for variant in range(3):
        tb = TraceBackend()
        tb.beginFrame((400, 300), clearMain=True, clearMainColor=(0.9, 0.9, 0.95, 1.0))
        # level 1, used once and popped: leaves values behind at depth 1
        tb.beginMask((30.0, 20.0, 340.0, 260.0), circularRadii((30, 30, 30, 30)))
        tb.endMask()
        tb.drawRoundedRectSdf((0.0, 0.0, 400.0, 300.0), solid(rgba(230, 60, 40, 120)), ZeroRadii)
        tb.popMask()
        # level 1 again, built from two shapes that do not cover the first mask's area
        if variant == 2:
            tb.beginMask((50.0, 50.0, 0.0, 40.0), ZeroRadii)  # dropped: zero width
        else:
            tb.beginMask((50.0, 40.0, 90.0, 70.0), circularRadii((12, 0, 12, 0)))
        tb.drawRoundedRectSdf((220.0, 150.0, 120.0, 100.0), solid(rgba(255, 255, 255, 255)), circularRadii((20, 20, 20, 20)))
        if variant == 1:
            tb.drawQuadraticBezierSdf((100.0, 100.0, 200.0, 120.0), solid(rgba(255, 255, 255, 255)), (-80.0, -40.0), (0.0, 50.0),
                                      (80.0, -30.0), 9.0, 1)
        tb.endMask()
        tb.drawRoundedRectSdf((0.0, 0.0, 400.0, 300.0), solid(rgba(20, 70, 220, 200)), ZeroRadii)
        tb.popMask()
        tb.endFrame()
        tr = tb.trace()
        got, want = render_trace(tr), oracle.render_trace(tr)
        d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
        assert int(d.max()) <= 2, f"variant {variant}: max {int(d.max())} LSB at {np.argwhere(d == d.max())[0]}"
        ctx = CudaContext()
        render_trace(tr, ctx)
        for seg, (off_ref, ent_ref) in enumerate(oracle.reference_bins(tr)):
            off, ent = ctx.debugBins(seg)
            assert np.array_equal(off, off_ref) and np.array_equal(ent, ent_ref), f"variant {variant}: bins differ"
        ctx.close()


def test_pixelate_context_magnifies_with_nearest():
    """`newContext(pixelate = true)` (glcontext.nim:165-168): GL_NEAREST magnification of the atlas -- magnified images,
    1:1 glyphs and MSDF quads (textureLod(.., 0) goes through the magnification filter); minified images stay trilinear."""
    from figdraw_b200 import scenes, scenes_fuzz

    def magnified():
        rng = np.random.default_rng(5)
        tb = TraceBackend(atlasSize=512)
        photo = rng.integers(0, 256, size=(48, 64, 4), dtype=np.uint8)
        tb.putImage(900, photo)
        tb.beginFrame((640, 400), clearMain=True, clearMainColor=(0.1, 0.1, 0.1, 1.0))
        tb.drawImage(900, (10.0, 8.0), [0xFFFFFFFF] * 4, (256.0, 192.0))           # x4
        tb.drawImage(900, (280.5, 10.25), [0xFFFFFFFF] * 4, (64.0, 48.0))          # 1:1 at a fractional position
        tb.drawImage(900, (290.0, 90.0), [0xFF80FFFF] * 4, (320.0, 240.0), True)   # flipped, x5
        tb.drawImage(900, (20.0, 220.0), [0xFFFFFFFF] * 4, (192.0, 144.0))         # x3
        tb.drawImage(900, (230.0, 340.0), [0xFFFFFFFF] * 4, (30.0, 20.0))          # minified: stays trilinear
        tb.endFrame()
        return tb.trace()

    # Left out on purpose, because GL_NEAREST makes them implementation-defined: rotated quads at ~1 texel per pixel
    # (lambda sits on 0; which side -- NEAREST or trilinear -- is decided by the last ulp) and non-integer magnifications
    # (some pixel centres map EXACTLY onto a texel boundary, e.g. 64 texels over 237 pixels at pixel 118).
    traces = [ss.config_trace(3, 1280, 720, n_glyphs=1500, msdf_glyphs=300), magnified()]
    n_changed = 0
    for tr in traces:
        ctx = CudaContext(atlasSize=tr.atlas_size, pixelate=True)
        got = render_trace(tr, ctx)
        ctx.close()
        want = oracle.render_trace(tr, pixelate=True)
        d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
        # GL_NEAREST is discontinuous: a pixel centre that maps onto a texel boundary picks one side or the other on the
        # last ulp of the coordinate (FMA here, separate multiply-add in the oracle), so a handful of boundary pixels may
        # land on the neighbouring texel; everything else must meet the usual 2 LSB.
        outliers = float((d > 2).mean())
        assert outliers <= 2e-4, f"{tr.width}x{tr.height}: {outliers:.5%} of pixels differ by more than 2 LSB (max {int(d.max())})"
        assert float((d > 0).mean()) <= 0.03
        n_changed += int((want != oracle.render_trace(tr)).any())
    assert n_changed >= 2  # the filter really changes these frames


def test_atlas_residency_eviction_and_native_replay_on_regrow():
    """entry kinds, owner tokens, eviction, and the atlas doubling WITHOUT losing the live images
    (fdc_set_atlas_replay): every image drawn after the regrow still matches the oracle, evicted ones are gone and their
    space is reclaimed."""
    rng = np.random.default_rng(21)
    imgs = {}
    for k in range(60):
        w, h = int(rng.integers(8, 40)), int(rng.integers(8, 40))
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        img[..., 3] = 255
        imgs[5000 + k] = img
    ctx = CudaContext(atlasSize=128)
    ctx.setAtlasReplay(True)
    keys = list(imgs)
    grew = 0
    for i, key in enumerate(keys[:40]):
        _rect, rebuilt = ctx.putImage(key, imgs[key])
        grew += int(rebuilt)
        if i < 20:
            ctx.markEntry(key, 2, idA=7 if i % 2 else 8, idB=3)  # glyphs of font 7 / 8, typeface 3
        else:
            ctx.markEntry(key, 1, idA=key)
            ctx.retainOwner(0, key, 111)
            ctx.retainOwner(0, key, 222)
    assert grew >= 1 and ctx.atlasSize() > 128
    u = ctx.atlasUsage()
    assert (u.entry_count, u.glyph_count, u.image_count) == (40, 20, 20) and u.rebuild_count >= 1
    assert all(ctx.hasImage(k) for k in keys[:40])  # nothing was dropped by the regrow
    # eviction: font 7's glyphs, and an image whose last owner lets go
    assert ctx.clearFontGlyphs(7) == 10
    assert not ctx.releaseOwner(0, keys[25], 111) and ctx.hasImage(keys[25])
    assert ctx.releaseOwner(0, keys[25], 222) and not ctx.hasImage(keys[25])
    assert ctx.clearTypefaceGlyphs(3) == 10
    live = [k for k in keys[20:40] if k != keys[25]]
    assert ctx.atlasUsage().entry_count == len(live) == 19
    # more images: the next regrow carries only the live ones over
    for key in keys[40:]:
        ctx.putImage(key, imgs[key])
        live.append(key)
    assert all(ctx.hasImage(k) for k in live) and not ctx.hasImage(keys[0])

    # draw every live image 1:1 and magnified; the oracle gets the same images (its packer places them elsewhere, the
    # pixels must not care)
    tb = TraceBackend(atlasSize=1024)
    for key in live:
        tb.putImage(key, imgs[key])
    tb.beginFrame((640, 480), clearMain=True, clearMainColor=(0.2, 0.2, 0.25, 1.0))
    for n, key in enumerate(live):
        x, y = 8 + (n % 10) * 62, 6 + (n // 10) * 90
        tb.drawImage(key, (float(x), float(y)), [0xFFFFFFFF] * 4)
        tb.drawImage(key, (float(x), float(y) + 44.0), [0xFFFFFFFF] * 4, (50.0, 40.0))
    tb.endFrame()
    tr = tb.trace()
    ctx.beginFrame((640, 480), clearMain=True, clearMainColor=(0.2, 0.2, 0.25, 1.0))
    ctx.submitCalls(tr.calls)
    ctx.endFrame()
    got = ctx.readPixels()
    want = oracle.render_trace(tr)
    d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
    assert int(d.max()) <= 2, f"max {int(d.max())} LSB"
    assert ctx.missing_images == 0
    ctx.close()
