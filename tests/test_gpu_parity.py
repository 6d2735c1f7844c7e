"""Parity of the CUDA backend (through the C ABI) against the CPU oracle and the reference's golden images.

Tolerance: per-channel RGBA8 within +-2 LSB (BASELINE.json north_star); bin lists and draw order bit-exact.
The shading arithmetic is float32 on both sides but the kernel uses FMA contraction and MUFU approximations,
so a pixel whose exact value sits on a rounding boundary may land one LSB away; MAX_DIFF states the bar and
MAX_FRACTION bounds how many pixels may differ at all.
"""
import os

import numpy as np
import pytest

from figdraw_b200 import scenes
from figdraw_b200 import scenes_fuzz
from figdraw_b200 import scenes_synth as ss
from figdraw_b200 import abi
from figdraw_b200.abi import Op
from figdraw_b200.cuda_context import CudaContext, FigDrawError, render_trace
from figdraw_b200.figbackend import Trace
from oracle import oracle

pytestmark = pytest.mark.gpu

MAX_DIFF = 2  # LSB, north_star tolerance
MAX_FRACTION = 0.01  # pixels allowed to differ from the oracle at all


def diff_stats(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=2)
    return int(d.max()), float((d > 0).mean())


def check(trace, max_diff=MAX_DIFF, max_fraction=MAX_FRACTION):
    got = render_trace(trace)
    want = oracle.render_trace(trace)
    mx, frac = diff_stats(got, want)
    assert mx <= max_diff, f"max diff {mx} LSB"
    assert frac <= max_fraction, f"{frac:.4%} of pixels differ"
    return got, want


@pytest.mark.parametrize("name", sorted(scenes.GOLDEN_SCENES))
def test_golden_scenes(name):
    tr = scenes.golden_trace(name)
    got, _ = check(tr)
    mx, _ = diff_stats(got, scenes.load_golden(name))
    assert mx <= MAX_DIFF  # against the reference's own PNG


def test_rect_mask_scenes():
    check(scenes.trace_scene(scenes.layers_rect_mask, 800, 375))
    check(scenes.trace_scene(scenes.mixed_rect_mask_batch, 480, 180))


def test_cfg2_renderlist_100():
    check(ss.config_trace(2))


def test_cfg2_without_blur_is_one_segment():
    tr = ss.config_trace(2, with_blur=False)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    render_trace(tr, ctx)
    assert ctx.frameStats().n_segments == 1
    ctx.close()


def test_cfg3_text_and_msdf():
    check(ss.config_trace(3, 1920, 1080, n_glyphs=6000, msdf_glyphs=500))


@pytest.mark.parametrize("rect_mask", [False, True])
def test_cfg4_clip_table(rect_mask):
    check(ss.config_trace(4, 1920, 1080, rows=60, cols=8, rect_mask=rect_mask))


def test_cfg5_small():
    check(ss.config_trace(5, 1280, 720, n_rects=12000, n_glyphs=2400))


def test_cfg5_scaled_x2():
    check(ss.config_trace(5, 1920, 1080, n_rects=6000, n_glyphs=1200, scale=2.0))


# ------------------------------------------------------------------ BASELINE.json configs at their stated sizes
def _check_bins(tr, ctx):
    ref = oracle.reference_bins(tr)
    assert ctx.frameStats().n_segments == len(ref)
    for seg, (off_ref, ent_ref) in enumerate(ref):
        off, ent = ctx.debugBins(seg)
        assert np.array_equal(off, off_ref), f"segment {seg}: tile offsets differ"
        assert np.array_equal(ent, ent_ref), f"segment {seg}: draw order differs"


FULL_SIZE = {
    "cfg3 text page + star 3840x2160": lambda: ss.config_trace(3),
    "cfg3 + 20k MSDF glyph quads 3840x2160": lambda: ss.config_trace(3, msdf_glyphs=20000),
    "cfg4 180x12 sub-clip 3840x2160": lambda: ss.config_trace(4),
    "cfg4 180x12 rect-mask 3840x2160": lambda: ss.config_trace(4, rect_mask=True),
    "cfg5 100k rects + 20k glyphs 3840x2160": lambda: ss.config_trace(5),
    "cfg5 x2 7680x4320": lambda: ss.config_trace(5, 7680, 4320, scale=2.0),
}


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_baseline_configs_at_full_size(name):
    """every BASELINE config at the size BASELINE.json states -- pixels within 2 LSB of the oracle
    AND bin lists / draw order bit-exact."""
    tr = FULL_SIZE[name]()
    assert (tr.width, tr.height) in ((3840, 2160), (7680, 4320))
    ctx = CudaContext(atlasSize=tr.atlas_size)
    got = render_trace(tr, ctx)
    want = oracle.render_trace(tr)
    mx, frac = diff_stats(got, want)
    assert mx <= MAX_DIFF, f"max diff {mx} LSB"
    assert frac <= MAX_FRACTION, f"{frac:.4%} of pixels differ"
    _check_bins(tr, ctx)
    ctx.close()


def test_more_chunks_than_one_segment_table_block():
    """541 893 primitives = 1 059 chunks of 512: the fine binner addresses a bin's segments 1 024 chunks at a time, so this
    scene takes the second block of the table (no BASELINE config does).  Pixels within 2 LSB, bin lists bit-exact."""
    tr = ss.config_trace(5, 1920, 1080, n_rects=240000, n_glyphs=2000, scale=0.35)
    assert tr.n_draws > 1024 * 512
    ctx = CudaContext(atlasSize=tr.atlas_size)
    got = render_trace(tr, ctx)
    want = oracle.render_trace(tr)
    mx, frac = diff_stats(got, want)
    assert mx <= MAX_DIFF, f"max diff {mx} LSB"
    assert frac <= MAX_FRACTION, f"{frac:.4%} of pixels differ"
    _check_bins(tr, ctx)
    ctx.close()


def test_cfg5_8k_bands_reassemble_the_frame():
    """configs[4]: the 8K frame partitioned into 8 tile-row bands equals the single-context frame, bit for bit."""
    tr = ss.config_trace(5, 7680, 4320, scale=2.0)
    full = render_trace(tr)
    out = np.zeros_like(full)
    for r in range(8):
        ctx = CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=8)
        img = render_trace(tr, ctx)
        y0, y1 = ctx.bandRows()
        out[y0:y1] = img[y0:y1]
        ctx.close()
    assert np.array_equal(out, full)


@pytest.mark.parametrize("builder", [
    lambda: scenes.golden_trace("layers_clip"),
    lambda: ss.config_trace(2),
    lambda: ss.config_trace(4, 1280, 720, rows=24, cols=6),
    lambda: ss.config_trace(5, 1280, 720, n_rects=6000, n_glyphs=1000),
])
def test_bin_lists_bit_exact(builder):
    tr = builder()
    ctx = CudaContext(atlasSize=tr.atlas_size)
    render_trace(tr, ctx)
    ref = oracle.reference_bins(tr)
    st = ctx.frameStats()
    assert st.n_segments == len(ref)
    assert (st.tile_w, st.tile_h) == (16, 16)
    for seg, (off_ref, ent_ref) in enumerate(ref):
        off, ent = ctx.debugBins(seg)
        assert np.array_equal(off, off_ref), f"segment {seg}: tile offsets differ"
        assert np.array_equal(ent, ent_ref), f"segment {seg}: draw order differs"
    ctx.close()


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzzed_call_streams(seed):
    """Random streams over every op/mode: rotation, mirroring, nested clips, rect masks, elliptical corners, Beziers,
    filled quads, minified/flipped images, MSDF strokes, AA changes, backdrop blurs."""
    tr = scenes_fuzz.random_trace(seed)
    got, want = check(tr, max_fraction=0.03)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    render_trace(tr, ctx)
    ref = oracle.reference_bins(tr)
    assert ctx.frameStats().n_segments == len(ref)
    for seg, (off_ref, ent_ref) in enumerate(ref):
        off, ent = ctx.debugBins(seg)
        assert np.array_equal(off, off_ref) and np.array_equal(ent, ent_ref), f"segment {seg}: bins differ"
    ctx.close()


def test_replay_and_rerender_are_idempotent():
    tr = ss.config_trace(5, 1280, 720, n_rects=4000, n_glyphs=800)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    a = render_trace(tr, ctx).copy()
    ctx.replayFrame()
    b = ctx.readPixels().copy()
    c = render_trace(tr, ctx)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    ctx.close()


def test_graph_replay_equals_plain_replay():
    """fdc_replay_frame as ONE CUDA-graph launch (default) == the launches issued one by one, for frames with blurs,
    masks and several segments; a new recording, regrown lists or a rebound framebuffer drop the graph."""
    for tr in (ss.config_trace(2, 1280, 720), ss.config_trace(4, 1280, 720, rows=30, cols=6), scenes.golden_trace("layers_clip"),
               ss.config_trace(5, 1280, 720, n_rects=3000, n_glyphs=500)):
        ctx = CudaContext(atlasSize=tr.atlas_size)
        want = render_trace(tr, ctx).copy()
        ctx.setReplayGraph(False)
        ctx.replayFrame()
        assert np.array_equal(ctx.readPixels(), want)
        ctx.setReplayGraph(True)
        for _ in range(3):  # first one captures, the others launch the graph
            ctx.replayFrame()
            assert np.array_equal(ctx.readPixels(), want)
        assert ctx.frameStats().gpu_ms > 0.0
        # a different recording on the same context drops the graph; then the first one again
        k = len(tr.calls) // 2
        while tr.calls[k]["op"] < 32 and k + 1 < len(tr.calls):
            k += 1
        if not (tr.calls["op"][:k] == Op.BEGIN_MASK).any() and not (tr.calls["op"][:k] == Op.SAVE_TRANSFORM).any():
            ctx.beginFrame((tr.width, tr.height), clearMain=True)
            ctx.submitCalls(tr.calls[:k])
            ctx.endFrame()
            half = ctx.readPixels().copy()
            ctx.replayFrame()
            ctx.replayFrame()
            assert np.array_equal(ctx.readPixels(), half)
        assert np.array_equal(render_trace(tr, ctx), want)
        ctx.debugLimitLists(0, 64)  # the replay overflows, regrows and re-captures
        ctx.replayFrame()
        assert np.array_equal(ctx.readPixels(), want)
        ctx.replayFrame()
        assert np.array_equal(ctx.readPixels(), want)
        ctx.close()


def test_split_frame_equals_single_frame():
    """Per-draw UNORM8 quantisation makes the frame splittable at any draw: second half with clearMain=false."""
    tr = ss.config_trace(5, 1280, 720, n_rects=4000, n_glyphs=0)
    full = render_trace(tr)
    calls = tr.calls
    k = len(calls) // 2
    while calls[k]["op"] < 32:
        k += 1
    ctx = CudaContext(atlasSize=tr.atlas_size)
    for _i, key, img in tr.images:
        ctx.putImage(key, img)
    ctx.beginFrame((tr.width, tr.height), clearMain=True)
    ctx.submitCalls(calls[:k])
    ctx.restoreTransform()
    ctx.endFrame()
    ctx.beginFrame((tr.width, tr.height), clearMain=False)
    ctx.saveTransform()
    ctx.submitCalls(calls[k:])
    ctx.endFrame()
    got = ctx.readPixels()
    mx, frac = diff_stats(got, full)
    assert np.array_equal(got, full), f"split frame differs: max {mx} LSB, {frac:.5%} of pixels"
    ctx.close()


def test_tile_bands_reassemble_the_frame():
    tr = ss.config_trace(5, 1280, 720, n_rects=4000, n_glyphs=800)
    full = render_trace(tr)
    out = np.zeros_like(full)
    for n in (2, 3):
        for r in range(n):
            ctx = CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=n)
            img = render_trace(tr, ctx)
            y0, y1 = ctx.bandRows()
            out[y0:y1] = img[y0:y1]
            ctx.close()
        assert np.array_equal(out, full)


def test_host_chosen_bands_reassemble_the_frame_and_row_costs_add_up():
    """fdc_set_band_tile_rows: unequal bands (one of them a single tile row) give the same frame; a tile's list does not
    depend on the partition, so the ranks' per-row entry counts (fdc_get_tile_row_costs) sum to the single-context
    profile -- which is what bands.balance_rows splits."""
    from figdraw_b200.bands import balance_rows

    tr = ss.config_trace(5, 1280, 720, n_rects=4000, n_glyphs=800)
    one = CudaContext(atlasSize=tr.atlas_size)
    full = render_trace(tr, one)
    profile = one.tileRowCosts().astype(np.int64)
    one.close()
    tiles_y = (tr.height + 15) // 16
    assert profile.shape == (tiles_y,) and profile.sum() > 0
    balanced = balance_rows(profile + 3 * ((tr.width + 15) // 16), 4)
    for bounds in ([0, 3, 20, 21, tiles_y], balanced):
        out = np.zeros_like(full)
        total = np.zeros(tiles_y, dtype=np.int64)
        for r in range(4):
            ctx = CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=4)
            ctx.setBandTileRows(bounds)
            img = render_trace(tr, ctx)
            y0, y1 = ctx.bandRows()
            assert (y0, y1) == (bounds[r] * 16, min(bounds[r + 1] * 16, tr.height))
            out[y0:y1] = img[y0:y1]
            costs = ctx.tileRowCosts().astype(np.int64)
            assert not costs[: bounds[r]].any() and not costs[bounds[r + 1]:].any()
            total += costs
            ctx.close()
        assert np.array_equal(out, full)
        assert np.array_equal(total, profile)
    # the balanced split is no worse than the equal one on the cost it was given
    cost = profile + 3 * ((tr.width + 15) // 16)
    per = (tiles_y + 3) // 4
    worst = lambda b: max(int(cost[b[i]:b[i + 1]].sum()) for i in range(4))  # noqa: E731
    assert worst(balanced) <= worst([min(i * per, tiles_y) for i in range(5)])
    # bad boundaries are refused
    ctx = CudaContext(atlasSize=tr.atlas_size, rank=0, nRanks=4)
    with pytest.raises(FigDrawError):
        ctx.setBandTileRows([0, 5, 4, 20, tiles_y])
    with pytest.raises(FigDrawError):
        ctx.setBandTileRows([0, tiles_y])
    ctx.close()


def _band_contexts(tr, n, gather="stores"):
    from figdraw_b200.bands import padded_rows

    ctxs = [CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=n) for r in range(n)]
    for c in ctxs:
        c.reserveFramebuffer(tr.width, padded_rows(tr.height, n))
    ptrs = [c.framebufferPtr() for c in ctxs]
    for c in ctxs:
        c.setPeerFramebuffers(ptrs)
        c.setPeerGather(gather, 3)
        for _idx, key, img in tr.images:
            c.putImage(key, img)
    return ctxs


def _submit(c, tr, calls):
    c.beginFrame((tr.width, tr.height), clearMain=tr.clear is not None, clearMainColor=tr.clear or (1.0, 1.0, 1.0, 1.0))
    c.submitCalls(calls)
    c.endFrame()


def _render_banded_with_peers(tr, n, gather="stores", bounds=None):
    """n band contexts on one device, framebuffers cross-registered as peers: every rank submits the frame, the shade
    kernel's last segment stores each band into every framebuffer (or the copy engines ship it slice by slice), blur
    halo rows are read from the owner."""
    ctxs = _band_contexts(tr, n, gather)
    try:
        if bounds is not None:
            for c in ctxs:
                c.setBandTileRows(bounds)
        # Size every context's buffers one rank at a time with the blur calls removed (no cross-rank waits): all ranks
        # share this process and device, and an allocation while a peer spins on our flags would stall both.
        for c in ctxs:
            _submit(c, tr, tr.calls[tr.calls["op"] != Op.BACKDROP_BLUR])
            c.sync()
        for _ in range(2):  # twice: the cross-rank flags carry over from frame to frame
            for c in ctxs:
                _submit(c, tr, tr.calls)
            for c in ctxs:
                c.sync()
        return [c.readPixels() for c in ctxs]
    finally:
        for c in ctxs:
            c.close()


def test_blur_halo_barrier_times_out_instead_of_hanging():
    tr = ss.config_trace(2, 640, 360)
    ctxs = _band_contexts(tr, 2)
    try:
        _submit(ctxs[0], tr, tr.calls)  # rank 1 never submits
        with pytest.raises(FigDrawError, match="timed out"):
            ctxs[0].sync()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("n", [2, 3])
def test_backdrop_blur_halo_exchange_across_bands(n):
    """north_star: 'halo exchange for blur'.  A blur panel straddling band boundaries reads rows the neighbours shaded."""
    traces = [ss.config_trace(2, 1280, 720), ss.config_trace(4, 1280, 720, rows=40, cols=8)]
    from figdraw_b200.scenes_fuzz import random_trace
    traces += [t for t in (random_trace(s) for s in range(40)) if (t.calls["op"] == Op.BACKDROP_BLUR).any()][:4]
    assert len(traces) >= 4
    for tr in traces:
        # same frame sequence on one context: state such as the SDF AA factor carries over from frame to frame
        full = _render_banded_with_peers(tr, 1)[0]
        for r, img in enumerate(_render_banded_with_peers(tr, n)):
            mx, frac = diff_stats(img, full)
            assert np.array_equal(img, full), f"rank {r}/{n}: max {mx} LSB, {frac:.5%} of pixels"


def test_backdrop_blur_halo_exchange_across_unequal_bands():
    """Host-chosen bands: the blur's halo rows come from whichever rank owns them under the boundaries in force."""
    tr = ss.config_trace(2, 1280, 720)
    tiles_y = (tr.height + 15) // 16
    full = _render_banded_with_peers(tr, 1)[0]
    for bounds in ([0, 10, 12, tiles_y], [0, 30, 31, tiles_y]):
        for r, img in enumerate(_render_banded_with_peers(tr, 3, bounds=bounds)):
            assert np.array_equal(img, full), f"rank {r}, bounds {bounds}"


@pytest.mark.parametrize("n", [2])
def test_copy_engine_gather_assembles_the_frame(n):
    """FDC_GATHER_COPY: the last segment is shaded in slices, finished slices are copied to the peers.
    (Two ranks only: all ranks of this test share one process, and with 1 + 4 streams per context a third context
    exceeds the device's 8 hardware queues -- a context's kernels could then queue behind another context's barrier.)"""
    for tr in (ss.config_trace(5, 1280, 720, n_rects=3000, n_glyphs=600), ss.config_trace(2, 1280, 720)):
        full = _render_banded_with_peers(tr, 1)[0]
        for r, img in enumerate(_render_banded_with_peers(tr, n, gather="copy")):
            assert np.array_equal(img, full), f"rank {r}/{n}"


@pytest.mark.parametrize("n", [2, 3])
def test_shared_framebuffer_fuses_the_gather(n):
    """fdc_bind_shared_framebuffer: host-allocated framebuffers every rank can reach (here: torch tensors on one device,
    no multicast mapping).  The copy-out stores every finished chunk into every copy and each frame ends with a flag
    barrier; blur halos are read from the peers' copies.  Every copy must hold the single-context frame."""
    import torch
    from figdraw_b200.bands import padded_rows

    for tr in (ss.config_trace(5, 1280, 720, n_rects=3000, n_glyphs=600), ss.config_trace(2, 1280, 720),
               ss.config_trace(5, 333, 217, n_rects=300, n_glyphs=60)):  # 333: rows are not 16-byte multiples
        full = render_trace(tr)
        rows = padded_rows(tr.height, n)
        nbytes = ((tr.width * rows * 4 + 255) & ~255) + 4096
        bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(n)]
        ctxs = [CudaContext(atlasSize=tr.atlas_size, rank=r, nRanks=n) for r in range(n)]
        try:
            for r, c in enumerate(ctxs):
                c.bindSharedFramebuffer(bufs[r].data_ptr(), nbytes, [b.data_ptr() for b in bufs], 0, tr.width, rows)
                for _idx, key, img in tr.images:
                    c.putImage(key, img)
            # Size every context's buffers one rank at a time without cross-rank waits: all ranks share this process and
            # device, and an allocation (a device-wide sync) while a peer spins on our flags would stall both.
            for c in ctxs:
                c.setFrameBarrier(False)
                _submit(c, tr, tr.calls[tr.calls["op"] != Op.BACKDROP_BLUR])
                c.sync()
                c.setFrameBarrier(True)
            for _ in range(2):
                for c in ctxs:
                    _submit(c, tr, tr.calls)
                for c in ctxs:
                    c.sync()
            torch.cuda.synchronize()
            for r in range(n):
                got = bufs[r][: tr.height * tr.width * 4].view(tr.height, tr.width, 4).cpu().numpy()
                assert np.array_equal(got, full), f"copy {r}/{n} of a {tr.width}x{tr.height} frame"
        finally:
            for c in ctxs:
                c.close()


def test_backdrop_blur_under_bands_needs_peers():
    tr = ss.config_trace(2, 640, 360)
    ctx = CudaContext(atlasSize=tr.atlas_size, rank=0, nRanks=2)
    with pytest.raises(FigDrawError, match="fdc_set_peer_framebuffers"):
        render_trace(tr, ctx)
    ctx.close()


def test_empty_and_degenerate_inputs():
    ctx = CudaContext()
    ctx.beginFrame((64, 48), clearMain=True, clearMainColor=(0.2, 0.4, 0.6, 1.0))
    ctx.endFrame()
    img = ctx.readPixels()
    assert img.shape == (48, 64, 4) and tuple(img[0, 0]) == (51, 102, 153, 255)
    from figdraw_b200.figbackend import ZeroRadii, solid

    ctx.beginFrame((64, 48), clearMain=True)
    ctx.drawRoundedRectSdf((10, 10, 0, 5), solid(0xFF000000), ZeroRadii)      # zero width: dropped
    ctx.drawRoundedRectSdf((-500, -500, 100, 100), solid(0xFF000000), ZeroRadii)  # off-screen
    ctx.drawImage(12345, (0, 0), [0xFFFFFFFF] * 4)                              # missing image: warn + skip
    ctx.endFrame()
    assert ctx.missing_images == 1
    assert (ctx.readPixels() == 255).all()
    ctx.close()


def test_state_errors_match_reference_asserts():
    from figdraw_b200.cuda_context import FigDrawError
    from figdraw_b200.figbackend import ZeroRadii

    ctx = CudaContext()
    with pytest.raises(FigDrawError):
        ctx.endFrame()  # "ctx.beginFrame was not called first."
    ctx.beginFrame((32, 32), clearMain=True)
    ctx.beginMask((0, 0, 8, 8), ZeroRadii)
    with pytest.raises(FigDrawError):
        ctx.beginMask((0, 0, 8, 8), ZeroRadii)  # "ctx.beginMask has already been called."
    ctx.endMask()
    with pytest.raises(FigDrawError):
        ctx.endFrame()  # "Not all masks have been popped."
    ctx.close()


def test_atlas_packer_matches_reference_placement():
    """putImage places images exactly where the reference's skyline packer would (glcontext.nim:541-586)."""
    ctx = CudaContext(atlasSize=256)
    o = oracle.Oracle(256)
    rng = np.random.default_rng(7)
    for k in range(40):
        w, h = int(rng.integers(2, 60)), int(rng.integers(2, 60))
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        r1, g1 = ctx.putImage(1000 + k, img)
        r2, g2 = o.put_image(1000 + k, img)
        assert r1 == pytest.approx(r2, abs=0) and g1 == g2
    assert ctx.atlasSize() == o.atlas_size and ctx.atlasSize() > 256  # it grew
    ctx.close()


def test_submit_draws_equals_submit_calls():
    """fdc_submit_draws (no host inspection, one async copy per run) renders exactly what fdc_submit_calls does."""
    from figdraw_b200.cuda_context import prepare_calls

    tr = ss.config_trace(5, 1280, 720, n_rects=6000, n_glyphs=1200)
    want = render_trace(tr)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    for _i, key, img in tr.images:
        ctx.putImage(key, img)
    prepared = prepare_calls(tr.calls)
    assert any(is_draw and b - a >= 2048 for is_draw, a, b in prepared[1])
    ctx.beginFrame((tr.width, tr.height), clearMain=True)
    ctx.submitPrepared(prepared)
    ctx.endFrame()
    assert np.array_equal(ctx.readPixels(), want)
    off, ent = ctx.debugBins(0)
    ref_off, ref_ent = oracle.reference_bins(tr)[0]
    assert np.array_equal(off, ref_off) and np.array_equal(ent, ref_ent)
    ctx.close()


# ------------------------------------------------------------------ native front-end (fdc_render_frame)
def test_native_front_end_renders_the_same_pixels():
    """One fdc_render_frame call over POD scene records == the per-call front-end driving the same backend."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from test_flatten import random_renders

    from figdraw_b200 import figrender, native_scene

    cases = [(scenes.rgb_boxes_sdf(800.0, 600.0), 800, 600, []), (scenes.layers_clip(800.0, 600.0), 800, 600, []),
             (scenes.image_scene(800.0, 600.0), 800, 600, [(scenes.IMG1_KEY, scenes.load_img1())]),
             (ss.renderlist_100(1280.0, 720.0), 1280, 720, []),
             (ss.text_page(1280.0, 720.0, n_glyphs=2000, msdf_glyphs=300), 1280, 720, ss.text_page_images()),
             (ss.clip_mask_table(1280.0, 720.0, rows=30, cols=6), 1280, 720, [])]
    cases += [(random_renders(s), 640, 480, [(1000 + k, np.full((6, 5, 4), 200, np.uint8)) for k in range(0, 40, 3)]) for s in range(6)]
    for renders, w, h, images in cases:
        a, b = CudaContext(atlasSize=2048), CudaContext(atlasSize=2048)
        for key, img in images:
            a.putImage(key, img)
            b.putImage(key, img)
        figrender.setFigUiScale(1.0)
        figrender.renderFrame(a, renders, (float(w), float(h)))
        b.renderFrameNative(native_scene.pack_renders(renders), (w, h))
        pa, pb = a.readPixels(), b.readPixels()
        assert np.array_equal(pa, pb)
        a.close()
        b.close()


def test_native_front_end_cfg5_scene_equals_call_stream():
    tr = ss.config_trace(5, 1280, 720, n_rects=6000, n_glyphs=1500)
    scene = ss.rects_and_glyphs_scene(1280, 720, n_rects=6000, n_glyphs=1500)
    want = render_trace(tr)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    for _i, key, img in tr.images:
        ctx.putImage(key, img)
    ctx.renderFrameNative(scene, (1280, 720))
    assert np.array_equal(ctx.readPixels(), want)
    ctx.close()


def test_async_readback_matches_readpixels():
    import torch

    tr = ss.config_trace(2, 1280, 720)
    ctx = CudaContext(atlasSize=tr.atlas_size)
    want = render_trace(tr, ctx)
    host = torch.empty((tr.height, tr.width, 4), dtype=torch.uint8).pin_memory().numpy()
    for _ in range(2):
        host[:] = 0
        ctx.beginFrame((tr.width, tr.height), clearMain=tr.clear is not None, clearMainColor=tr.clear or (1.0, 1.0, 1.0, 1.0))
        ctx.submitCalls(tr.calls)
        ctx.endFrame()
        ctx.readPixelsAsync(host)
        ctx.sync()
        assert np.array_equal(host, want)
    # a much larger scene on the same context: the bin lists overflow, the frame is re-run at sync and read back again
    big = ss.config_trace(5, 1280, 720, n_rects=6000, n_glyphs=0)
    want_big = render_trace(big)
    ctx.beginFrame((big.width, big.height), clearMain=True)
    ctx.submitCalls(big.calls)
    ctx.endFrame()
    ctx.readPixelsAsync(host)
    ctx.sync()
    assert np.array_equal(host, want_big)
    ctx.close()


def test_ragged_and_extreme_sizes():
    """Frames that are not a multiple of the 16-px tile / 128-px bin, a 1x1 frame, a very wide one, huge coordinates."""
    from figdraw_b200.figbackend import TraceBackend, circularRadii, solid
    from figdraw_b200.fignodes import rgba

    def scene(w, h):
        tb = TraceBackend()
        tb.beginFrame((w, h), clearMain=True, clearMainColor=(0.9, 0.9, 0.8, 1.0))
        tb.drawRoundedRectSdf((-3.5, -2.25, w * 0.7, h * 0.6), solid(rgba(200, 30, 40, 200)), circularRadii((9, 3, 0, 14)))
        tb.drawRoundedRectSdf((w * 0.3, h * 0.2, w, h), solid(rgba(10, 90, 220, 255)), circularRadii((5, 5, 5, 5)),
                              mode=abi.SdfMode.sdfModeDropShadow, factor=6.0, spread=3.0, shapeSize=(w * 0.5, h * 0.5))
        tb.drawRoundedRectSdf((-1.0e6, -1.0e6, 2.0e6, 2.0e6), solid(rgba(0, 0, 0, 20)), circularRadii((0, 0, 0, 0)))
        tb.saveTransform()
        tb.translate((w * 0.5, h * 0.5))
        tb.rotate(0.4)
        tb.drawRoundedRectSdf((-w * 0.2, -h * 0.1, w * 0.4, h * 0.2), solid(rgba(20, 160, 60, 180)), circularRadii((4, 4, 4, 4)))
        tb.restoreTransform()
        tb.endFrame()
        return tb.trace()

    for w, h in ((1, 1), (17, 5), (333, 217), (4099, 33), (130, 2051)):
        tr = scene(w, h)
        got, want = render_trace(tr), oracle.render_trace(tr)
        mx, frac = diff_stats(got, want)
        assert mx <= MAX_DIFF, f"{w}x{h}: max {mx} LSB"
    # ragged bands: 217 rows = 14 tile rows over 3 ranks (5 + 5 + 4), the last tile row only 9 px high
    tr = scene(333, 217)
    full = render_trace(tr)
    out = np.zeros_like(full)
    for r in range(3):
        ctx = CudaContext(rank=r, nRanks=3)
        img = render_trace(tr, ctx)
        y0, y1 = ctx.bandRows()
        out[y0:y1] = img[y0:y1]
        ctx.close()
    assert np.array_equal(out, full)


def test_reference_spot_pixels():
    """The spot pixels asserted by the reference's own render tests, on the CUDA backend."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from test_oracle_golden import check_reference_spot_pixels

    check_reference_spot_pixels(render_trace)


def test_compact_rect_records_render_identically():
    """fdc_submit_rects64 (64-byte records expanded by the setup kernel) == the same frame from 128-byte records."""
    from figdraw_b200.cuda_context import prepare_calls

    for tr in (ss.config_trace(5, 1280, 720, n_rects=4000, n_glyphs=800), ss.config_trace(2, 1280, 720),
               ss.config_trace(4, 1280, 720, rows=30, cols=6)):
        want = render_trace(tr)
        ctx = CudaContext(atlasSize=tr.atlas_size)
        for _i, key, img in tr.images:
            ctx.putImage(key, img)
        prepared = prepare_calls(tr.calls, compact=True, min_compact_run=4)
        assert any(r[0] == "rects64" for r in prepared[1])
        for _ in range(2):
            ctx.beginFrame((tr.width, tr.height), clearMain=tr.clear is not None, clearMainColor=tr.clear or (1.0, 1.0, 1.0, 1.0))
            ctx.submitPrepared(prepared)
            ctx.endFrame()
            assert np.array_equal(ctx.readPixels(), want)
        ctx.replayFrame()
        assert np.array_equal(ctx.readPixels(), want)
        ctx.close()
