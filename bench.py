#!/usr/bin/env python
"""bench.py -- headline benchmark of the figdraw B200 render path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg5_4k|cfg5_8k|...]

A "step" is ONE FRAME of the hot path (setup + binning + shade [+ blur]) over one synthetic scene.
  N = 1   : BASELINE.json's target scene -- 100k shadowed rounded rects + 20k glyph quads at 3840x2160 (cfg5).
  N > 1   : the same scene at 7680x4320 (all sizes x2, configs[4]), framebuffer partitioned into tile-row bands,
            one rank per GPU, NCCL all-gather of the bands at the end of every frame (strong scaling of one frame).
            The line then carries `single_gpu_same_workload` -- the 8K frame timed on rank 0's GPU alone in the same
            run -- because the N = 1 line is the 4K scene: scaling of THIS workload is value / (N x that value).
`value` is Mpixels/s with the frame's inputs already resident in HBM (kernels only, CUDA events on the context's
stream, max over ranks); `e2e` is the same metric through the C ABI with HOST buffers: fdc_begin_frame +
fdc_submit_calls(host records) + fdc_end_frame + fdc_read_pixels(host), copies inside the timed region.
`--impl reference` times the CPU restatement of the reference's GL path (oracle/) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixels/s"

# Algorithmic flops per fragment by SdfMode (FMA = 2, sqrt/exp/div/cvt = 1).
FLOPS = {0: 89, 3: 62, 7: 63, 8: 68, 9: 88, 11: 66, 12: 66, 13: 99, 14: 99, 15: 99, 16: 99, 17: 78, 18: 150, 19: 150, 20: 150}
FLOPS_GRADIENT_EXTRA = 16
FLOPS_BLUR_PER_PIXEL_PASS = 140
N_MODES = 24


def workload_trace(name: str):
    from figdraw_b200 import scenes_synth as ss

    if name == "cfg5_4k":
        return ss.config_trace(5, 3840, 2160), "cfg5: 100k shadowed rounded rects + 20k glyph quads, 3840x2160, seed 5"
    if name == "cfg5_8k":
        return ss.config_trace(5, 7680, 4320, scale=2.0), "cfg5 x2: 100k shadowed rounded rects + 20k glyph quads, 7680x4320, seed 5"
    if name == "cfg2":
        return ss.config_trace(2), "cfg2: renderlist_100 shape (300 boxes, shadows, elliptical corners, 1 backdrop blur), 1920x1080"
    if name == "cfg3":
        return ss.config_trace(3), "cfg3: 20k atlas glyph quads + MSDF/MTSDF star, 3840x2160"
    if name == "cfg4":
        return ss.config_trace(4), "cfg4: clip-mask table 180x12 (sub-clip), 3-stop gradients, 2 backdrop blurs, 3840x2160"
    if name == "cfg4_rectmask":
        return ss.config_trace(4, rect_mask=True), "cfg4: clip-mask table 180x12 (rect-mask), 2 backdrop blurs, 3840x2160"
    raise SystemExit(f"unknown workload {name}")


def algorithmic_flops(counts: np.ndarray) -> float:
    total = 0.0
    for m, f in FLOPS.items():
        total += float(counts[m]) * f + float(counts[N_MODES + m]) * (f + FLOPS_GRADIENT_EXTRA)
    return total


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled in-process every ~2 ms (the timed region
    of the default run is only ~25 ms long), `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.nvml = None
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.sm_max = None

    def _nvml_loop(self):
        nv, h = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nvml = (nv, h)
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            nv, _h = self.nvml
            self.stop_flag = True
            self.thread.join(timeout=1)
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            reasons = sorted({n for _sm, bits in self.samples for n, b in names if bits & b})
            sm = [float(x) for x, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def run_reference(args, rank: int, out=sys.stdout):
    """CPU arm: the oracle port of the reference's GL path, all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as orc

    name = args.workload or ("cfg5_4k" if args.gpus == 1 else "cfg5_8k")
    trace, desc = workload_trace(name)
    cores = orc.max_threads()
    o = orc.Oracle(trace.atlas_size)
    for _i, key, img in trace.images:
        o.put_image(key, img)
    has_blur = bool((trace.calls["op"] == 13).any())
    H = trace.height
    rows = (0, H) if has_blur else (H // 2 - H // 32, H // 2 + H // 32)  # 1/16 of the frame, centred
    for _ in range(args.warmup):
        o.render(trace.width, H, trace.calls, clear=trace.clear, n_threads=cores, rows=rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.render(trace.width, H, trace.calls, clear=trace.clear, n_threads=cores, rows=rows)
    dt = (time.perf_counter() - t0) / args.steps
    px = trace.width * (rows[1] - rows[0])
    val = px / dt / 1e6
    sample = f"rows {rows[0]}..{rows[1]} of {H} ({px} px) of the same frame per step"
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": METRIC, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3 * (trace.width * H) / px, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "note": "restated-reference CPU rasteriser (oracle port of the GL path), not llvmpipe; "
                       "ms_per_step extrapolated from the sample to the whole frame"},
            "cpu_baseline": {"value": round(val, 3), "unit": METRIC, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 3), "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "frames_per_s": round(val * 1e6 / (trace.width * H), 4)}
    print(json.dumps(line), file=out, flush=True)


def native_frontend_probe(name: str, calls_np):
    """Informational: host time of fdc_flatten_renders for the node-level form of the workload --
    100k nkRectangle + nkText records -> the call stream the timed steps replay (checked byte for byte)."""
    import ctypes

    from figdraw_b200 import abi, scenes_synth as ss

    lib = abi.load_library()
    big = name == "cfg5_8k"
    scene = ss.rects_and_glyphs_scene(7680 if big else 3840, 4320 if big else 2160, scale=2.0 if big else 1.0)
    keys = np.asarray(sorted(ss.glyph_image_keys()), dtype=np.uint64)
    env = abi.FdcFlattenEnv(1.0, 1.0, 1.2, 0, keys.ctypes.data, len(keys))
    out = np.zeros(len(calls_np) + 64, dtype=abi.CALL_DTYPE)
    n = ctypes.c_size_t(0)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        rc = lib.fdc_flatten_renders(ctypes.byref(scene.scene), ctypes.byref(env), out.ctypes.data, len(out), ctypes.byref(n))
        ts.append(time.perf_counter() - t0)
    same = rc == 0 and n.value == len(calls_np) and out[: n.value].tobytes() == calls_np.tobytes()
    return {"nodes": scene.n_nodes, "records": int(n.value), "flatten_ms": round(float(np.median(ts[1:])) * 1e3, 3),
            "threads": min(16, os.cpu_count() or 1), "records_equal_benchmark_stream": bool(same)}


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun warnings) print to fd 1; the driver wants ONE JSON line there.
    Point fd 1 at stderr for the run and keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="nccl", choices=["p2p", "nccl", "ce"],
                    help="N>1: NCCL all-gather of the bands (nccl), fused peer stores from the shade kernel (p2p), or copy "
                         "engines shipping finished band slices to the peers while the next slice is shaded (ce)")
    ap.add_argument("--sub-bands", type=int, default=4)
    ap.add_argument("--records", default=None, choices=["compact", "full"],
                    help="e2e upload: 64-byte fdc_rect64 records for rounded rects with circular corners, or 128-byte fdc_call only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return

    import torch
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from figdraw_b200.cuda_context import CudaContext, prepare_calls, prepared_upload_bytes

    name = args.workload or ("cfg5_4k" if world == 1 else "cfg5_8k")
    trace, desc = workload_trace(name)
    W, H = trace.width, trace.height
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    ctx = CudaContext(atlasSize=trace.atlas_size, device=local_rank, rank=rank, nRanks=world)
    for _i, key, img in trace.images:
        ctx.putImage(key, img)

    from figdraw_b200 import bands

    band_rows, _layout = bands.band_layout(H, world)
    # a backdrop blur under a band partition reads halo rows out of the neighbours' framebuffers: peer mappings needed
    has_blur = bool((trace.calls["op"] == 13).any())
    use_p2p = world > 1 and (args.gather in ("p2p", "ce") or has_blur)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    if use_p2p:
        # Fused all-gather: every rank's shade kernel stores its finished pixels into all peers' framebuffers over
        # NVLink (CUDA IPC mappings); a one-element NCCL all-reduce on the same stream is the completion barrier.
        ctx.reserveFramebuffer(W, band_rows * world)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.framebufferIpcHandle())
        peers = [0 if r == rank else ctx.openPeerFramebuffer(handles[r]) for r in range(world)]
        ctx.setPeerFramebuffers(peers)
        if args.gather == "ce":
            ctx.setPeerGather("copy", args.sub_bands)
        token = torch.zeros(1, dtype=torch.int32, device=dev)
        fb = None
    else:
        # Framebuffer owned by torch so NCCL can all-gather the bands in place; rows padded to equal bands.
        fb = torch.zeros((band_rows * world, W, 4), dtype=torch.uint8, device=dev)
        ctx.bindFramebuffer(fb.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    calls_host = torch.from_numpy(trace.calls.view(np.uint8).reshape(-1, 128).copy()).pin_memory()
    calls_np = calls_host.numpy().view(trace.calls.dtype).reshape(-1)
    out_host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    out_np = out_host.numpy()
    keep_pinned = []
    # upload format of the e2e arm: compact records on one GPU (validated there, profiles/r01_e2e_records.md); the
    # multi-GPU runs keep the plain 128-byte records unless asked otherwise
    records = args.records or ("compact" if world == 1 else "full")

    def prepare_pinned(calls_pinned_np):
        """Run boundaries computed once (a host that emits the calls knows them); rounded rects with circular corners go
        as 64-byte fdc_rect64 records (`--records full` keeps everything at 128 bytes); every buffer page-locked."""
        calls_, runs = prepare_calls(calls_pinned_np, compact=records == "compact")
        out_runs = []
        for run in runs:
            if run[0] == "rects64":
                t = torch.from_numpy(run[3].view(np.uint8).reshape(-1, 64).copy()).pin_memory()
                keep_pinned.append(t)
                out_runs.append((run[0], run[1], run[2], t.numpy().view(run[3].dtype).reshape(-1)))
            else:
                out_runs.append(run)
        return calls_, out_runs

    prepared = prepare_pinned(calls_np)

    def gather():
        if use_p2p:
            dist.all_reduce(token)  # all ranks' shade kernels (and their peer stores) are complete after this
        else:
            bands.allgather_bands(fb, rank, world)

    def frame_e2e():
        ctx.beginFrame((W, H), clearMain=trace.clear is not None, clearMainColor=trace.clear or (1, 1, 1, 1))
        ctx.submitPrepared(prepared)
        ctx.endFrame()
        with torch.cuda.stream(stream):
            gather()
        y0, y1 = ctx.bandRows() if world > 1 else (0, H)
        ctx.readPixels((0, y0, W, y1 - y0), out=out_np[y0:y1])

    # first frame: uploads the recording, allocates everything
    frame_e2e()
    st = ctx.frameStats()
    launches_per_frame = int(st.n_launches)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn_step, steps):
        """Each step individually bracketed by CUDA events on the context stream; L2 flushed between steps."""
        evs = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn_step()
                e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def step_resident():
        ctx.replayFrame()
        gather()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    times = timed_loop(step_resident, args.steps)
    barrier()
    stats = ctx.frameStats()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = float(np.sum(times)) / args.steps

    # e2e: host records in, host pixels out, wall clock bracketed by synchronisation (copies are inside)
    for _ in range(2):
        frame_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(5, min(args.steps, 20))
    for _ in range(e2e_steps):
        frame_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # e2e, pipelined (N = 1): three contexts on three streams, like a three-image swap chain.  Every step still copies
    # its own records host->device and its own frame device->host; in steady state frame k's readback, frame k+1's
    # kernels and frame k+2's upload are in flight together.  Throughput over the steps, not latency.  (Measured: deeper
    # rings and fdc_read_pixels_async change nothing -- 33 MB down + 31 MB up per frame is what PCIe sustains in ~0.96 ms.)
    e2e_pipe_ms = None
    if world == 1:
        ring = [(ctx, prepared, out_np)]
        extra_ctx = []
        for _ in range(max(1, int(os.environ.get("FDC_E2E_RING", "3")) - 1)):
            c2 = CudaContext(atlasSize=trace.atlas_size, device=local_rank)
            for _i, key, img in trace.images:
                c2.putImage(key, img)
            o2 = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            k2 = calls_host.clone().pin_memory()
            ring.append((c2, prepare_pinned(k2.numpy().view(trace.calls.dtype).reshape(-1)), o2.numpy()))
            extra_ctx.append((c2, o2, k2))
        depth = len(ring)

        def submit(c, calls):
            c.beginFrame((W, H), clearMain=trace.clear is not None, clearMainColor=trace.clear or (1, 1, 1, 1))
            c.submitPrepared(calls)
            c.endFrame()

        def pipelined(steps):
            for k in range(min(depth - 1, steps)):
                submit(ring[k % depth][0], ring[k % depth][1])
            for k in range(steps):
                if k + depth - 1 < steps:
                    nxt = ring[(k + depth - 1) % depth]
                    submit(nxt[0], nxt[1])
                cur = ring[k % depth]
                cur[0].readPixels((0, 0, W, H), out=cur[2])

        pipelined(6)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipelined(e2e_steps)
        torch.cuda.synchronize()
        e2e_pipe_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        same = all(bool(np.array_equal(out_np, r[2])) for r in ring[1:])
        if not same:
            raise SystemExit("pipelined contexts produced different frames")
        for c2, _o, _k in extra_ctx:
            c2.close()

    gathered_ok = None
    if world > 1 and rank == 0:
        # the gathered frame on rank 0 must equal a single-context render of the whole frame, bit for bit
        ctx.replayFrame()
        with torch.cuda.stream(stream):
            gather()
        torch.cuda.synchronize()
    if world > 1 and rank != 0:
        ctx.replayFrame()
        with torch.cuda.stream(stream):
            gather()
        torch.cuda.synchronize()
    if world > 1 and rank == 0:
        import ctypes

        whole = np.empty((H, W, 4), dtype=np.uint8)
        if use_p2p:
            ctx._ck(ctx._lib.fdc_read_pixels(ctx._h, 0, 0, W, H, whole.ctypes.data))
        else:
            whole[:] = fb[:H].cpu().numpy()
        ref_ctx = CudaContext(atlasSize=trace.atlas_size, device=local_rank)
        for _i, key, img in trace.images:
            ref_ctx.putImage(key, img)
        ref_ctx.beginFrame((W, H), clearMain=trace.clear is not None, clearMainColor=trace.clear or (1, 1, 1, 1))
        ref_ctx.submitCalls(calls_np)
        ref_ctx.endFrame()
        gathered_ok = bool(np.array_equal(ref_ctx.readPixels(), whole))
        # the same frame on ONE GPU, measured here so that a strong-scaling ratio for this workload can be formed
        # (the N = 1 bench line runs the 4K target frame, BASELINE's multi-GPU config is the 8K one)
        single_ms = []
        for _ in range(7):
            flush.fill_(0)
            torch.cuda.synchronize()
            ref_ctx.replayFrame()
            single_ms.append(float(ref_ctx.frameStats().gpu_ms))
        single_gpu_ms = float(np.median(single_ms[2:]))
        ref_ctx.close()
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, stats.shade_ms, stats.bin_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, shade_ms, bin_ms = (float(v) for v in t.tolist())
    else:
        shade_ms, bin_ms = float(stats.shade_ms), float(stats.bin_ms)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        mpx = W * H / 1e6
        value = mpx / (ms_step * 1e-3)
        # roofline of the dominant kernel (shade): algorithmic flops from the oracle's exact fragment counts
        cpu_base, roof = None, None
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        peak_fp32 = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s per GPU, non-tensor FP32 at the clock seen under load
        from oracle import oracle as orc

        counts = orc.count_fragments(trace)  # exact fragments per mode (checker; counts only, nothing is shaded)
        n_frag = int(counts.sum())
        flops = algorithmic_flops(counts)
        bytes_alg = W * H * 4 + trace.n_draws * 128 * 2 + int(stats.n_tile_entries) * 4 * world
        ach = flops / (shade_ms * 1e-3) / 1e12 / world  # per GPU: every rank shades 1/world of the frame in shade_ms
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_shade_traffic.json")
        if name == "cfg5_4k" and world == 1 and os.path.exists(tp):
            traffic = json.load(open(tp)).get("traffic_bytes_per_launch")  # dram read+write from the ncu --set full capture
        roof = {"bound": "fp32", "kernel": "shade_kernel", "achieved": round(ach, 3), "peak": round(peak_fp32, 2),
                "unit": "TFLOP/s", "frac": round(ach / peak_fp32, 4), "traffic": traffic,
                "algorithmic_flops_per_frame": flops, "fragments_per_frame": n_frag,
                "note": "algorithmic = every fragment the reference's GL path would shade; the kernel provably skips "
                        "occluded and trivially covered ones, so this can exceed what is executed (see profiles/)",
                "peak_source": f"148 SM x 128 lanes x 2 x {sm_mhz:.0f} MHz (SM clock sampled during the timed region), per GPU",
                "shade_ms": round(shade_ms, 4), "bin_ms": round(bin_ms, 4),
                "hbm": {"algorithmic_bytes": bytes_alg, "achieved_gbs": round(bytes_alg / (ms_step * 1e-3) / 1e9 / world, 1),
                        "peak_gbs": peaks.get("hbm_gbs"),
                        "frac": round(bytes_alg / (ms_step * 1e-3) / 1e9 / world / peaks.get("hbm_gbs", 6650.0), 4),
                        "peak_source": peak_src}}
        if world == 1 and not args.no_cpu_baseline:
            cores = orc.max_threads()
            o = orc.Oracle(trace.atlas_size)
            for _i, key, img in trace.images:
                o.put_image(key, img)
            t0 = time.perf_counter()
            ref_img = o.render(W, H, trace.calls, clear=trace.clear, n_threads=cores)
            cpu_s = time.perf_counter() - t0
            d = np.abs(out_np.astype(np.int16) - ref_img.astype(np.int16)).max(axis=2)
            cpu_base = {"value": round(mpx / cpu_s, 3), "unit": METRIC, "cores": cores, "kind": "port",
                        "sample": f"1 full frame of the same scene ({W}x{H}, {n_frag} fragments) in {cpu_s:.2f} s"}
            roof["parity_vs_oracle"] = {"max_abs_diff_lsb": int(d.max()), "pixels_differing": int((d > 0).sum())}
        h2d = int(prepared_upload_bytes(prepared))
        d2h = int(W * H * 4)
        line = {"metric": METRIC, "value": round(value, 2), "unit": METRIC, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "frame": [W, H], "primitives": trace.n_draws, "l2": "flushed between steps (256 MiB fill)",
                           "partition": "single GPU" if world == 1 else (
                               (f"{world} tile-row bands, finished band slices copied to the peers by the copy engines (NVLink) while "
                                "the next slice is shaded" if args.gather == "ce" else
                                f"{world} tile-row bands, band all-gather fused into the shade kernel (peer stores over NVLink)")
                               if use_p2p else f"{world} tile-row bands + NCCL all-gather")},
                "frames_per_s": round(1e3 / ms_step, 2),
                "e2e": {"value": round(mpx / ((e2e_pipe_ms or e2e_ms) * 1e-3), 2), "unit": METRIC,
                        "ms_per_step": round(e2e_pipe_ms or e2e_ms, 4), "latency_ms": round(e2e_ms, 4),
                        "mode": (f"{depth} contexts in flight: step k's readback, step k+1's kernels and step k+2's upload overlap"
                                 if e2e_pipe_ms else "one frame at a time"),
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches_per_frame * args.steps, "launches_per_frame": launches_per_frame,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base}
        if gathered_ok is not None:
            line["gathered_frame_equals_single_gpu"] = gathered_ok
            line["single_gpu_same_workload"] = {"ms_per_step": round(single_gpu_ms, 4), "value": round(mpx / (single_gpu_ms * 1e-3), 2),
                                                "unit": METRIC, "note": "this rank-0 GPU alone on the same frame (no gather); "
                                                "value / (n_gpus x this) is the strong-scaling efficiency of the workload"}
        if world == 1 and name in ("cfg5_4k", "cfg5_8k"):
            line["native_frontend"] = native_frontend_probe(name, calls_np)
        print(json.dumps(line), file=real_stdout, flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
