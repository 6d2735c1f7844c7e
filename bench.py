#!/usr/bin/env python
"""bench.py -- headline benchmark of the figdraw B200 render path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg5_4k|cfg5_8k|cfg2|...]

A "step" is ONE FRAME of the hot path (setup + binning + shade [+ blur] [+ band all-gather]) over one synthetic scene.
The workload is the SAME at every N: BASELINE.json's target scene -- 100k shadowed rounded rects + 20k glyph quads at
3840x2160 (cfg5) -- so value(N) / (N x value(1)) is a real strong-scaling curve.  For N > 1 the framebuffer is
partitioned into tile-row bands, one rank per GPU, and the bands are all-gathered at the end of every frame; the line
then also carries `cfg5_8k`: BASELINE's multi-GPU config (the same scene at 7680x4320, all sizes x2) timed the same way
next to the same frame on one GPU.
`value` is Mpixels/s with the frame's inputs already resident in HBM (kernels only, CUDA events on the context's
stream, max over ranks); `e2e` is the same metric through the C ABI with HOST buffers: fdc_begin_frame +
fdc_submit_*(host records) + fdc_end_frame + fdc_read_pixels(host), copies inside the timed region.
`--impl reference` times the CPU restatement of the reference's GL path (oracle/) on all host cores, whole frames.
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

WORKLOADS = {
    "cfg5_4k": "cfg5: 100k shadowed rounded rects + 20k glyph quads, 3840x2160, seed 5",
    "cfg5_8k": "cfg5 x2: 100k shadowed rounded rects + 20k glyph quads, 7680x4320, seed 5",
    "cfg2": "cfg2: renderlist_100 shape (300 boxes, shadows, elliptical corners, 1 backdrop blur), 1920x1080",
    "cfg3": "cfg3: 20k atlas glyph quads + MSDF/MTSDF star, 3840x2160",
    "cfg3_msdf": "cfg3: 20k atlas glyph quads + 20k MSDF glyph quads + MSDF/MTSDF star, 3840x2160",
    "cfg4": "cfg4: clip-mask table 180x12 (sub-clip), 3-stop gradients, 2 backdrop blurs, 3840x2160",
    "cfg4_rectmask": "cfg4: clip-mask table 180x12 (rect-mask), 2 backdrop blurs, 3840x2160",
}


def workload_trace(name: str):
    from figdraw_b200 import scenes_synth as ss

    if name not in WORKLOADS:
        raise SystemExit(f"unknown workload {name}")
    if name == "cfg5_4k":
        tr = ss.config_trace(5, 3840, 2160)
    elif name == "cfg5_8k":
        tr = ss.config_trace(5, 7680, 4320, scale=2.0)
    elif name == "cfg2":
        tr = ss.config_trace(2)
    elif name == "cfg3":
        tr = ss.config_trace(3)
    elif name == "cfg3_msdf":
        tr = ss.config_trace(3, msdf_glyphs=20000)
    elif name == "cfg4":
        tr = ss.config_trace(4)
    else:
        tr = ss.config_trace(4, rect_mask=True)
    return tr, WORKLOADS[name]


def bench_config(name: str, trace, world: int, gather: str = "nccl", band_policy: str = "balanced") -> dict:
    """The `config` object of the JSON line -- ONE function for both arms, so `--impl reference` reports the very same
    object as the GPU arm it is compared with."""
    if world == 1:
        part = "single GPU"
    elif gather in ("mc", "auto"):
        part = f"{world} tile-row bands, band all-gather fused into the shade kernel's copy-out (NVSwitch multicast stores)"
    elif gather == "symm-p2p":
        part = f"{world} tile-row bands, band all-gather fused into the shade kernel's copy-out (peer stores, symmetric memory)"
    elif gather == "p2p":
        part = f"{world} tile-row bands, band all-gather fused into the shade kernel (peer stores over NVLink)"
    elif gather == "ce":
        part = f"{world} tile-row bands, finished band slices copied to the peers by the copy engines (NVLink)"
    else:
        part = f"{world} tile-row bands + NCCL all-gather"
    if world > 1 and gather in ("mc", "auto", "symm-p2p"):
        part += ("; bands of equal tile-entry cost (profile of the previous frame)" if band_policy == "balanced" else "; bands of equal height")
    return {"workload": WORKLOADS[name], "frame": [trace.width, trace.height], "primitives": int(trace.n_draws),
            "l2": "flushed between steps (256 MiB fill)", "partition": part,
            "replay": "one CUDA-graph launch per frame (setup, binning, shade, barriers captured once)"}


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
    """CPU arm: the oracle port of the reference's GL path on all host threads.  Every step renders the WHOLE frame of the
    GPU arm's workload (no row sample, no extrapolation); warm-up is capped at one frame so K + W steps stay in minutes."""
    if rank != 0:
        return
    from oracle import oracle as orc

    name = args.workload or "cfg5_4k"
    trace, _desc = workload_trace(name)
    cores = orc.max_threads()
    o = orc.Oracle(trace.atlas_size)
    for _i, key, img in trace.images:
        o.put_image(key, img)
    W, H = trace.width, trace.height
    warm = min(args.warmup, 1)
    for _ in range(warm):
        o.render(W, H, trace.calls, clear=trace.clear, n_threads=cores)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        o.render(W, H, trace.calls, clear=trace.clear, n_threads=cores)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    val = W * H / dt / 1e6
    sample = f"{args.steps} whole frames of the same scene ({W}x{H}), {dt:.2f} s each (min {min(ts):.2f} s), {warm} warm-up frame(s)"
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": METRIC, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(name, trace, args.gpus, args.gather, args.bands),
            "note": "restated-reference CPU rasteriser (oracle port of the GL path, OpenMP over 64-row strips), not llvmpipe; "
                    "whole frames, nothing extrapolated",
            "cpu_baseline": {"value": round(val, 3), "unit": METRIC, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 3), "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "frames_per_s": round(1.0 / dt, 4)}
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


def executed_view(name: str, world: int, shade_ms: float, sm_mhz: float):
    """What the hardware executed for the dominant kernel, from the committed ncu --set full summary of this kernel
    (profiles/shade_ncu_summary.json: which commit, which command).  NOT measured in this run -- ncu cannot run inside a
    timed benchmark -- and labelled so; the live part is `issue_slots_per_cycle_live`: the stored instruction count over
    this run's kernel time and clock."""
    p = os.path.join(ROOT, "profiles", "shade_ncu_summary.json")
    if name != "cfg5_4k" or world != 1 or not os.path.exists(p):
        return None, None, None
    s = json.load(open(p))
    inst = float(s["warp_instructions"])
    cycles = shade_ms * 1e-3 * sm_mhz * 1e6
    ex = {"source": f"stored ncu --set full capture, {s['source']}", "captured_at_commit": s.get("commit"),
          "warp_instructions": inst, "issue_slots_per_cycle_ncu": s.get("issue_active"),
          "pipe_fma_pct_of_peak": s.get("pipe_fma_pct"), "pipe_alu_pct_of_peak": s.get("pipe_alu_pct"),
          "pipe_xu_pct_of_peak": s.get("pipe_xu_pct"), "kernel_us_under_ncu": s.get("duration_us"),
          # 4 schedulers per SM issue at most one warp instruction per cycle each
          "issue_slots_per_cycle_live": round(inst / (cycles * 148 * 4), 4) if cycles > 0 else None,
          "frac_executed": round(float(s.get("pipe_fma_pct") or 0.0) / 100.0, 4),
          "note": "frac_executed = share of the FP32 (FMA) pipe's issue capacity the kernel used under ncu; the kernel is "
                  "issue/latency-bound, the algorithmic `frac` above counts fragments it legitimately skips"}
    return ex, s.get("traffic_bytes_per_launch"), f"stored: dram__bytes_read.sum + dram__bytes_write.sum of one launch, {s['source']}"


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun warnings) print to fd 1; the driver wants ONE JSON line there.
    Point fd 1 at stderr for the run and keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


class Env:
    """torch / torch.distributed state shared by the measurements of one bench.py process."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            dist.barrier()
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.keep = []

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.dist is None:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]


def measure(env: Env, name: str, steps: int, warmup: int, headline: bool):
    """One workload on env.world GPUs.  Returns a dict of raw results on every rank (values reduced to max over ranks)."""
    torch, dist, args = env.torch, env.dist, env.args
    rank, world, dev = env.rank, env.world, env.dev
    from figdraw_b200 import bands
    from figdraw_b200.cuda_context import CudaContext, prepare_calls, prepared_upload_bytes

    trace, _desc = workload_trace(name)
    W, H = trace.width, trace.height
    ctx = CudaContext(atlasSize=trace.atlas_size, device=env.local_rank, rank=rank, nRanks=world)
    for _i, key, img in trace.images:
        ctx.putImage(key, img)
    band_rows, _layout = bands.band_layout(H, world)
    # a backdrop blur under a band partition reads halo rows out of the neighbours' framebuffers: peer mappings needed
    has_blur = bool((trace.calls["op"] == 13).any())
    gather_mode = args.gather if world > 1 else "none"
    symm_t = None

    def bind_symmetric(c):
        """Framebuffer (+ flags + record exchange area) in torch symmetric memory, bound to context `c` on every rank."""
        import torch.distributed._symmetric_memory as symm

        nbytes = ((W * band_rows * world * 4 + 255) & ~255) + 4096 + 64 * (len(trace.calls) + 1024)
        t = symm.empty(nbytes, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        mc_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
        ok = torch.tensor([1 if (mc_ptr or args.gather == "mc") else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            raise RuntimeError("no multicast mapping on some rank")
        t.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        c.bindSharedFramebuffer(t.data_ptr(), nbytes, [int(p) for p in hdl.buffer_ptrs], mc_ptr, W, band_rows * world)
        # kept for the life of the process: torch caches rendezvous handles per allocation, a freed block that is handed out
        # again for a larger tensor would come back with the stale (smaller) peer mappings
        env.keep.append((t, hdl))
        return t, mc_ptr

    if gather_mode in ("auto", "mc"):
        # Framebuffer in torch symmetric memory: every rank maps every copy, and behind an NVSwitch there is a multicast
        # mapping -- the shade kernel's copy-out then writes each finished chunk once and the switch delivers it to all
        # ranks (fdc_bind_shared_framebuffer).  No collective call in the frame loop; a flag barrier ends the frame.
        try:
            symm_t, mc_ptr = bind_symmetric(ctx)
            gather_mode = "mc" if mc_ptr else "symm-p2p"
        except Exception as e:  # noqa: BLE001
            if args.gather == "mc":
                raise
            sys.stderr.write(f"[bench] symmetric-memory framebuffer unavailable ({e!r}); falling back to NCCL all-gather\n")
            symm_t = None
            gather_mode = "nccl"
    use_p2p = world > 1 and symm_t is None and (gather_mode in ("p2p", "ce") or has_blur)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    token = None
    fb = None
    if symm_t is not None:
        fb = symm_t[: band_rows * world * W * 4].view(band_rows * world, W, 4)
    elif use_p2p:
        # Fused all-gather: every rank's shade kernel stores its finished pixels into all peers' framebuffers over
        # NVLink (CUDA IPC mappings); a one-element NCCL all-reduce on the same stream is the completion barrier.
        ctx.reserveFramebuffer(W, band_rows * world)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.framebufferIpcHandle())
        peers = [0 if r == rank else ctx.openPeerFramebuffer(handles[r]) for r in range(world)]
        ctx.setPeerFramebuffers(peers)
        if gather_mode == "ce":
            ctx.setPeerGather("copy", args.sub_bands)
        token = torch.zeros(1, dtype=torch.int32, device=dev)
        if gather_mode == "nccl":
            gather_mode = "p2p"
    else:
        # Framebuffer owned by torch so NCCL can all-gather the bands in place; rows padded to equal bands.
        fb = torch.zeros((band_rows * world, W, 4), dtype=torch.uint8, device=dev)
        ctx.bindFramebuffer(fb.data_ptr())

    calls_host = torch.from_numpy(trace.calls.view(np.uint8).reshape(-1, 128).copy()).pin_memory()
    calls_np = calls_host.numpy().view(trace.calls.dtype).reshape(-1)
    out_host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    out_np = out_host.numpy()
    keep_pinned = []
    records = args.records or "compact"

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
        if world == 1 or symm_t is not None:
            return  # single GPU, or fused into the shade kernel's copy-out (+ the frame's own flag barrier)
        if use_p2p:
            dist.all_reduce(token)  # all ranks' shade kernels (and their peer stores) are complete after this
        else:
            bands.allgather_bands(fb, rank, world)

    def submit(c, prep):
        c.beginFrame((W, H), clearMain=trace.clear is not None, clearMainColor=trace.clear or (1, 1, 1, 1))
        c.submitPrepared(prep)
        c.endFrame()

    trace_host = os.environ.get("FDC_E2E_TRACE", "0") != "0"  # host-side split of a single frame's latency, to stderr
    host_t = [0.0, 0.0, 0.0, 0]

    def frame_e2e():
        t0 = time.perf_counter()
        submit(ctx, prepared)
        t1 = time.perf_counter()
        with torch.cuda.stream(stream):
            gather()
        y0, y1 = ctx.bandRows() if world > 1 else (0, H)
        if trace_host:
            ctx.sync()
        t2 = time.perf_counter()
        ctx.readPixels((0, y0, W, y1 - y0), out=out_np[y0:y1])
        t3 = time.perf_counter()
        host_t[0] += t1 - t0; host_t[1] += t2 - t1; host_t[2] += t3 - t2; host_t[3] += 1

    # first frame: uploads the recording, allocates everything; a bin-list overflow on ANY rank re-runs it on all
    submit(ctx, prepared)
    bands.resolve_across_ranks(ctx, world, dist)
    # Bands of equal COST instead of equal height: the frame just rendered tells every rank how many tile entries each
    # of its tile rows holds; the summed profile is split into `world` contiguous bands of equal cost and every rank
    # adopts the same boundaries (fdc_get_tile_row_costs / fdc_set_band_tile_rows).  Only with a framebuffer the ranks
    # share: an all-gather of equal slices does not apply to unequal bands.
    band_bounds = None
    if world > 1 and symm_t is not None and args.bands == "balanced" and not has_blur:
        band_bounds = bands.rebalance_across_ranks(ctx, world, (W + 15) // 16, dist)
        submit(ctx, prepared)
        bands.resolve_across_ranks(ctx, world, dist)
    frame_e2e()
    launches_per_frame = int(ctx.frameStats().n_launches)

    def timed_loop(fn_step, n):
        """Each step individually bracketed by CUDA events on the context stream; L2 flushed between steps."""
        evs = []
        for _ in range(n):
            with torch.cuda.stream(stream):
                env.flush.fill_(0)
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

    for _ in range(warmup):
        step_resident()
    env.barrier()
    sampler = ClockSampler(env.local_rank) if (headline and rank == 0) else None
    if sampler:
        sampler.start()
    env.barrier()
    times = timed_loop(step_resident, steps)
    env.barrier()
    clocks = sampler.stop() if sampler else None
    ms_step = float(np.sum(times)) / steps
    # per-phase times (setup + binning / shade) of the same frame: replayed launch by launch, because the timed steps above
    # are single CUDA-graph launches and events inside a graph cannot be timed
    ctx.setReplayGraph(False)
    phase = []
    for _ in range(5):
        with torch.cuda.stream(stream):
            env.flush.fill_(0)
        step_resident()
        st_ = ctx.frameStats()
        phase.append((float(st_.shade_ms), float(st_.bin_ms), float(st_.gpu_ms)))
    ctx.setReplayGraph(True)
    stats = ctx.frameStats()
    shade_plain, bin_plain, gpu_plain = (float(v) for v in np.median(np.array(phase), axis=0))

    # e2e: host records in, host pixels out, wall clock bracketed by synchronisation (copies are inside)
    for _ in range(2):
        frame_e2e()
    env.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(5, min(steps, 20))
    for _ in range(e2e_steps):
        frame_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if trace_host and host_t[3]:
        sys.stderr.write("[bench] rank %d single-frame split: submit (host) %.3f ms, wait for the frame %.3f ms, read-back %.3f ms\n"
                         % (rank, host_t[0] / host_t[3] * 1e3, host_t[1] / host_t[3] * 1e3, host_t[2] / host_t[3] * 1e3))

    # e2e, pipelined (N = 1): three contexts on three streams, like a three-image swap chain.  Every step still copies
    # its own records host->device and its own frame device->host; in steady state frame k's readback, frame k+1's
    # kernels and frame k+2's upload are in flight together.  Throughput over the steps, not latency.
    e2e_pipe_ms, depth = None, 1
    present_ms, present_ok = None, None
    y0b, y1b = ctx.bandRows() if world > 1 else (0, H)
    if world == 1 or symm_t is not None:
        ring = [(ctx, prepared, out_np)]
        extra_ctx = []
        for _ in range(max(1, int(os.environ.get("FDC_E2E_RING", "3")) - 1)):
            c2 = CudaContext(atlasSize=trace.atlas_size, device=env.local_rank, rank=rank, nRanks=world)
            if world > 1:
                bind_symmetric(c2)  # its own shared framebuffer, flags and record exchange area
                if band_bounds is not None:
                    c2.setBandTileRows(band_bounds)
            for _i, key, img in trace.images:
                c2.putImage(key, img)
            o2 = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            k2 = calls_host.clone().pin_memory()
            ring.append((c2, prepare_pinned(k2.numpy().view(trace.calls.dtype).reshape(-1)), o2.numpy()))
            extra_ctx.append((c2, o2, k2))
        depth = len(ring)

        # FDC_E2E_ASYNC_READ=1 queues every read-back right behind its frame (fdc_read_pixels_async).  Measured slower on
        # this platform (0.97 vs 0.85 ms per 4K frame, same box): the device-to-host copy then runs against the next
        # frames' uploads the whole time, and the PCIe link does worse in both directions at once than taking turns.
        async_read = os.environ.get("FDC_E2E_ASYNC_READ", "0") != "0"

        def queue_frame(slot):
            # records up, kernels, and the read-back of this rank's band queued right behind them (fdc_read_pixels_async):
            # the device-to-host copy starts the moment the frame is finished, not when the host gets round to asking
            submit(slot[0], slot[1])
            if async_read:
                slot[0].readPixelsAsync(slot[2][y0b:y1b], (0, y0b, W, y1b - y0b))

        def pipelined(n):
            for k in range(min(depth - 1, n)):
                queue_frame(ring[k % depth])
            for k in range(n):
                if k + depth - 1 < n:
                    queue_frame(ring[(k + depth - 1) % depth])
                cur = ring[k % depth]
                if async_read:
                    cur[0].sync()  # frame k's pixels are in host memory
                else:
                    cur[0].readPixels((0, y0b, W, y1b - y0b), out=cur[2][y0b:y1b])

        for c2, prep2, _o in ring[1:]:  # size the new contexts' buffers (a bin-list overflow re-runs on every rank)
            submit(c2, prep2)
            bands.resolve_across_ranks(c2, world, dist)
        pipelined(6)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipelined(e2e_steps)
        torch.cuda.synchronize()
        e2e_pipe_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        if not all(bool(np.array_equal(out_np[y0b:y1b], r[2][y0b:y1b])) for r in ring[1:]):
            raise SystemExit("pipelined contexts produced different frames")
        # e2e, presenter variant (N = 1): the frame stays on the device in an exported allocation a presenter imports
        # (fdc_export_framebuffer) -- records still cross PCIe every step, pixels never do.
        if world == 1 and headline:
            try:
                for c, _p, _o in ring:
                    c.exportFramebuffer(W, H)

                def presented(n):
                    for k in range(n):
                        c, prep, _o = ring[k % depth]
                        submit(c, prep)  # beginFrame waits for the frame this context rendered `depth` steps ago
                    for c, _p, _o in ring:
                        c.sync()

                presented(6)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                presented(e2e_steps)
                torch.cuda.synchronize()
                present_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
                present_ok = bool(np.array_equal(ring[0][0].readPixels(), out_np))
            except Exception as e:  # noqa: BLE001
                present_ms, present_ok = None, repr(e)[:200]
        for c2, _o, _k in extra_ctx:
            c2.close()

    gathered_ok, single_gpu_ms = None, None
    if world > 1:
        # the gathered frame on rank 0 must equal a single-context render of the whole frame, bit for bit
        ctx.replayFrame()
        with torch.cuda.stream(stream):
            gather()
        torch.cuda.synchronize()
        if rank == 0:
            whole = np.empty((H, W, 4), dtype=np.uint8)
            if fb is None:
                ctx._ck(ctx._lib.fdc_read_pixels(ctx._h, 0, 0, W, H, whole.ctypes.data))
            else:
                whole[:] = fb[:H].cpu().numpy()
            ref_ctx = CudaContext(atlasSize=trace.atlas_size, device=env.local_rank)
            for _i, key, img in trace.images:
                ref_ctx.putImage(key, img)
            submit(ref_ctx, prepared)
            gathered_ok = bool(np.array_equal(ref_ctx.readPixels(), whole))
            # the same frame on ONE GPU (this rank's), timed in the same process
            single_ms = []
            for _ in range(9):
                env.flush.fill_(0)
                torch.cuda.synchronize()
                ref_ctx.replayFrame()
                single_ms.append(float(ref_ctx.frameStats().gpu_ms))
            single_gpu_ms = float(np.median(single_ms[2:]))
            ref_ctx.close()
        env.barrier()
    ms_step, e2e_ms, shade_ms, bin_ms, e2e_pipe_max = env.max_over_ranks([ms_step, e2e_ms, shade_plain, bin_plain, e2e_pipe_ms or 0.0])
    if e2e_pipe_ms is not None:
        e2e_pipe_ms = e2e_pipe_max
    res = {"name": name, "trace": trace, "W": W, "H": H, "ms_step": ms_step, "e2e_ms": e2e_ms, "e2e_pipe_ms": e2e_pipe_ms,
           "depth": depth, "shade_ms": shade_ms, "bin_ms": bin_ms, "clocks": clocks, "launches_per_frame": launches_per_frame,
           "n_tile_entries": int(stats.n_tile_entries), "h2d": int(prepared_upload_bytes(prepared)), "out_np": out_np, "band_bounds": band_bounds,
           "sharded_upload": bool(world > 1 and symm_t is not None), "gpu_plain_ms": gpu_plain, "present_ms": present_ms, "present_ok": present_ok,
           "calls_np": calls_np, "gathered_ok": gathered_ok, "single_gpu_ms": single_gpu_ms, "use_p2p": use_p2p,
           "gather": gather_mode}
    ctx.close()
    return res


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-8k", action="store_true", help="N > 1: skip the extra cfg5_8k measurement")
    ap.add_argument("--gather", default="auto", choices=["auto", "mc", "p2p", "nccl", "ce"],
                    help="N>1: how the bands reach every rank -- NVSwitch multicast stores fused into the shade kernel's copy-out "
                         "(mc; needs torch symmetric memory), peer stores over IPC mappings (p2p), NCCL all-gather after the frame "
                         "(nccl), copy engines shipping band slices (ce); auto = mc when available, else nccl")
    ap.add_argument("--sub-bands", type=int, default=4)
    ap.add_argument("--bands", default="balanced", choices=["balanced", "equal"],
                    help="N > 1 with a shared framebuffer: bands of equal tile-entry cost (default) or of equal height")
    ap.add_argument("--records", default=None, choices=["compact", "full"],
                    help="e2e upload: 64-byte fdc_rect64 records for rounded rects with circular corners, or 128-byte fdc_call only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return

    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    env = Env(args)
    world = env.world
    name = args.workload or "cfg5_4k"
    r = measure(env, name, args.steps, args.warmup, headline=True)
    extra_8k = None
    if world > 1 and args.workload is None and not args.no_8k:
        extra_8k = measure(env, "cfg5_8k", max(5, min(args.steps, 20)), 3, headline=False)

    if rank == 0:
        trace, W, H = r["trace"], r["W"], r["H"]
        ms_step, shade_ms, bin_ms, clocks = r["ms_step"], r["shade_ms"], r["bin_ms"], r["clocks"]
        peaks, peak_src = measured_peaks()
        mpx = W * H / 1e6
        value = mpx / (ms_step * 1e-3)
        # roofline of the dominant kernel (shade): algorithmic flops from the oracle's exact fragment counts
        cpu_base = None
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        peak_fp32 = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s per GPU, non-tensor FP32 at the clock seen under load
        from oracle import oracle as orc

        counts = orc.count_fragments(trace)  # exact fragments per mode (checker; counts only, nothing is shaded)
        n_frag = int(counts.sum())
        flops = algorithmic_flops(counts)
        bytes_alg = W * H * 4 + trace.n_draws * 128 * 2 + r["n_tile_entries"] * 8 * world
        ach = flops / (shade_ms * 1e-3) / 1e12 / world  # per GPU: every rank shades 1/world of the frame in shade_ms
        executed, traffic, traffic_src = executed_view(name, world, shade_ms, sm_mhz)
        roof = {"bound": "fp32", "kernel": "shade_kernel", "achieved": round(ach, 3), "peak": round(peak_fp32, 2),
                "unit": "TFLOP/s", "frac": round(ach / peak_fp32, 4), "traffic": traffic, "traffic_source": traffic_src,
                "definition": "ALGORITHMIC (SURVEY 8d): every fragment the reference's GL path would shade x flops per fragment "
                              "by mode / shade-kernel time; the kernel provably skips occluded and trivially covered fragments, "
                              "so this is not a utilisation figure (it can exceed 1) -- see `executed`",
                "executed": executed,
                "algorithmic_flops_per_frame": flops, "fragments_per_frame": n_frag,
                "peak_source": f"148 SM x 128 lanes x 2 x {sm_mhz:.0f} MHz (SM clock sampled during the timed region), per GPU",
                "shade_ms": round(shade_ms, 4), "bin_ms": round(bin_ms, 4),
                "phase_note": "shade_ms / bin_ms: the same frame replayed launch by launch (%.4f ms per frame that way); the timed "
                              "steps are single CUDA-graph launches" % r["gpu_plain_ms"],
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
            d = np.abs(r["out_np"].astype(np.int16) - ref_img.astype(np.int16)).max(axis=2)
            cpu_base = {"value": round(mpx / cpu_s, 3), "unit": METRIC, "cores": cores, "kind": "port",
                        "sample": f"1 full frame of the same scene ({W}x{H}, {n_frag} fragments) in {cpu_s:.2f} s"}
            roof["parity_vs_oracle"] = {"max_abs_diff_lsb": int(d.max()), "pixels_differing": int((d > 0).sum())}
        e2e_best = r["e2e_pipe_ms"] or r["e2e_ms"]
        line = {"metric": METRIC, "value": round(value, 2), "unit": METRIC, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": bench_config(name, trace, world, r["gather"], "balanced" if r.get("band_bounds") else "equal"),
                "frames_per_s": round(1e3 / ms_step, 2),
                "e2e": {"value": round(mpx / (e2e_best * 1e-3), 2), "unit": METRIC,
                        "ms_per_step": round(e2e_best, 4), "latency_ms": round(r["e2e_ms"], 4),
                        "mode": (f"{r['depth']} contexts in flight: step k's readback, step k+1's kernels and step k+2's upload overlap"
                                 if r["e2e_pipe_ms"] else "one frame at a time"),
                        "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": int(W * H * 4),
                        "bytes_note": ("whole job: every rank uploads 1/n of each long run of compact records over its own PCIe link "
                                       "and pushes it to all ranks over NVLink; every rank reads its own band back"
                                       if r["sharded_upload"] else
                                       ("every rank uploads the whole stream and reads its own band back" if world > 1 else "one GPU"))},
                "gpu_launches": r["launches_per_frame"] * args.steps, "launches_per_frame": r["launches_per_frame"],
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base}
        if r["present_ms"]:
            line["e2e_present"] = {"value": round(mpx / (r["present_ms"] * 1e-3), 2), "unit": METRIC, "ms_per_step": round(r["present_ms"], 4),
                                   "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 0, "frame_equals_readback": r["present_ok"],
                                   "note": "same steps as e2e but the frame is left in an exported device allocation a presenter "
                                           "imports (fdc_export_framebuffer): no read-back.  Not the headline: e2e above reads pixels to the host"}
        if r["gathered_ok"] is not None:
            line["gathered_frame_equals_single_gpu"] = r["gathered_ok"]
            line["single_gpu_same_workload"] = {"ms_per_step": round(r["single_gpu_ms"], 4),
                                                "value": round(mpx / (r["single_gpu_ms"] * 1e-3), 2), "unit": METRIC,
                                                "note": "rank 0's GPU alone on the same frame in the same process (no gather)"}
        if r.get("band_bounds"):
            line["band_tile_rows"] = r["band_bounds"]
        if extra_8k is not None:
            x = extra_8k
            mpx8 = x["W"] * x["H"] / 1e6
            line["cfg5_8k"] = {"workload": WORKLOADS["cfg5_8k"], "ms_per_step": round(x["ms_step"], 4),
                               "value": round(mpx8 / (x["ms_step"] * 1e-3), 2), "unit": METRIC,
                               "single_gpu_ms": round(x["single_gpu_ms"], 4),
                               "efficiency": round(x["single_gpu_ms"] / (world * x["ms_step"]), 4),
                               "shade_ms": round(x["shade_ms"], 4), "bin_ms": round(x["bin_ms"], 4),
                               "e2e_ms_per_step": round(x["e2e_pipe_ms"] or x["e2e_ms"], 4), "e2e_latency_ms": round(x["e2e_ms"], 4),
                               "h2d_bytes_per_step": x["h2d"],
                               "gathered_frame_equals_single_gpu": x["gathered_ok"], "band_tile_rows": x.get("band_bounds"),
                               "note": "BASELINE configs[4] (all sizes x2) on the same ranks; efficiency = single_gpu_ms / (n_gpus x ms_per_step)"}
        if world == 1 and name in ("cfg5_4k", "cfg5_8k"):
            line["native_frontend"] = native_frontend_probe(name, r["calls_np"])
        print(json.dumps(line), file=real_stdout, flush=True)
    if env.dist is not None:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
