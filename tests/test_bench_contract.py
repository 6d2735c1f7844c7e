"""bench.py's host-side helpers (no GPU): both arms report the SAME `config` object, the executed-work block is the
committed ncu summary labelled as stored, the algorithmic flop count follows the oracle's fragment counts, and the
reference arm prints one contract-shaped JSON line (a tiny frame so the CPU suite stays short)."""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _Trace:
    width, height, n_draws = 3840, 2160, 245012


def test_both_arms_report_the_same_config():
    b = _bench()
    one = b.bench_config("cfg5_4k", _Trace, 1)
    assert one["partition"] == "single GPU" and one["frame"] == [3840, 2160] and one["primitives"] == 245012
    assert one["workload"] == b.WORKLOADS["cfg5_4k"] and "flushed" in one["l2"]
    # the GPU arm resolves `auto` to the multicast gather; the reference arm only knows the flags -- same text either way
    ours = b.bench_config("cfg5_4k", _Trace, 8, "mc", "balanced")
    ref = b.bench_config("cfg5_4k", _Trace, 8, "auto", "balanced")
    assert ours == ref and "8 tile-row bands" in ours["partition"] and "equal tile-entry cost" in ours["partition"]
    assert "equal height" in b.bench_config("cfg5_4k", _Trace, 8, "mc", "equal")["partition"]
    assert "model" not in ours  # a rasteriser has no model keys


def test_executed_view_is_the_stored_capture_labelled_as_such():
    b = _bench()
    ex, traffic, src = b.executed_view("cfg5_4k", 1, 0.432, 1965.0)
    stored = json.load(open(os.path.join(ROOT, "profiles", "shade_ncu_summary.json")))
    assert ex["warp_instructions"] == stored["warp_instructions"] and ex["captured_at_commit"] == stored["commit"]
    assert "stored" in ex["source"] and "stored" in src and traffic == stored["traffic_bytes_per_launch"]
    # live issue rate = stored instruction count over THIS run's kernel time: 4 schedulers x 148 SMs, <= 1 per cycle each
    assert 0.5 < ex["issue_slots_per_cycle_live"] <= 1.0
    assert abs(ex["frac_executed"] - stored["pipe_fma_pct"] / 100.0) < 1e-3
    assert b.executed_view("cfg5_4k", 8, 0.1, 1965.0) == (None, None, None)  # only the single-GPU headline workload
    assert b.executed_view("cfg2", 1, 0.1, 1965.0) == (None, None, None)


def test_algorithmic_flops_follow_the_fragment_counts():
    b = _bench()
    counts = np.zeros(2 * b.N_MODES, dtype=np.float64)
    assert b.algorithmic_flops(counts) == 0.0
    mode = next(iter(b.FLOPS))
    counts[mode] = 1000.0
    base = b.algorithmic_flops(counts)
    assert base == 1000.0 * b.FLOPS[mode]
    counts[b.N_MODES + mode] = 10.0  # fragments with a 3-stop gradient cost the extra evaluation
    assert b.algorithmic_flops(counts) == base + 10.0 * (b.FLOPS[mode] + b.FLOPS_GRADIENT_EXTRA)


def test_measured_peaks_fall_back_when_the_driver_file_is_absent(tmp_path, monkeypatch):
    b = _bench()
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    peaks, kind = b.measured_peaks()
    assert kind == "fallback" and peaks["hbm_gbs"] > 0
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6545.3, "sm_max_mhz": 1965.0}))
    peaks, kind = b.measured_peaks()
    assert kind == "measured" and peaks["hbm_gbs"] == 6545.3
