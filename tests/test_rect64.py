"""fdc_rect64, the 64-byte form of the common rounded-rect draw: pack -> expand is the identity on the 128-byte record
(host helpers of the C library; the setup kernel runs the same expander on the device)."""
import ctypes

import numpy as np
import pytest

from figdraw_b200 import abi, scenes, scenes_fuzz, scenes_synth as ss
from figdraw_b200.cuda_context import pack_rects64, prepare_calls, prepared_upload_bytes


def expand_c(rects):
    lib = abi.load_library()
    out = np.zeros(len(rects), dtype=abi.CALL_DTYPE)
    for i in range(len(rects)):
        lib.fdc_expand_rect64(rects[i:i + 1].ctypes.data, out[i:i + 1].ctypes.data)
    return out


def pack_c(calls):
    lib = abi.load_library()
    ok = np.zeros(len(calls), dtype=bool)
    out = np.zeros(len(calls), dtype=abi.RECT64_DTYPE)
    for i in range(len(calls)):
        ok[i] = bool(lib.fdc_pack_rect64(calls[i:i + 1].ctypes.data, out[i:i + 1].ctypes.data))
    return ok, out[ok]


def traces():
    yield ss.config_trace(2, 1280, 720)
    yield ss.config_trace(4, 1280, 720, rows=20, cols=6)
    yield ss.config_trace(5, 1280, 720, n_rects=1500, n_glyphs=300)
    for name in sorted(scenes.GOLDEN_SCENES):
        yield scenes.golden_trace(name)
    for seed in range(6):
        yield scenes_fuzz.random_trace(seed)


def test_pack_expand_round_trip_and_c_python_agreement():
    n_packed = 0
    for tr in traces():
        calls = tr.calls
        ok_py, r_py = pack_rects64(calls)
        ok_c, r_c = pack_c(calls)
        assert np.array_equal(ok_py, ok_c)
        assert r_py.tobytes() == r_c.tobytes()
        back = expand_c(r_py)
        assert back.tobytes() == np.ascontiguousarray(calls[ok_py]).tobytes()
        n_packed += int(ok_py.sum())
    assert n_packed > 4000


def test_prepare_calls_compact_covers_the_frame():
    tr = ss.config_trace(5, 1280, 720, n_rects=3000, n_glyphs=600)
    calls, runs = prepare_calls(tr.calls, compact=True)
    covered = 0
    for run in runs:
        assert run[1] == covered
        covered = run[2]
        if run[0] == "rects64":
            assert len(run[3]) == run[2] - run[1] >= 256
            assert expand_c(run[3][:5]).tobytes() == np.ascontiguousarray(calls[run[1]:run[1] + 5]).tobytes()
    assert covered == len(calls)
    assert prepared_upload_bytes((calls, runs)) < 0.6 * calls.nbytes
