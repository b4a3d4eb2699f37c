"""Real-time path under contention on a real GPU (tools/native/stress_probe.cpp, C ABI only): the audio thread pushes
512-sample stereo blocks while the GUI thread fetches columns and keeps changing range / palette / recolouring the ring."""
import json
import pathlib
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_push_latency_and_columns_under_gui_contention():
    exe = ROOT / "tools" / "native" / "stress_probe"
    if not exe.exists():
        subprocess.run(["make", "-C", str(ROOT / "tools" / "native")], check=True)
    out = subprocess.run([str(exe), "3000", "150"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    print(r)
    assert r["gui_errors"] == 0
    assert r["incomplete_columns"] == 0 and r["out_of_order"] == 0
    assert r["fetched_columns"] == r["columns"] >= 3000 - 3  # one column per block (the first ones start in the pre-roll zeros)
    assert r["db_ring_identical"] is True                   # bit-identical to an undisturbed single-threaded engine
    assert r["gui_ticks"] > 100
    # north star: the 512-sample block path stays under 100 us; the push alone must do so at p99 even while the GUI thread
    # recolours the ring and swaps palettes
    assert r["push_p99_us"] < 100.0, r
