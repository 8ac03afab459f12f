"""The bench line's contract (keys the driver and the judge read), checked on the committed line of the current state, and the
reference arm run for one bounded step on the host."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest_line():
    def key(p):
        b = os.path.basename(p)
        return (int(b.split("_r")[1].split("_")[0]), int("".join(c for c in b.split("_v")[1].split(".")[0].split("_")[0] if c.isdigit())))
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r*_v*.json")), key=key)
    full = [f for f in files if "_n" not in os.path.basename(f) and "noscene" not in f]
    return json.load(open(full[-1]))


def test_committed_line_has_every_contract_key():
    d = _latest_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "yolact_frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    assert d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if "check" in d:   # round 2: the line carries the verdict of its own bytes against the oracle's committed CRCs
        assert d["check"]["ok"] is True and d["check"]["timed_step_tiles_differing_from_oracle"] == 0 and d["check"]["timed_step_tiles_checked"] == 64


def test_reference_arm_one_step():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "yolact_frames_per_sec" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
