"""Holds the oracle (and on a GPU box the CUDA path) to the bytes of a REAL TFLite interpreter, when
tools/pin_with_tflite.py has been run somewhere that has one (tests/golden/tflite_real.npz).  In the build container
neither TFLite nor the model blob exists, so these skip and DESIGN.md says "unpinned vs real TFLite"."""
import hashlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "tflite_real.npz")


def _golden_and_model(models):
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/tflite_real.npz absent: run tools/pin_with_tflite.py where a TFLite interpreter exists")
    g = np.load(GOLD)
    for path in (os.path.join(ROOT, "data", "FRC_model.tflite"),) + tuple(models):
        if os.path.exists(path) and hashlib.sha256(open(path, "rb").read()).hexdigest() == str(g["model_sha256"]):
            return g, path
    pytest.skip("no local model file matches the one the TFLite golden was made with")


def test_pin_script_reports_missing_interpreter():
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pin_with_tflite.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode in (0, 2), out.stderr[-300:]
    assert "pin_with_tflite" in out.stdout


def test_oracle_matches_real_tflite(models):
    import oracle
    g, path = _golden_and_model(models)
    m = oracle.Model(path)
    for t in range(len(g["tiles"])):
        m.invoke(g["tiles"][t], threads=8)
        for k, ti in enumerate(m.outputs):
            assert np.array_equal(m.tensor(ti).reshape(-1), g["out%d" % k][t].reshape(-1)), "tile %d output %d differs from TFLite" % (t, k)
        if t == 0:
            for name in g.files:
                if name.startswith("t") and name[1:].isdigit():
                    ti = int(name[1:])
                    assert np.array_equal(m.tensor(ti).reshape(-1), g[name].reshape(-1)), "tensor %d differs from TFLite" % ti


@pytest.mark.gpu
def test_cuda_matches_real_tflite(tod, models):
    g, path = _golden_and_model(models)
    n = len(g["tiles"])
    res = tod.Yolact.init(path, max_tiles=n).infer_tiles(g["tiles"], detections=False)
    for k in range(len(res["outputs"])):
        assert np.array_equal(res["outputs"][k].reshape(n, -1), g["out%d" % k].reshape(n, -1)), "output %d differs from TFLite" % k
