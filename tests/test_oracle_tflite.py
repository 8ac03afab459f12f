"""The C++ TFLite-integer oracle against the independent numpy restatement that generated the synthetic model
(oracle/synth_model.py keeps the calibration activations of every tensor): two implementations written from the
published rules must agree bit for bit on every tensor.  PARITY UNPINNED against a real TFLite build."""
import numpy as np

import oracle
from oracle import synth_model


def test_cpp_oracle_matches_numpy_restatement_every_tensor():
    g, blob = synth_model.build_frc_model(small=True)
    _, small = synth_model.ensure_models()
    assert open(small, "rb").read() == blob, "the generator is not deterministic"
    m = oracle.Model(small)
    calib = synth_model.calib_images()[:, :64, :64]
    for img in range(2):
        m.invoke(calib[img])
        checked = 0
        for code, ins, outs, _ in g.ops:
            t = g.tensors[outs[0]]
            want = t.val[img].reshape(-1)
            got = m.tensor(t.idx).reshape(-1).astype(np.int64)
            assert np.array_equal(got, want), "op code %d tensor %s" % (code, t.name)
            checked += 1
        assert checked == 141


def test_operator_histogram_matches_reference_log():
    """/root/reference/data/FRC_model_edgetpu.log:7-19"""
    from collections import Counter
    full, _ = synth_model.ensure_models()
    m = oracle.Model(full)
    hist = Counter(m.op_code(i) for i in range(m.num_ops))
    assert hist == {3: 64, 4: 15, 114: 19, 22: 15, 0: 11, 28: 5, 34: 5, 23: 3, 2: 3, 19: 1}
    shapes = [m.tensor_info(t)["shape"] for t in m.outputs]
    assert shapes[4] == (1, 28, 28, 81)  # yolact.rs:91,108,118: results[4] is 28*28 chunks of 81


def test_golden_outputs_unchanged():
    import os
    import zlib
    from tests import synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))
    _, small = synth_model.ensure_models()
    m = oracle.Model(small)
    tiles = synth.rgb_tiles(2, S=64, seed=103)
    for t in range(2):
        m.invoke(tiles[t])
        got = np.array([zlib.crc32(m.tensor(m.op_output(i)).tobytes()) for i in range(m.num_ops)], np.uint32)
        assert np.array_equal(got, g["small_all_t%d" % t])
