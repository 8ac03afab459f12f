"""`path::modify_path` on the GPU (csrc/path.cu) against the CPU oracle (oracle/path.cpp): converged costs bit-exact,
predecessors identical, the (magnitude, rotation) list bit-exact; `Path::serialize` against struct.pack.
The reference function itself panics on every input (path.rs:29-30,38,99,105) - both sides implement the intent mode."""
import struct

import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def _scene(tod, W, H, seed, weights_mode, balls=None):
    depth = synth.depth_frames(1, W=W, H=H, seed=seed)
    target = synth.target_frames(1, W=W, H=H, seed=seed + 1)
    sb = tod.SceneBuilder(width=W, height=H, max_batch=1, weights_mode=weights_mode)
    sb.append_batch(depth, target, want=())
    s = sb.materialize(0)
    if balls is not None:
        s.balls[:3] = balls
    return s


@pytest.mark.parametrize("W,H,weights_mode,balls", [(640, 480, 0, None), (640, 480, 1, [(100, 60), (500, 200), (320, 400)]),
                                                    (320, 240, 1, [(10, 10), (300, 20), (160, 120)]), (100, 75, 0, [(50, 40), (0, 0), (0, 0)])])
def test_modify_path_matches_oracle(tod, W, H, weights_mode, balls):
    s = _scene(tod, W, H, 90 + W, weights_mode, balls)
    want = oracle.path_modify(s.height, s.pos, s.balls, s.connections, W, H, cap=8192)
    path, cost, pred = tod.modify_path(s, height_px=H, want_fields=True)
    assert np.array_equal(cost.view(np.uint32), want["cost"].view(np.uint32)), "%d costs differ" % (cost.view(np.uint32) != want["cost"].view(np.uint32)).sum()
    assert np.array_equal(pred, want["pred"])
    if want["directions"] is None:
        assert path is None
    else:
        assert np.array_equal(path.directions.view(np.uint32), want["directions"].view(np.uint32))
        assert len(path.directions) > 0
        assert path.serialize() == oracle.path_serialize(path.created, path.directions)


def test_unreachable_start_and_serialize(tod):
    W, H = 64, 48
    n = W * H
    from tod_b200 import Scene
    s = Scene(np.zeros(n, np.float32), np.zeros((n, 3), np.float32), np.zeros((100, 2), np.int32), -np.ones((n, 8), np.float32), W)
    assert tod.modify_path(s, height_px=H) is None          # every connection is the -1 "absent" marker: nothing is reachable
    p = tod.Path(1700000000, np.array([[1.5, 0.0], [2.0, 0.25]], np.float32))
    assert p.serialize() == struct.pack(">Qffff", 1700000000, 1.5, 0.0, 2.0, 0.25)   # path.rs:17-21
    assert tod.Path(7).serialize() == struct.pack(">Q", 7)
    assert tod.lib().tod_path_reference_panics() == 1
