"""compute-sanitizer is closed on this pool, so lost updates in the new shared-memory stamp (lanes of a warp update a
shared tile with plain read / max / write, ordered by __syncwarp) are hunted the blunt way: the same dense batch many
times through one handle and through fresh handles - every byte identical, every time, and equal to the oracle."""
import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def test_scene_bytes_are_identical_over_repeats(tod):
    rng = np.random.default_rng(5)
    depth = rng.integers(300, 4001, (6, 480, 640)).astype(np.uint16)       # noise: every pixel lands on its own row
    depth[3:] = synth.depth_frames(3, seed=6)
    target = synth.target_frames(6, seed=7, blobs=12)
    sb = tod.SceneBuilder(max_batch=6)
    first = sb.append_batch(depth, target)
    for f in (0, 4):
        m, _ = oracle.pt_cloud(depth[f], target[f])
        assert np.array_equal(first["map"][f], m)
    for rep in range(12):
        h = sb if rep % 3 else tod.SceneBuilder(max_batch=6)
        got = h.append_batch(depth, target)
        for k in ("map", "world", "conn0", "conn1"):
            assert np.array_equal(got[k].view(np.uint32), first[k].view(np.uint32)), "%s changed on repeat %d" % (k, rep)
