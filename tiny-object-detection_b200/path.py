"""Host-side mirror of the reference's `Path` (src/path.rs:11-22) and `modify_path` (path.rs:25-120) over the C ABI.

    Path.serialize()        <- Path::serialize                (path.rs:17-21)
    modify_path(scene)      <- path::modify_path(path, scene) (path.rs:25-120), intent mode

The reference function panics on every input and is unreachable (see include/tod.h); `REFERENCE_PANICS` says so.
"""
import ctypes as C
import time
from dataclasses import dataclass, field

import numpy as np

from ._lib import check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


REFERENCE_PANICS = True


@dataclass
class Path:
    created: int = 0                                                            # seconds since the epoch (SystemTime, path.rs:12)
    directions: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.float32))   # (magnitude, rotation), path.rs:14

    def serialize(self):
        d = np.ascontiguousarray(self.directions, np.float32).reshape(-1, 2)
        out = np.zeros(8 + 8 * len(d), np.uint8)
        n = C.c_size_t(0)
        check(lib().tod_path_serialize(int(self.created), _ptr(d), len(d), _ptr(out), out.nbytes, C.byref(n)))
        return out[:n.value].tobytes()


def modify_path(scene, height_px=480, device=0, cap=8192, want_fields=False):
    """`modify_path` on a materialised Scene (tod_b200.Scene): returns the new Path, or None when the start node cannot
    reach a ball; with want_fields=True also the converged cost / predecessor arrays."""
    W = scene.width
    n = W * height_px
    h = np.ascontiguousarray(scene.height, np.float32).reshape(n)
    pos = np.ascontiguousarray(scene.pos, np.float32).reshape(n, 3)
    balls = np.ascontiguousarray(scene.balls, np.int32).reshape(-1, 2)
    conn = np.ascontiguousarray(scene.connections, np.float32).reshape(n, 8)
    cost = np.zeros(n, np.float32) if want_fields else None
    pred = np.zeros(n, np.int32) if want_fields else None
    dirs = np.zeros((cap, 2), np.float32)
    k = C.c_int32(0)
    check(lib().tod_path_modify(int(device), W, height_px, _ptr(h), _ptr(pos), _ptr(balls), _ptr(conn), _ptr(cost), _ptr(pred), _ptr(dirs), cap, C.byref(k)))
    path = None if k.value < 0 else Path(int(time.time()), dirs[:min(k.value, cap)].copy())
    return (path, cost, pred) if want_fields else path
