"""Host-side mirror of the reference's scene builder (src/scene.rs:122-331) over the C ABI.

    Scene                         <- struct Scene { height, pos, balls, connections }   (scene.rs:122-132)
    SceneBuilder.append_scene()   <- append_scene(...)                                  (scene.rs:147-331)

`append_scene` keeps the reference's meaning: take the newest depth frame and target frame (both
u16[H*W]; target low byte = class, high byte = id, scene.rs:198), run pt_cloud + pt_cloud_weights on
the GPU, block until done (scene.rs:282) and return the overwritten Scene (scene.rs:329-330).  The
Vulkan objects the reference re-creates per call (scene.rs:152-224) live in the handle instead.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import SceneParams, check, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@dataclass
class Scene:
    height: np.ndarray        # f32[H*W]        (scene.rs:312-314)
    pos: np.ndarray           # f32[H*W, 3]     (scene.rs:316-318)
    balls: np.ndarray         # i32[100, 2]     (scene.rs:320-322)
    connections: np.ndarray   # f32[H*W, 8]     (scene.rs:324-327)
    width: int = 640

    def neighbors(self, px):
        """scene.rs:134-143, including its literal 680 typo."""
        out = []
        if px > 0:
            out.append(px - 1)
        if px < 680 * 480 - 1:
            out.append(px + 1)
        if px // 640 > 0:
            out.append(px - 640)
        if px // 640 < 480 - 1:
            out.append(px + 640)
        return out


def default_params(**kw):
    p = SceneParams()
    lib().tod_scene_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown scene parameter %r" % k)
        setattr(p, k, v)
    return p


class SceneBuilder:
    def __init__(self, device=0, **params):
        self.params = default_params(**params)
        h = C.c_void_p()
        check(lib().tod_scene_create(int(device), C.byref(self.params), C.byref(h)))
        self._h = h
        self.W, self.H = self.params.width, self.params.height

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().tod_scene_destroy(self._h)
            except Exception:  # interpreter shutdown: the module globals may already be gone
                pass
            self._h = None

    __del__ = close

    def append_batch(self, depth, target, want=("map", "world", "conn0", "conn1", "balls")):
        """depth, target: u16[n, H, W].  Returns dict of host arrays."""
        depth = np.ascontiguousarray(depth, np.uint16).reshape(-1, self.H, self.W)
        target = np.ascontiguousarray(target, np.uint16).reshape(-1, self.H, self.W)
        n = depth.shape[0]
        assert target.shape == depth.shape
        out = dict(map=np.zeros((n, self.H, self.W), np.uint32) if "map" in want else None,
                   world=np.zeros((n, self.H, self.W, 4), np.float32) if "world" in want else None,
                   conn0=np.zeros((n, self.H, self.W, 4), np.float32) if "conn0" in want else None,
                   conn1=np.zeros((n, self.H, self.W, 4), np.float32) if "conn1" in want else None,
                   balls=np.zeros((n, 100, 4), np.float32) if "balls" in want else None)
        check(lib().tod_scene_append_batch(self._h, _ptr(depth), _ptr(target), n, _ptr(out["map"]), _ptr(out["world"]),
                                           _ptr(out["conn0"]), _ptr(out["conn1"]), _ptr(out["balls"])))
        return out

    def append_batch_device(self, d_depth, d_target, n, d_map=None, d_world=None, d_conn0=None, d_conn1=None, d_balls=None, stream=None):
        check(lib().tod_scene_append_batch_device(self._h, d_depth, d_target, n, d_map, d_world, d_conn0, d_conn1, d_balls, stream))

    def materialize(self, frame=0):
        npx = self.W * self.H
        s = Scene(np.zeros(npx, np.float32), np.zeros((npx, 3), np.float32), np.zeros((100, 2), np.int32),
                  np.zeros((npx, 8), np.float32), self.W)
        check(lib().tod_scene_materialize(self._h, frame, _ptr(s.height), _ptr(s.pos), _ptr(s.balls), _ptr(s.connections)))
        return s

    def append_scene(self, depth_queue, target_queue):
        """`append_scene` (scene.rs:147): pops the newest entries (LIFO `pop`, scene.rs:186-187) and returns the Scene."""
        depth = depth_queue.pop()
        target = target_queue.pop()
        self.append_batch(depth, target, want=())
        return self.materialize(0)

    def last_kernel_ms(self):
        a, b = C.c_float(), C.c_float()
        check(lib().tod_scene_last_kernel_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value
