"""tod_b200 - B200-native per-frame perception hot path of icf3ver/tiny-object-detection.

The product is lib/libtod_b200.so (hand-written CUDA for sm_100a behind the C ABI of include/tod.h).
This package is the thin host-side mirror of the reference's own interface for the path:
`Yolact.init()/classify()` (src/yolact.rs) and `append_scene()` / `Scene` (src/scene.rs).
Nothing here computes on the CPU and nothing here imports the test oracle.
"""
from ._lib import LIB_PATH, SYMBOLS, TodError, build, device_count, lib  # noqa: F401
from .scene import Scene, SceneBuilder, default_params  # noqa: F401
from . import shard  # noqa: F401
from .path import Path, modify_path  # noqa: F401
from .yolact import Pool, Yolact, YolactPool, model_inspect  # noqa: F401
