"""ctypes binding of lib/libtod_b200.so (the C ABI of include/tod.h).

This is plumbing only: every computation happens in the CUDA library.  There is no CPU path - if the
library is missing, or no B200 is visible, the calls fail loudly.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOD_B200_LIB") or os.path.join(_HERE, "lib", "libtod_b200.so")  # override: A/B runs of two builds

TOD_OK = 0
TOD_WARN_REFERENCE_DIVERGES = 1
ERRORS = {-1: "INVALID_ARG", -2: "IO", -3: "MODEL", -4: "CUDA", -5: "NO_DEVICE", -6: "CAPACITY", -7: "UNSUPPORTED"}

# every symbol include/tod.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "tod_last_error", "tod_abi_version", "tod_device_count",
    "tod_scene_default_params", "tod_scene_create", "tod_scene_destroy", "tod_scene_append_batch",
    "tod_scene_append_batch_device", "tod_scene_materialize", "tod_scene_last_kernel_ms",
    "tod_yolact_default_options", "tod_yolact_create", "tod_yolact_destroy", "tod_yolact_classify",
    "tod_yolact_classify_batch", "tod_yolact_classify_batch_device", "tod_model_inspect", "tod_yolact_set_priors",
    "tod_yolact_last_diverged", "tod_yolact_num_outputs", "tod_yolact_output_info", "tod_yolact_num_tensors",
    "tod_yolact_num_ops", "tod_yolact_tensor_info", "tod_yolact_infer_tiles", "tod_yolact_infer_tiles_cells", "tod_yolact_infer_tiles_device",
    "tod_yolact_fetch_output", "tod_yolact_fetch_tensor", "tod_yolact_fetch_tile_classes", "tod_yolact_fetch_detections",
    "tod_yolact_stats", "tod_yolact_profile_ops", "tod_i8_gemm_selftest", "tod_conv_selftest",
    "tod_conv_selftest_ex", "tod_i8_mma_peak", "tod_yolact_step_macs", "tod_yolact_trace_steps", "tod_yolact_fetch_output_f32",
    "tod_yolact_input_info", "tod_pool_create", "tod_pool_destroy", "tod_pool_num_devices", "tod_pool_num_handles",
    "tod_pool_infer_tiles", "tod_pool_classify_batch", "tod_pool_rgbd_batch",
    "tod_path_reference_panics", "tod_path_modify", "tod_path_serialize",
]


class TodError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tod_b200: %s (%d): %s" % (ERRORS.get(code, "?"), code, msg))
        self.code = code


class SceneParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_depth_in", C.c_float), ("x_fov", C.c_float),
                ("y_fov", C.c_float), ("bot_avoidance_const", C.c_float), ("bot_norm_const", C.c_int32),
                ("terrain_norm_const", C.c_int32), ("bump_err", C.c_float), ("sample_shift", C.c_int32),
                ("weights_mode", C.c_int32), ("max_batch", C.c_int32)]


class YolactOptions(C.Structure):
    _fields_ = [("max_tiles", C.c_int32), ("id_mode", C.c_int32), ("conf_thresh", C.c_float), ("nms_thresh", C.c_float),
                ("top_k", C.c_int32), ("max_dets", C.c_int32), ("use_cuda_graph", C.c_int32), ("conv_impl", C.c_int32),
                ("fusion", C.c_int32), ("use_pdl", C.c_int32), ("batches_in_flight", C.c_int32)]


class Detections(C.Structure):
    _fields_ = [("max_dets", C.c_int32), ("count", C.c_void_p), ("boxes", C.c_void_p), ("scores", C.c_void_p),
                ("classes", C.c_void_p), ("priors", C.c_void_p), ("masks", C.c_void_p), ("masks_bin", C.c_void_p), ("masks_bits", C.c_void_p),
                ("masks_tile_bits", C.c_void_p)]


def build(force=False):
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    cmd = ["make", "-C", _HERE, "-s", "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TodError(-5, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.tod_last_error.restype = C.c_char_p
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.tod_scene_create.argtypes = [C.c_int, vp, vp]
        L.tod_scene_destroy.argtypes = [vp]
        L.tod_scene_destroy.restype = None
        L.tod_scene_append_batch.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]
        L.tod_scene_append_batch_device.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
        L.tod_scene_materialize.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.tod_scene_last_kernel_ms.argtypes = [vp, vp, vp]
        L.tod_yolact_create.argtypes = [C.c_char_p, C.c_int, vp, vp]
        L.tod_yolact_destroy.argtypes = [vp]
        L.tod_yolact_destroy.restype = None
        L.tod_yolact_classify.argtypes = [vp, vp, C.c_int, C.c_int]
        L.tod_yolact_classify_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.tod_yolact_classify_batch_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.tod_model_inspect.argtypes = [C.c_char_p, vp, vp, vp]
        L.tod_yolact_set_priors.argtypes = [vp, vp, C.c_int]
        L.tod_yolact_last_diverged.argtypes = [vp, vp]
        for f in ("num_outputs", "num_tensors", "num_ops"):
            getattr(L, "tod_yolact_" + f).argtypes = [vp]
        L.tod_yolact_output_info.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        L.tod_yolact_tensor_info.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.tod_yolact_infer_tiles.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        L.tod_yolact_infer_tiles_cells.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
        L.tod_yolact_infer_tiles_device.argtypes = [vp, vp, C.c_int, vp]
        L.tod_yolact_fetch_output.argtypes = [vp, C.c_int, C.c_int, vp]
        L.tod_yolact_fetch_tensor.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
        L.tod_yolact_fetch_tile_classes.argtypes = [vp, C.c_int, vp]
        L.tod_yolact_fetch_detections.argtypes = [vp, C.c_int, vp]
        L.tod_yolact_stats.argtypes = [vp, vp, vp, vp]
        L.tod_yolact_profile_ops.argtypes = [vp, C.c_int, vp, vp, C.c_int]
        L.tod_i8_gemm_selftest.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.tod_conv_selftest.argtypes = [C.c_int] * 8 + [vp, vp, vp]
        L.tod_conv_selftest_ex.argtypes = [C.c_int] * 9 + [vp, vp, vp]
        L.tod_i8_mma_peak.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.tod_yolact_step_macs.argtypes = [vp, vp, C.c_int]
        L.tod_yolact_trace_steps.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int]
        L.tod_yolact_fetch_output_f32.argtypes = [vp, C.c_int, C.c_int, vp]
        L.tod_yolact_input_info.argtypes = [vp, vp]
        L.tod_pool_create.argtypes = [C.c_char_p, vp, C.c_int, C.c_int, vp, vp]
        L.tod_pool_destroy.argtypes = [vp]
        L.tod_pool_destroy.restype = None
        L.tod_pool_num_devices.argtypes = [vp]
        L.tod_pool_num_handles.argtypes = [vp]
        L.tod_pool_infer_tiles.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
        L.tod_pool_classify_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.tod_pool_rgbd_batch.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]
        L.tod_path_modify.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
        L.tod_path_serialize.argtypes = [C.c_uint64, vp, C.c_int, vp, C.c_size_t, vp]
        _lib = L
    return _lib


def check(rc):
    """0 / positive warning codes pass through; negative codes raise (the Rust shim `.expect()`s them)."""
    if rc < 0:
        raise TodError(rc, lib().tod_last_error().decode(errors="replace"))
    return rc


def conv_selftest(tiles, H, W, IC, OC, K, iters=3, device=0, flags=0):
    """(ms_tcgen05, ms_direct, mismatching bytes) of one convolution run through both kernels."""
    a, b, bad = C.c_float(), C.c_float(), C.c_longlong(-1)
    check(lib().tod_conv_selftest_ex(device, tiles, H, W, IC, OC, K, iters, flags, C.byref(a), C.byref(b), C.byref(bad)))
    return a.value, b.value, bad.value


def device_count():
    n = C.c_int(0)
    rc = lib().tod_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def i8_mma_peak(n_mma=20000, iters=5, device=0):
    """Measured tensor-pipe peak of tcgen05.mma.kind::i8 in TOP/s (no loads, no epilogue)."""
    v = C.c_double(0.0)
    check(lib().tod_i8_mma_peak(device, n_mma, iters, C.byref(v)))
    return v.value
