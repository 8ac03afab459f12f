"""The C-ABI library loads without a GPU and exports every symbol include/tod.h declares; host-side logic
(model parsing, defaults, error reporting) works; device entry points fail loudly when there is no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "tod.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tod_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(tod):
    lib = tod.lib()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libtod_b200.so does not export %s" % n
    assert sorted(tod.SYMBOLS) == names, "the Python binding's symbol list and include/tod.h disagree"


def test_abi_version_and_defaults(tod):
    lib = tod.lib()
    assert lib.tod_abi_version() == 1
    p = tod.default_params()
    # pt_cloud.comp:23-39
    assert (p.width, p.height, p.bot_norm_const, p.terrain_norm_const) == (640, 480, 20, 10)
    assert abs(p.max_depth_in - 4000.0) < 1e-6 and abs(p.bump_err - 0.1) < 1e-7 and abs(p.bot_avoidance_const - 100.0) < 1e-6
    assert abs(p.x_fov - 1.51843644924) < 1e-6 and abs(p.y_fov - 1.01229096616) < 1e-6
    from tod_b200._lib import YolactOptions
    o = YolactOptions()
    lib.tod_yolact_default_options(C.byref(o))
    assert (o.top_k, o.max_dets, o.id_mode) == (200, 100, 0)
    assert abs(o.conf_thresh - 0.05) < 1e-7 and abs(o.nms_thresh - 0.5) < 1e-7
    # the struct's last fields: a layout drift between include/tod.h and the ctypes mirror would land elsewhere
    assert (o.use_cuda_graph, o.conv_impl, o.fusion, o.use_pdl, o.batches_in_flight) == (1, 0, 1, 1, 1)
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "tod.h")).read()
    body = re.search(r"typedef struct tod_yolact_options \{(.*?)\} tod_yolact_options;", hdr, re.S).group(1)
    fields = re.findall(r"^\s*(?:int32_t|float)\s+(\w+);", body, re.M)
    assert fields == [f[0] for f in YolactOptions._fields_]


def test_model_inspect_matches_reference_op_log(tod, models):
    full, small = models
    info = tod.model_inspect(full)
    # /root/reference/data/FRC_model_edgetpu.log:7-19 -> 141 operators; SURVEY §8d -> 5.62 GMAC per tile
    assert info["num_ops"] == 141
    assert info["macs"] == 5618874112
    assert tod.model_inspect(small)["num_ops"] == 141


def test_model_errors_are_reported_not_fatal(tod, models, tmp_path):
    _, small = models
    blob = open(small, "rb").read()
    with pytest.raises(tod.TodError) as e:
        tod.model_inspect(str(tmp_path / "nope.tflite"))
    assert e.value.code == -2
    bad = tmp_path / "magic.tflite"
    bad.write_bytes(blob[:4] + b"XXXX" + blob[8:])
    with pytest.raises(tod.TodError) as e:
        tod.model_inspect(str(bad))
    assert e.value.code == -3
    for cut in (100, 3000, len(blob) // 2):
        t = tmp_path / ("cut%d.tflite" % cut)
        t.write_bytes(blob[:cut])
        with pytest.raises(tod.TodError):
            tod.model_inspect(str(t))
    tiny = tmp_path / "tiny.tflite"
    tiny.write_bytes(b"1234")
    with pytest.raises(tod.TodError):
        tod.model_inspect(str(tiny))


def test_edgetpu_custom_op_is_rejected(tod, tmp_path):
    """yolact.rs:19 loads FRC_model_edgetpu.tflite (one edgetpu-custom-op); this build must refuse it with a message."""
    from oracle.fbwriter import Builder
    b = Builder()
    code = b.table([(0, "b", 32), (1, "o", b.string("edgetpu-custom-op")), (2, "i", 1), (3, "i", 32)])
    sg = b.table([(0, "o", b.vector_of_offsets([])), (1, "o", b.vector_of("i", [])), (2, "o", b.vector_of("i", [])),
                  (3, "o", b.vector_of_offsets([])), (4, "o", b.string("main"))])
    root = b.table([(0, "I", 3), (1, "o", b.vector_of_offsets([code])), (2, "o", b.vector_of_offsets([sg])),
                    (3, "o", b.string("x")), (4, "o", b.vector_of_offsets([b.table([])]))])
    path = tmp_path / "edgetpu.tflite"
    path.write_bytes(b.finish(root))
    with pytest.raises(tod.TodError) as e:
        tod.model_inspect(str(path))
    assert e.value.code == -3 and "edgetpu-custom-op" in str(e.value)


def test_no_cpu_fallback(tod, models):
    """Without a GPU every create call fails with NO_DEVICE: nothing computes on the CPU."""
    if tod.device_count() > 0:
        pytest.skip("a GPU is visible")
    _, small = models
    with pytest.raises(tod.TodError) as e:
        tod.Yolact.init(small)
    assert e.value.code == -5
    with pytest.raises(tod.TodError) as e:
        tod.SceneBuilder()
    assert e.value.code == -5


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tiny-object-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith(("//", "#", "*", '"""')))
                assert not re.search(r"^\s*(import|from)\s+oracle", code, flags=re.M), f
                assert "liboracle" not in code and "#include \"../../oracle" not in code, f
