"""tcgen05 int8 implicit-GEMM convolution vs the CUDA-core direct kernel on the same device buffers
(the direct kernel itself is pinned against the CPU oracle in test_gpu_graph.py).  Bit-exact bar."""
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [
    # tiles, H, W, IC, OC, K
    (2, 28, 28, 256, 256, 3),    # FPN / protonet / head tower
    (3, 56, 56, 256, 256, 3),    # proto_conv3 (odd tile count: pn = 2 tail masking)
    (2, 28, 28, 256, 243, 3),    # class head: OC padded to 256, byte stores
    (2, 14, 14, 256, 12, 3),     # box head
    (5, 7, 7, 256, 96, 3),       # coefficient head on P5
    (9, 4, 4, 256, 256, 3),      # P6
    (33, 2, 2, 256, 256, 3),     # P7
    (2, 56, 56, 256, 32, 1),     # proto_out
    (2, 28, 28, 192, 32, 1),     # MobileNetV2 project (BK = 64)
    (1, 112, 112, 16, 96, 1),    # block_1 expand: IC = 16 -> K chunk zero-filled to 32
    (2, 56, 56, 144, 32, 1),     # IC = 144 -> 5 chunks of 32, last half zero-filled
    (2, 14, 14, 96, 576, 1),     # OC = 576 -> 3 N tiles of 192
    (2, 7, 7, 960, 160, 1),      # BK = 64, 15 chunks
    (2, 28, 28, 32, 81, 1),      # seg head style: OC = 81
    (1, 1, 1000, 256, 256, 1),   # plain GEMM, M not a multiple of 128
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_matches_direct(tod, shape):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(*shape, iters=2)
    assert bad == 0, "%d bytes differ (tcgen05 %.3f ms, direct %.3f ms)" % (bad, ms_tc, ms_direct)


# epilogue variants: (flags, meaning) — see tod_conv_selftest_ex in include/tod.h
VARIANTS = [(1, "relu6-clamp"), (2, "byte-map"), (3, "clamp+byte-map"), (16, "act-min"), (4, "general-epilogue"), (8, "shift0-fallback")]
VARIANT_SHAPES = [(3, 28, 28, 256, 256, 3), (2, 14, 14, 256, 243, 3), (5, 7, 7, 256, 96, 3), (2, 56, 56, 144, 32, 1),
                  (1, 112, 112, 16, 96, 1), (2, 7, 7, 960, 160, 1), (2, 14, 14, 96, 576, 1), (3, 28, 28, 32, 16, 1)]


@pytest.mark.parametrize("flags", [v[0] for v in VARIANTS], ids=[v[1] for v in VARIANTS])
@pytest.mark.parametrize("shape", VARIANT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_epilogue_variants(tod, shape, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(*shape, iters=1, flags=flags)
    assert bad == 0, "%d bytes differ with epilogue flags %d" % (bad, flags)


# stride-2 3x3 (FPN P6 / P7 down-samplers): the A box walks the input with TMA element strides
@pytest.mark.parametrize("shape", [(9, 7, 7, 256, 256, 3), (33, 4, 4, 256, 256, 3), (2, 15, 15, 64, 32, 3), (3, 28, 28, 32, 48, 3)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("flags", [32, 32 | 1, 32 | 4], ids=["s2", "s2-clamp", "s2-general"])
def test_conv_tc_stride2(tod, shape, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(*shape, iters=1, flags=flags)
    assert bad == 0, "%d bytes differ (stride 2, flags %d)" % (bad, flags)


# CTA-pair kernel (cta_group::2): shapes with >= 64 M tiles and >= 8 K iterations; odd tile counts leave the odd CTA a dummy tile
PAIR_SHAPES = [(16, 28, 28, 256, 256, 3), (9, 56, 56, 256, 256, 3), (33, 14, 14, 256, 256, 3), (1, 1, 20000, 1024, 256, 1),
               (17, 28, 28, 128, 96, 3), (40, 14, 14, 256, 64, 3)]


@pytest.mark.parametrize("flags", [0, 1], ids=["sat", "clamp"])
@pytest.mark.parametrize("shape", PAIR_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_pair(tod, shape, flags):
    from tod_b200 import _lib
    ms_tc, ms_direct, bad = _lib.conv_selftest(*shape, iters=1, flags=flags)
    assert bad == 0, "%d bytes differ (CTA-pair path, flags %d)" % (bad, flags)
