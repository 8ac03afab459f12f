"""CUDA int8 graph vs the CPU oracle's TFLite restatement: every tensor bit-exact (fusion off),
outputs bit-exact (fusion on, CUDA graph on), batch order independence.  Calls go through the C ABI."""
import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu

OPS_WITH_DATA = None


def _oracle_run(path, tile):
    m = oracle.Model(path)
    m.invoke(tile, threads=8)
    return m


@pytest.mark.parametrize("conv_impl", [1, 0, 2])
def test_small_model_every_tensor(tod, models, conv_impl):
    _, small = models
    tiles = synth.rgb_tiles(3, S=64, seed=31)
    y = tod.Yolact.init(small, max_tiles=3, fusion=0, use_cuda_graph=0, conv_impl=conv_impl)
    y.infer_tiles(tiles, detections=False)
    for t in range(3):
        m = _oracle_run(small, tiles[t])
        for op in range(m.num_ops):
            ti = m.op_output(op)
            want = m.tensor(ti)
            got = y.fetch_tensor(ti, 3)[t]
            assert got.dtype == want.dtype
            assert np.array_equal(got.reshape(-1), want.reshape(-1)), "tile %d op %d (code %d) tensor %d differs in %d of %d" % (
                t, op, m.op_code(op), ti, (got.reshape(-1) != want.reshape(-1)).sum(), want.size)


@pytest.mark.parametrize("conv_impl", [1, 0, 2])
def test_full_model_outputs(tod, models, conv_impl):
    full, _ = models
    tiles = synth.rgb_tiles(2, seed=32)
    y = tod.Yolact.init(full, max_tiles=4, conv_impl=conv_impl)  # fusion + CUDA graph (defaults)
    res = y.infer_tiles(tiles, detections=False)
    for t in range(2):
        m = _oracle_run(full, tiles[t])
        for k, ti in enumerate(m.outputs):
            want = m.tensor(ti)
            assert np.array_equal(res["outputs"][k][t].reshape(-1), want.reshape(-1)), "tile %d output %d" % (t, k)
    st = y.stats()
    assert st["macs_per_tile"] == oracle.Model(full).macs() or st["macs_per_tile"] == 5618874112


def test_fusion_and_graph_do_not_change_bytes(tod, models):
    _, small = models
    tiles = synth.rgb_tiles(5, S=64, seed=33)
    a = tod.Yolact.init(small, max_tiles=5, fusion=0, use_cuda_graph=0).infer_tiles(tiles)
    b = tod.Yolact.init(small, max_tiles=8, fusion=1, use_cuda_graph=1)
    rb = b.infer_tiles(tiles)
    rb2 = b.infer_tiles(tiles[::-1].copy())  # replayed graph, reversed batch
    for k in range(5):
        assert np.array_equal(a["outputs"][k], rb["outputs"][k])
        assert np.array_equal(a["outputs"][k], rb2["outputs"][k][::-1])
    assert np.array_equal(a["tile_classes"], rb["tile_classes"])


def test_pdl_does_not_change_bytes(tod, models):
    """Programmatic dependent launch only overlaps prologues: same bytes with it on and off, over repeated replays."""
    full, _ = models
    tiles = synth.rgb_tiles(4, seed=34)
    a = tod.Yolact.init(full, max_tiles=4, use_pdl=0).infer_tiles(tiles)
    b = tod.Yolact.init(full, max_tiles=4, use_pdl=1)
    for _ in range(3):
        rb = b.infer_tiles(tiles)
        for k in range(5):
            assert np.array_equal(a["outputs"][k], rb["outputs"][k])
        assert np.array_equal(a["tile_classes"], rb["tile_classes"])


def test_sibling_head_fusion_and_launch_sizing_do_not_change_bytes(tod, models, monkeypatch):
    """The planner merges the box head of every pyramid level into the coefficient head's launch (58 instead of 63 tcgen05
    launches); handles that share a GPU give each convolution CTA two tiles.  Neither changes a byte: all five outputs, the class
    maps and the detections of the merged / throughput-sized plan against the separate / latency-sized one, odd tile count."""
    full, _ = models
    tiles = synth.rgb_tiles(7, seed=41)
    monkeypatch.setenv("TOD_HEAD_MERGE", "0")
    ya = tod.Yolact.init(full, max_tiles=7)
    monkeypatch.delenv("TOD_HEAD_MERGE")
    yb = tod.Yolact.init(full, max_tiles=7, batches_in_flight=3)
    assert ya.stats()["tc_conv_layers"] == yb.stats()["tc_conv_layers"] + 5
    a = ya.infer_tiles(tiles, detections=True)
    for _ in range(2):
        b = yb.infer_tiles(tiles, detections=True)
        for k in range(5):
            assert np.array_equal(a["outputs"][k], b["outputs"][k]), k
        assert np.array_equal(a["tile_classes"], b["tile_classes"])
        for da, db in zip(a["dets"], b["dets"]):
            assert da["n"] == db["n"] and np.array_equal(da["prior"], db["prior"])
            assert np.array_equal(da["score"].view(np.uint32), db["score"].view(np.uint32))
            assert np.array_equal(da["box"].view(np.uint32), db["box"].view(np.uint32))


@pytest.mark.parametrize("n,max_tiles", [(1, 1), (3, 3), (5, 7), (13, 16)])
def test_tile_counts_cross_check(tod, models, n, max_tiles):
    """Odd and partial tile counts (patch groups with tail masking, CTA pairs with a dummy tile, flat GEMMs whose last
    M tile is ragged): the tensor-core path against the CUDA-core direct path, full model, every output and the class maps."""
    full, _ = models
    tiles = synth.rgb_tiles(n, seed=60 + n)
    a = tod.Yolact.init(full, max_tiles=max_tiles, conv_impl=1).infer_tiles(tiles)
    b = tod.Yolact.init(full, max_tiles=max_tiles, conv_impl=0).infer_tiles(tiles)
    for k in range(5):
        assert np.array_equal(a["outputs"][k], b["outputs"][k]), "output %d differs at n=%d" % (k, n)
    assert np.array_equal(a["tile_classes"], b["tile_classes"])


def test_errors(tod, models, tmp_path):
    _, small = models
    with pytest.raises(tod.TodError):
        tod.Yolact.init(str(tmp_path / "missing.tflite"))
    bad = tmp_path / "bad.tflite"
    bad.write_bytes(open(small, "rb").read()[:4000])
    with pytest.raises(tod.TodError):
        tod.Yolact.init(str(bad))
    y = tod.Yolact.init(small, max_tiles=2)
    with pytest.raises(tod.TodError):
        y.infer_tiles(synth.rgb_tiles(3, S=64))


def test_pool_matches_single_handle(tod, models):
    """Batches through a pool of three handles (alternate handles, one host thread each) give exactly what one handle gives."""
    full, _ = models
    batches = [synth.rgb_tiles(4, seed=300 + k) for k in range(7)]
    y = tod.Yolact.init(full, max_tiles=4)
    want = [y.infer_tiles(b, detections=True, float_masks=False) for b in batches]
    y.close()
    pool = tod.YolactPool(full, depth=3, max_tiles=4)
    got = pool.map(batches, detections=True, float_masks=False)
    pool.close()
    for g, w in zip(got, want):
        assert np.array_equal(g["tile_classes"], w["tile_classes"])
        for a, b in zip(g["outputs"], w["outputs"]):
            assert np.array_equal(a, b)
        for dg, dw in zip(g["dets"], w["dets"]):
            assert dg["n"] == dw["n"]
            assert np.array_equal(dg["prior"], dw["prior"]) and np.array_equal(dg["cls"], dw["cls"])
            assert np.array_equal(dg["score"].view(np.uint32), dw["score"].view(np.uint32))
            assert np.array_equal(dg["masks_bits"], dw["masks_bits"])


def test_trace_steps(tod, models):
    """The diagnostics entry point: one entry per planned step plus the three post-processing stages, end times
    non-negative and bounded by the step end, lanes inside the executor's lane range."""
    _, small = models
    tiles = synth.rgb_tiles(2, S=64, seed=5)
    y = tod.Yolact.init(small, max_tiles=2)
    y.infer_tiles(tiles, detections=False)
    iso, kinds = y.profile_ops(2)
    end, lanes, kinds2 = y.trace_steps(2)
    assert len(end) == len(iso) + 3
    assert np.array_equal(kinds, kinds2[:len(iso)])
    assert (end[:len(iso)] > 0).all() and (lanes[:len(iso)] >= 0).all() and (lanes[:len(iso)] < 16).all()
    assert end[len(iso)] > 0  # segmentation pass ran
    # the class maps of a later normal call are unaffected by the traced (graph-less) run
    a = y.infer_tiles(tiles, detections=False)["tile_classes"]
    y2 = tod.Yolact.init(small, max_tiles=2)
    b = y2.infer_tiles(tiles, detections=False)["tile_classes"]
    assert np.array_equal(a, b)
