#!/usr/bin/env python
"""One command that closes "parity unpinned vs real TFLite" wherever a TFLite interpreter and a model exist
(neither does in the build container: SURVEY.md §8c; /root/reference/.MISSING_LARGE_BLOBS lists both blobs).

    python tools/pin_with_tflite.py [path/to/FRC_model.tflite]

With `tflite_runtime`, `ai_edge_litert` or `tensorflow` importable it runs the interpreter the reference wraps
(src/yolact.rs:18-35: FlatBufferModel::build_from_file -> InterpreterBuilder -> allocate_tensors -> invoke) on

  * the reference's two model-input fixtures (tests/golden/ref_tiles.npz = data/frc_balls.png, data/red_robot.png) and
  * four seeded tiles (tests.synth.rgb_tiles(4, seed=2)),

keeps every tensor (experimental_preserve_all_tensors) and writes tests/golden/tflite_real.npz:
    tiles u8[6,224,224,3], out{k} = output tensor k for the 6 tiles, t{idx} = every intermediate activation tensor of
    tile 0, model_sha256.  tests/test_tflite_golden.py then holds the CPU oracle (and, on a GPU box, the CUDA path) to
those bytes.  Default model: data/FRC_model.tflite (BASELINE config 1); without it the synthetic stand-in
oracle/_build/FRC_model_synth.tflite is used, which still pins the ARITHMETIC (same ten builtin operators).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_interpreter():
    for mod, attr in (("tflite_runtime.interpreter", "Interpreter"), ("ai_edge_litert.interpreter", "Interpreter"), ("tensorflow.lite", "Interpreter")):
        try:
            m = __import__(mod, fromlist=[attr])
            return getattr(m, attr), mod
        except Exception:
            continue
    return None, None


def main():
    Interpreter, where = find_interpreter()
    if Interpreter is None:
        print("pin_with_tflite: no TFLite interpreter importable (tried tflite_runtime, ai_edge_litert, tensorflow) - nothing written")
        return 2
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "data", "FRC_model.tflite")
    if not os.path.exists(path):
        from oracle import synth_model
        path, _ = synth_model.ensure_models()
        print("pin_with_tflite: real blob absent, pinning the arithmetic on the synthetic stand-in", path)
    from tests import synth
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_tiles.npz"))["tiles"]
    tiles = np.concatenate([ref, synth.rgb_tiles(4, seed=2)])
    try:
        it = Interpreter(model_path=path, num_threads=4, experimental_preserve_all_tensors=True)   # set_num_threads(4), yolact.rs:34
    except TypeError:
        it = Interpreter(model_path=path, num_threads=4)
    it.allocate_tensors()
    inp = it.get_input_details()[0]
    outs = it.get_output_details()
    rec = {"tiles": tiles, "model_sha256": np.array(hashlib.sha256(open(path, "rb").read()).hexdigest()), "interpreter": np.array(where)}
    per_out = [[] for _ in outs]
    for t in range(len(tiles)):
        it.set_tensor(inp["index"], tiles[t][None])
        it.invoke()
        for k, o in enumerate(outs):
            per_out[k].append(it.get_tensor(o["index"]).copy())
        if t == 0:
            for d in it.get_tensor_details():
                try:
                    v = it.get_tensor(d["index"])
                except ValueError:
                    continue
                if v.dtype in (np.int8, np.uint8) and v.size > 4:
                    rec["t%d" % d["index"]] = v.copy()
    for k in range(len(outs)):
        rec["out%d" % k] = np.concatenate(per_out[k])
    dst = os.path.join(ROOT, "tests", "golden", "tflite_real.npz")
    np.savez_compressed(dst, **rec)
    print("pin_with_tflite: wrote %s (%d arrays) using %s on %s" % (dst, len(rec), where, path))
    return 0


if __name__ == "__main__":
    sys.exit(main())
