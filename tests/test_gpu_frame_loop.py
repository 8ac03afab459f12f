"""The C++ mirror of the reference's interface (include/tod.hpp) driven like src/main.rs drives the reference:
builds tools/frame_loop.cpp against libtod_b200.so and runs a few synthetic RGB-D frames through
Yolact::classify -> target extraction -> append_scene."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_frame_loop(models, tmp_path):
    full, _ = models
    libdir = os.path.join(ROOT, "tiny-object-detection_b200", "lib")
    exe = str(tmp_path / "frame_loop")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "frame_loop.cpp"),
                           "-L" + libdir, "-ltod_b200", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe, full, "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "frame_loop: 3 frames" in out.stdout and "scene.height.size()=307200" in out.stdout
