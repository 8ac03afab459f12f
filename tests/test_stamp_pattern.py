"""CPU checks of the generated terrain stamp (csrc/stamp_pattern.inc) that scene.cu::stamp_pruned_kernel unrolls."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "tiny-object-detection_b200", "csrc", "stamp_pattern.inc")


def test_committed_pattern_is_the_generators_output():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_stamp_pattern
    assert open(INC).read() == gen_stamp_pattern.emit()


def test_pattern_equals_direct_definition(tmp_path):
    """host emulation of PRMT / vmaxu2: the unrolled sequence == max over dx*dx + dy*dy <= 80, both row parities"""
    exe = str(tmp_path / "spc")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "stamp_pattern_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "stamp pattern: ok" in out.stdout
