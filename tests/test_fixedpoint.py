"""The product's fixed-point helpers (csrc/fixedpoint.cuh: literal and fast forms, QuantizeMultiplier) against the
oracle's literal restatement of the gemmlowp rules, compiled for the host: edge cases + 12 million random draws."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixedpoint_forms_agree(tmp_path):
    exe = str(tmp_path / "fpcheck")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "fixedpoint_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout


def test_known_answers():
    import oracle
    L = oracle.lib()
    # SaturatingRoundingDoublingHighMul: (a*b*2 + 2^31) >> 32 with the single saturating case
    assert L.tod_oracle_srdhm(1 << 30, 1 << 30) == 1 << 29
    assert L.tod_oracle_srdhm(-(1 << 31), -(1 << 31)) == (1 << 31) - 1
    assert L.tod_oracle_srdhm(3, 1 << 30) == 2          # 1.5 rounds away from zero... on the positive side
    assert L.tod_oracle_srdhm(-3, 1 << 30) == -1        # -1.5 + nudge truncates toward zero
    # RoundingDivideByPOT: ties away from zero
    assert [L.tod_oracle_rdivpot(v, 1) for v in (3, -3, 1, -1, 2, -2)] == [2, -2, 1, -1, 1, -1]
    assert L.tod_oracle_rdivpot(5, 0) == 5 and L.tod_oracle_rdivpot(-5, 0) == -5
    # MultiplyByQuantizedMultiplier(x, 0.5 * 2^31, 0) = round(x / 2)
    assert [L.tod_oracle_mbqm(v, 1 << 30, 0) for v in (10, 11, -11)] == [5, 6, -5]
