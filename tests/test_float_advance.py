"""Closed-form float32 accumulation (disinfect_slam_b200/csrc/float_advance.h: k additions `p += s` as integer
arithmetic on the bit pattern while the accumulator stays in one binade) against plain repeated addition, on the CPU:
random operands in the ranges the ray caster sees, ties at every step, binade edges, zero crossings, denormals, NaN / inf,
and the three-accumulator form."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_closed_form_advance_equals_repeated_addition(tmp_path):
    exe = tmp_path / "float_advance_check"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-msse2", "-mfpmath=sse", "-o", str(exe),
                           os.path.join(HERE, "cpp", "float_advance_check.cc")])
    res = subprocess.run([str(exe), "1500000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.startswith("ok "), res.stdout[-400:] + res.stderr[-400:]
    assert int(res.stdout.split()[1]) > 3_000_000
