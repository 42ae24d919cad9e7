"""
CPU check of the table-based correctly rounded sin/cos of the near-field walk
(tesseroid_variable_density_b200/csrc/tvd_crtrig.cuh), and of the correctly rounded exp of K1/K4:
their host/device cores are compiled by g++
and compared with quad precision (libquadmath) rounded to double.  The reference gets these
values from glibc (_tesseroid.pyx:113-121,168-169,242,401), which was correctly rounded at the
time of the paper.
"""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_cr_sincos_core_is_correctly_rounded(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "crtest")
    cmd = ["g++", "-O2", "-ffp-contract=off", os.path.join(HERE, "csrc", "crtrig_host_test.cpp"),
           "-o", exe, "-lquadmath", "-lm"]
    flags = open("/proc/cpuinfo").read()
    if " fma " in flags:
        cmd.insert(3, "-mfma")          # hardware FMA; without it libm's exact software fma is used
    subprocess.check_call(cmd)
    out = subprocess.run([exe, "400000"], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("range")]
    assert len(lines) == 6
    for l in lines:
        m = re.search(r"not-ok (\d+) .*sin != quad (\d+) \(with ok: (\d+)\)\s+cos != quad (\d+) "
                      r"\(with ok: (\d+)\)", l)
        notok, bad_s, bad_ok_s, bad_c, bad_ok_c = (int(g) for g in m.groups())
        assert bad_ok_s == 0 and bad_ok_c == 0, l      # accepted by Ziv's test => correctly rounded
        assert notok <= 400000 * 2e-4, l               # the slow path is taken by ~2e-5 of the calls
    elines = [l for l in out.splitlines() if l.startswith("exp range")]
    assert len(elines) == 4
    for l in elines:
        m = re.search(r"not-ok (\d+) .*exp != quad (\d+) \(with ok: (\d+)\)", l)
        notok, bad, bad_ok = (int(g) for g in m.groups())
        assert bad_ok == 0, l
        assert notok <= 400000 * 5e-4, l
