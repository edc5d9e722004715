"""csrc/glibc_f32.h — the fdlibm atan2f / acosf restatement the 3DSC kernel bins with — against the libm
of this machine, bit for bit (no GPU: the header is plain C++ on the host side).

glibc <= 2.40 (this image: 2.39; the reference's Ubuntu 14.04: 2.19) computes both with the Sun fdlibm
float routines; a host with glibc >= 2.41 has the correctly rounded CORE-MATH versions instead and must
use fe_set_angle_libm(ctx, FE_LIBM_CORRECTLY_ROUNDED) — there this test is skipped."""
import os
import platform
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _glibc():
    name, ver = platform.libc_ver()
    try:
        return name, tuple(int(x) for x in ver.split(".")[:2])
    except ValueError:
        return name, (0, 0)


def test_restated_fdlibm_equals_host_libm(tmp_path):
    name, ver = _glibc()
    if name != "glibc" or ver >= (2, 41):
        pytest.skip("host libm is not the fdlibm-based glibc (<= 2.40): %s %s" % (name, ver))
    exe = tmp_path / "check_glibc_f32"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread",
                           os.path.join(ROOT, "tests", "native", "check_glibc_f32.cpp"), "-o", str(exe)])
    # every 3rd float for acosf / atanf (1.4e9 values each) and 2e8 atan2f pairs: ~10 s on 8 threads;
    # the exhaustive run (stride 1, 1e9 pairs) is what profiles/r2_libm_parity.md records
    out = subprocess.run([str(exe), "3", "200000000"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "acosf" in out.stdout and " 0 differ" in out.stdout
