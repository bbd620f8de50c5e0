"""The kernels divide by compile-time constants with a 3-operation FMA sequence (csrc/f16_device.cuh: DC / operator/).
oracle/divc_check.c enumerates floats and compares that sequence with the IEEE quotient bit for bit; the full
enumeration (stride 1: 5.4e10 quotients, every float with 2^-100 <= |x| < 2^100 for 16 divisors, 0 mismatches, ~70 s on
8 cores) is recorded in profiles/r01_divc_exhaustive.txt -- here a strided subset keeps the CPU suite fast."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constant_division_is_correctly_rounded():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_build", "divc_check"), "127"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    m = re.search(r"TOTAL checked (\d+) mismatches (\d+)", r.stdout)
    assert m and int(m.group(1)) > 4e8 and int(m.group(2)) == 0, r.stdout


def test_every_kernel_divisor_is_enumerated():
    """Every DC(...) divisor in the device code is in the checker's list."""
    src = ""
    csrc = os.path.join(ROOT, "neuralplane_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        src += open(os.path.join(csrc, f)).read()
    used = set(re.findall(r"/ DC\(([^)]+)\)", src))
    names = {"kPi": "3.14159265358979323846f", "UAV_M": "300.0f"}
    chk = open(os.path.join(ROOT, "oracle", "divc_check.c")).read()
    listed = re.search(r"kDivisors\[\] = \{([^}]*)\}", chk).group(1)
    for d in used:
        assert names.get(d, d) in listed, d


def test_fast_fmod_twopi_equals_fmodf():
    """wrap_PI's `%` (envs/utils/utils.py:144-154) runs as a 10-instruction exact remainder in the kernels (fmod_twopi,
    csrc/f16_device.cuh); oracle/fmod_check.c compares it with fmodf bit for bit.  Stride 1 (2.5e9 floats, every float with
    |a| < 6e6, 0 mismatches) is recorded in profiles/r02_fmod_exhaustive.txt; a strided subset runs here."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_build", "fmod_check"), "61"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    m = re.search(r"(\d+) floats .* (\d+) mismatches", r.stdout)
    assert m and int(m.group(1)) > 4e7 and int(m.group(2)) == 0, r.stdout
    dev = open(os.path.join(ROOT, "neuralplane_b200", "csrc", "f16_device.cuh")).read()
    chk = open(os.path.join(ROOT, "oracle", "fmod_check.c")).read()
    body = lambda src: re.sub(r"\s+", " ", re.search(r"float q = truncf.*?return r == 0\.0f \? copysignf\(0\.0f, a\) : r;\s*\}", src, re.S).group(0))
    strip = lambda t: re.sub(r"//[^\n]*", "", t)
    assert body(strip(dev)) == body(strip(chk))          # the checker enumerates the very statements the kernel runs
