"""AddressSanitizer + UBSan over the kernels' index arithmetic: the device source compiled for the host (the emulation of tests/emu) with
-fsanitize=address,undefined, a cross-section of the parity checks run against it in a child process.  Found in round 1: a 7-double local array
receiving a 9-double speed-bias block in the solver's gradient-norm helper (harmless by luck on the GPU, undefined all the same)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(1500)
def test_kernel_logic_under_address_and_ub_sanitizers():
    from emu import build_emu
    built = build_emu.build_sanitized()
    if built is None:
        pytest.skip("no libasan in this toolchain")
    lib, asan = built
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "sanitized_checks.py"), lib], capture_output=True, text=True, env=env, timeout=1400)
    tail = (p.stdout + p.stderr)[-3000:]
    assert p.returncode == 0 and "AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr, tail
    assert "checks 18" in p.stdout, tail
