"""Out-of-bounds-write check of every kernel's scratch, staging and result buffers without compute-sanitizer (closed on
this pool): `scripts/guard_run.py` drives every host / file entry point at small odd sizes with PHSH_B200_GUARD=1, which
brackets each device buffer of the library with 256-byte canary bands and verifies them on release.  The detector proves
itself first (two deliberate one-byte overruns must be reported)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_kernel_writes_outside_its_buffers():
    env = dict(os.environ, PHSH_B200_GUARD="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "guard_run.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("guard: 0 violations in ") and int(last.split()[4]) > 50
