"""Drop-in evidence: the reference's OWN test files for the path's boundary (backend registry, phsh helpers,
wrappers helpers, conphas/ViPErLEED writer) still pass, unmodified, when ``phaseshifts_b200.backend.install()`` has
replaced ``phaseshifts.lib.libphsh`` and registered the b200 backend.  Runs only where the reference checkout is
available (this container); CPU only."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = os.environ.get("PHASESHIFTS_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["test_backends.py", "test_phsh_backend.py", "test_phsh_helpers.py", "test_wrappers_helpers.py",
         "test_conphas_viper.py"]

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference checkout not available")


def test_reference_boundary_tests_pass_with_b200_installed(tmp_path):
    plugin = tmp_path / "b200_plugin.py"
    plugin.write_text(textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from phaseshifts_b200.backend import install
        install()
        import phaseshifts.backends as B
        assert "b200" in B.DEFAULT_BACKENDS
    """ % (ROOT, REF)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT, REF]))
    cmd = [sys.executable, "-m", "pytest", "-p", "b200_plugin", "-q", "--no-header", "-p", "no:cacheprovider",
           "-c", os.devnull, "--rootdir", str(tmp_path)] + [os.path.join(REF, "tests", f) for f in FILES]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
