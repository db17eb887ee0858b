"""``python -m phaseshifts_b200 [phsh options] --backend b200`` -- the reference CLI with the b200 backend registered.

The reference has no entry-point discovery for backends (``phaseshifts/plugins/__init__.py`` is empty), so the
registration has to happen in-process before ``phaseshifts.phsh.main()`` parses ``--backend`` (``phsh.py:811``).
"""
import sys


def main(argv=None):
    from .backend import install
    install()
    import phaseshifts.phsh as phsh
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    if "--backend" not in " ".join(sys.argv[1:]):
        sys.argv += ["--backend", "b200"]
    rc = phsh.main()
    return 0 if rc is None else rc


if __name__ == "__main__":
    sys.exit(main())
