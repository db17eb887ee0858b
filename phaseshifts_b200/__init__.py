"""phaseshifts_b200 -- B200-native (sm_100a) replacement for the phase-shift hot path of
Liam-Deacon/phaseshifts: libphsh.phsh_rel / phsh_cav / phsh_wil + the conphas unwrapping.

Public surface:
  * ``phaseshifts_b200.api``      batched tensor API (torch device tensors or host arrays)
  * ``phaseshifts_b200.libphsh``  stand-in for ``phaseshifts.lib.libphsh`` (four-file functions)
  * ``phaseshifts_b200.backend``  ``B200Backend`` + ``install()`` for ``phaseshifts.backends``
  * ``python -m phaseshifts_b200 ...``  the reference CLI with ``--backend b200`` available
"""
__version__ = "0.1.0"

from ._lib import PhshB200Error  # noqa: F401
