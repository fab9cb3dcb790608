"""CPU oracle for the no-blank CTC hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``ctc_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / reported CPU baseline.

Parity status: PINNED.  ``oracle.restatement`` is checked against outputs of the
reference itself (``/root/reference/NoBlankCTC.py``, ``NoBlankBinaryCTC.py`` imported
unmodified in the build container and run in float64) that are committed as fixtures
under ``tests/golden/`` together with the generating script
``tests/golden/make_golden.py``.
"""
from .restatement import (  # noqa: F401
    nbctc_alpha_beta,
    nbctc_loss_grad,
    nbbctc_emissions,
    nbbctc_loss_grad,
    best_path,
    frame_argmax,
)
