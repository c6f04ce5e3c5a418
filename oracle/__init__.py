"""CPU oracle for the ManyTor step loop -- TEST INFRASTRUCTURE ONLY.

Nothing under ``manytor_b200/`` may import this package.  The only callers are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py``, and there only as the checker / the timed CPU
baseline, never as the product path.

Parity pin: the reference (victorkich/ManyTor) ships no tests or golden vectors
of its own, so the oracle is pinned against outputs of the
unmodified reference itself, run in the build container under fixed
``np.random.seed`` by ``tests/golden/make_golden.py`` and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them.
"""
from .manytor_oracle import (  # noqa: F401
    ArmSpec,
    REFERENCE_ARM,
    UR5_ARM,
    OracleEnvs,
    StepResult,
    dh,
    fk,
    fk_frames,
    r_theta,
    observations,
    sample_points_reference_stream,
    action_sample_reference_stream,
)
