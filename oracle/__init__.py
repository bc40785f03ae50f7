"""CPU oracle for the DBS-Gym hot path (``SpatialKuramoto.step`` / ``reset``).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import anything
from this package, and only as the checker or as the CPU arm that is timed
beside the GPU path -- never as part of the product path.

PARITY UNPINNED (see ``diffrax_restated.py``): the reference has no tests or
golden vectors for this path, and its integrator lives in diffrax 0.7.0, which is
neither under /root/reference nor installable here.  Goldens under
``tests/golden/`` come from the reference's own ``environment/env.py`` executed
verbatim on top of ``oracle/shims`` (``oracle/run_reference.py``).
"""
