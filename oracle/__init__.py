"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the AO-v0 env-step path.

Nothing in the product package may import from here.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker / timed CPU baseline.

PARITY UNPINNED: the reference ships no tests or golden vectors and its
arithmetic lives in hcipy==0.5.1 / scikit-image==0.22.0, neither of which is
present or installable here.  This oracle restates their published algorithms
(see ao_oracle.py, each function cites the reference call site) and is pinned
only by analytic known answers (tests/test_oracle_known_answers.py).
"""
