"""CPU oracle for nerfail_b200 — TEST INFRASTRUCTURE ONLY.

A torch-CPU / numpy restatement of the reference's algorithm for the hot path, each function citing the
reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package (nerfail_b200/) never does.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against the
reference ITSELF: tests/golden/make_golden.py imports the unmodified reference functions from
/root/reference in the build container, runs them on seeded inputs and stores input/output pairs under
tests/golden/*.npz; tests/test_oracle_golden.py checks this restatement against those files.
"""
