"""CPU oracle for the cluster-tracking hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing under
``pcseqlearning_b200/`` does.  See oracle/README.md for how the oracle is pinned.
"""
