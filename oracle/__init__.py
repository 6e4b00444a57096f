"""ORACLE — test infrastructure only.

CPU restatement of the reference's algorithm for the LECO training-step hot path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker.
Nothing under ``leco_b200/`` imports it.
"""
