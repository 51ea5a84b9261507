"""TEST INFRASTRUCTURE ONLY — CPU oracle for the VectorBase top-k hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker / the timed CPU comparator.
The product path (``typeagent-py_b200``) never imports this package and fails
loudly when its CUDA library is missing.
"""
