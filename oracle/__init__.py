"""TEST INFRASTRUCTURE ONLY — CPU oracle for the CAP-VSTNet stylization hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker / CPU baseline.
"""
