"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the WarpSTR caller hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.
The product package ``warpstr_b200`` never imports this package.
"""
