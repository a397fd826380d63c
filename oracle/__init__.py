"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the datasketch hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / reported baseline.
The product package ``datasketch_b200`` never imports this package and fails
loudly when its CUDA library is missing.
"""
