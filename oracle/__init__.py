"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's detection path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``meta-viterbinet_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to this code.
"""
