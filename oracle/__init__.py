"""Test infrastructure only (CPU restatement of the reference algorithm).

Importable from ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` -- never from ``pde-net_b200/``.
"""
