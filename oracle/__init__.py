"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is the CPU checker for the B200 hot path: restatements of the
reference's algorithms (each function cites the reference file:line it follows), the import
shim that runs the unmodified reference in the build container, synthetic input generators
and the script that produced ``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  The product (``biscuit_b200/``) never does: it fails loudly if
its CUDA library is missing instead of falling back to anything here.
"""
