"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's Yacht-Auction rules and MCTS.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU arm that is timed beside
the GPU path.  The product package (``nypc_yacht_auction_b200``) never imports it and
fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference``) and
committed under ``tests/golden/``.
"""
