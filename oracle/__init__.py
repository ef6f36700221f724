"""CPU oracle for the xdiffusion sampling hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 *restatement* (functional, state-dict driven) of what the
reference computes on the hot path: noise-schedule tables, the per-timestep
sampler updates, and the four score networks.  Every function cites the
reference file:line it follows.  It is checked against the real reference
(imported from /root/reference in the authoring container) by
``tests/golden/make_golden.py`` -> committed fixtures under ``tests/golden/``
and by ``tests/test_oracle_vs_reference.py``; parity is therefore *pinned by
outputs of the reference itself* (the reference ships no tests or golden
vectors of its own).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``xdiffusion_b200``) never does: it fails loudly when its CUDA library is
missing instead of falling back to this code.
"""
