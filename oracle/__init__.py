"""TEST INFRASTRUCTURE ONLY — CPU oracle for the TCDiff denoising hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker (never as the thing measured as the
product or shipped).  The product path (``tcdiff_b200``) never imports this package
and fails loudly when its CUDA library is missing.

Contents
--------
* ``p3d.py``            restatement of the five ``pytorch3d.transforms`` (pin 0.7.1,
                        /root/reference/requirements.txt:70) functions the path uses.
                        pytorch3d is un-vendored and absent offline and the reference has
                        no tests for it => **parity unpinned** at that boundary (the
                        restatement follows the published 0.7.x algorithm from memory).
* ``ref_shim.py``       imports the UNMODIFIED reference from /root/reference (exists in
                        the build container only, never on the GPU box) with in-memory
                        stand-ins for its missing render-only dependencies.
* ``tcdiff_oracle.py``  functional CPU (torch fp32) restatement of DanceDecoder /
                        GaussianDiffusion / SMPLSkeleton, each function citing the
                        reference file:line it follows.  Pinned against the real
                        reference by ``oracle/make_golden.py`` -> ``tests/golden/*.pt``.
* ``synth.py``          deterministic synthetic weights / inputs / noise banks.
"""
