"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference (yanx27/PointASNL) hot path, used only as the checker / timed CPU
baseline by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` and
``--impl reference`` legs).  ``pointasnl_b200`` never imports this package.

    cref        numpy bindings of cpu_ref.c            (FPS, gather, ball query, group, kNN, three_nn, ...)
    refnative   bindings of the REFERENCE's own code compiled untouched into oracle/_ref/ by build_ref.sh
    torch_ref   line-by-line PyTorch-CPU restatement of the TF graph code (AdaptiveSampling, SampleWeights,
                PointNonLocalCell, SA layer, decoding layers, the two models)

Parity status (details in cpu_ref.c's header and DESIGN.md): the reference has no golden vectors; the
restatements are pinned against the compiled reference (``oracle/_ref``) -- CPU pieces here, GPU pieces on
the B200 box -- and the resulting vectors are committed under ``tests/golden/``.  The TF graph code
(AS / NL cells, conv/BN glue) cannot be executed (TensorFlow is not installable in this image):
**parity unpinned** for ``torch_ref``.
"""
