"""CPU oracle for the adversarial-perturbation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``imagecompression_adversarial_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker / CPU baseline.

What it restates (plain ``torch`` fp32/fp64, eager, runs on CPU):

* the loop of ``/root/reference/attack_rd.py:332-379`` (``attack_our``) and ``:381-575``
  (``attack_``), the final metrics of ``/root/reference/self_ensemble.py:173-252`` (``eval``),
  the clamp-with-gradient ops of ``/root/reference/utils/ops.py:28-56``, the sign update of
  ``/root/reference/attack_ifgsm.py:348-438`` and the RD loss of ``/root/reference/train.py:37-96``;
* the arithmetic of the reference's un-vendored dependencies: ``compressai`` (unpinned, API era
  1.1.x-1.2.x: GDN, EntropyBottleneck, GaussianConditional, MaskedConv2d, the four zoo model
  families) and ``pytorch_msssim`` (unpinned).  Neither package is in ``/root/reference`` nor
  installed, so their published algorithms are restated and anchored on the
  reference's own call sites and in-repo partial restatements (``utils/ops.py:58-97`` GDN,
  ``anchors/model.py:86-108`` forwards, ``visual_distribution.py:85-100`` discretised Gaussian,
  ``utils/torch_msssim.py`` MS-SSIM variant 2) and on CompressAI's published parameter counts.

PARITY PINNING.  The reference has no tests and no golden vectors.  What *can* run
in the build container is pinned: ``tests/golden/make_golden.py`` imports the reference's own
``utils/ops.py`` (Low_bound/Up_bound/GDN), ``utils/torch_msssim.py``, ``anchors/utils.py`` and the
loop functions ``attack_rd.attack_``/``attack_our`` + ``self_ensemble.eval`` (with this oracle's
models standing in for the missing ``compressai`` zoo) and stores their outputs as fixtures that
``tests/test_oracle_golden.py`` replays.  The CompressAI / pytorch_msssim arithmetic itself has no
reference-side vector to check against: for those pieces parity is **unpinned** (structural
checksums only: parameter counts, closed-form identities, fp64 gradchecks).
"""
