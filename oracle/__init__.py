"""CPU oracle for the vision_mtl per-step multi-task hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vision_mtl_b200/`` imports this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` (``cpu_baseline`` / ``--impl reference``) may import it, and
there it is the checker (or the timed CPU baseline), never the product.

Parity status
-------------
* ``CrossStitchLayer``, ``MTANMiniUnet`` (gates + 1x1 heads), ``SILogLoss`` and
  ``nn.CrossEntropyLoss``: PINNED.  ``oracle/make_golden.py`` imports the
  unmodified reference modules from ``/root/reference`` (behind the import
  shims in ``oracle/ref_shims.py``), runs them on seeded inputs and commits the
  input/output vectors under ``tests/golden/``; ``tests/test_oracle_golden.py``
  checks this restatement against those vectors.
* torchmetrics==0.7.3 arithmetic (Accuracy / JaccardIndex / FBetaScore / MAE,
  reference call sites vision_mtl/lit_module.py:48-69,106-118): PARITY UNPINNED.
  The package is a requirements.txt dependency whose source is not in the
  container; the reference has no test pinning its outputs.  The restatement
  in ``oracle/metrics_np.py`` follows the published definitions and is
  cross-checked against scikit-learn.
* segmentation-models-pytorch 0.3.3 / timm backbones used by ``basic`` and
  ``csnet``: PARITY UNPINNED (not installable offline).  ``CSNet``'s flat-walk
  semantics are pinned by running the reference ``CSNet`` class over a
  stand-in backbone honouring its naming contract.
"""
