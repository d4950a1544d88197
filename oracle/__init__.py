"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Nothing under `oracle/` is part of the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker or the CPU baseline.  The product
package (`robust-e-nerf_b200/ren_b200`) never imports it and fails loudly when its
CUDA library is missing.

Contents
  shims/           pure-Python stand-ins for the un-installable third-party
                   packages the reference imports (easydict, pytorch_lightning,
                   nerfacc, roma, lpips, torchmetrics, tinycudann); the nerfacc and
                   roma shims restate those libraries' published algorithms.
  ref_loader.py    imports the UNMODIFIED reference from /root/reference on top of
                   the shims (this container only; the reference cannot travel).
  restatement.py   self-contained torch-CPU restatement of the hot path (travels to
                   the GPU box); validated against ref_loader in tests.
  synth.py         synthetic 640x480 event-stream scene in the reference's on-disk
                   schema + batch sampler.
  make_golden.py   mints tests/golden/*.pt by running the reference itself.

PARITY UNPINNED: the reference ships no tests / golden vectors (SURVEY.md
section 4), and nerfacc==0.3.1 / roma==1.2.7 are not vendored.  Golden vectors are
minted here from the reference's own Python running on the shims.
"""
