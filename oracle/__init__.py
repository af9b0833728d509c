"""CPU oracle for the offline actor-critic gradient step  --  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference
(zhaoyizhou1123/OfflineRL-Kit) for the hot path named in BASELINE.json.  The
reference is pure Python on top of PyTorch (third-party, torch 2.11.0+cu128 in
this image; the reference itself pins no version, setup.py:14-24), so the
oracle is functional PyTorch-on-CPU code: forward passes are written out op by
op, the backward pass is torch autograd (as in the reference) and the
optimiser is ``torch.optim.Adam`` (as constructed in run_example/run_cql.py:92-94).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``  --  as the checker
or as the timed CPU baseline, never as part of the product path.  The product
(``offlinerl-kit_b200``) never imports it and raises if its CUDA library is
missing.

Parity pinning: the reference ships NO tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, run in the authoring container: ``tests/golden/make_golden.py`` imports
the read-only reference (with stub gym/diffusers/wandb/matplotlib modules),
runs ``policy.learn`` / ``buffer.sample`` / ``dynamics.learn`` / ``dynamics.step``
on seeded inputs and stores inputs, noise, losses and post-step parameters in
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every oracle
function against those files (CPU, ``-m "not gpu"``).
"""
