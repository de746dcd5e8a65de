"""CPU oracle for the S3OD `remove_background` hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain fp32 CPU restatement of the reference algorithm (torch CPU ops for the
floating-point network, numpy integer arithmetic for the uint8 resize / composite).  It is the
checker, never the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
/ `--impl reference` legs may import it.  Nothing under `s3od_b200/` imports it, and the product
path raises when the CUDA library is missing instead of falling back to this code.

Parity pinning: the reference keeps no golden tensors for this path.  The oracle is
pinned against the *reference itself*, imported from /root/reference/src in the build container by
`oracle/make_golden.py`, which writes `tests/golden/*.npz` (reference outputs on seeded weights and
inputs).  `tests/test_oracle_golden.py` checks the oracle against those fixtures on every run.
"""
