"""CPU oracle for the SCONE / MACARONS next-best-view scoring path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU, the arithmetic of the reference
(Anttwo/MACARONS @ b23c180) for the hot path that `macarons_b200` implements in CUDA.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it, and only as the checker or the timed CPU baseline -- never as part of
the product path.  `macarons_b200` itself raises if its CUDA library is missing; it has no CPU
fallback and never imports this package.

Pinning status
--------------
The reference ships no tests, golden vectors or known-answer fixtures for this path
(SURVEY.md section 4 / 8c), so the pin is the reference's own Python executed in the build
container:

* `tests/golden/make_golden.py` imports the unmodified reference from `/root/reference`
  (through `tests/golden/ref_shim.py`, which only stubs the never-executed pytorch3d /
  matplotlib imports), runs it on seeded inputs and writes `tests/golden/*.npz`;
  at generation time it also asserts that this oracle reproduces the reference **bit for bit**
  on the same machine (same torch CPU kernels, same operation order).
* `tests/test_oracle_golden.py` (CPU, no reference needed) re-checks the oracle against those
  committed fixtures, and against an independent trig-free float64 closed form.

Each function cites the reference file:line it follows.
"""
