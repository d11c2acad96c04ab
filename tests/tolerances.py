"""Stated fp32 tolerances for the SH-integration stage (rows a3-a5), and why.

* coverage gain, (B,C) means in [0,1]:   |cuda - oracle| <= 1e-5   (observed ~5e-7).
* per-point visibility gains, (B,C,P):
    - against the float64 closed form (the mathematical truth):  <= 2e-5 * max(1, coef_scale)
    - against the fp32 oracle (= the reference's own arithmetic): <= 2e-5 on rays where the
      reference is well conditioned, <= 5e-3 everywhere.  The reference recovers the azimuth with
      acos(z / (r cos(elev))) (utility/CustomGeometry.py:39), which amplifies one ulp of its argument
      to ~3e-4 rad when |cos(phi)| -> 1 and that error is multiplied by m <= 7: its own fp32 output is
      up to ~1e-3 away from the exact value on such rays (measured: median 6e-8, p99.9 2e-5, max 1e-3;
      tests/test_oracle_golden.py::test_reference_fp32_conditioning).  The CUDA kernel is trig-free and
      stays within ~4e-6 of the exact value everywhere, so on those rays the two legitimately differ.
"""
import numpy as np

COVERAGE_ATOL = 1e-5
VIS_F64_ATOL = 2e-5
VIS_ORACLE_ATOL_WELL = 2e-5
VIS_ORACLE_ATOL_ANY = 5e-3


def ray_conditioning(pts, cams):
    """|cos(phi)| and |cos(theta)| of every camera->point ray, (B,C,P) each."""
    d = np.asarray(cams, np.float64)[:, :, None, :] - np.asarray(pts, np.float64)[:, None, :, :3]
    rho = np.hypot(d[..., 0], d[..., 2])
    return np.abs(d[..., 2]) / rho, np.abs(d[..., 1]) / np.linalg.norm(d, axis=-1)


def assert_visibility_close(got, oracle32, truth64, pts, cams, coef_scale=1.0):
    got = np.asarray(got, np.float64)
    s = max(1.0, float(coef_scale))
    e64 = np.abs(got - truth64)
    assert e64.max() <= VIS_F64_ATOL * s, "vs float64 closed form: %g" % e64.max()
    e32 = np.abs(got - np.asarray(oracle32, np.float64))
    cphi, ct = ray_conditioning(pts, cams)
    well = (cphi < 0.99) & (ct < 0.99)
    assert e32[well].max() <= VIS_ORACLE_ATOL_WELL * s, "vs fp32 oracle (well-conditioned rays): %g" % e32[well].max()
    assert e32.max() <= VIS_ORACLE_ATOL_ANY * s, "vs fp32 oracle: %g" % e32.max()
