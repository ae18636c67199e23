"""TEST INFRASTRUCTURE (oracle) -- PSS-78 practical salinity, restated.

The reference calls ``gsw.SP_from_C(C, T, z)`` (reference parse.py:132, package
gsw==3.3.1 pinned in reference requirements.txt:4).  gsw is a third-party
dependency that is neither vendored under /root/reference nor installed in
this image, so its published algorithm (GSW-C ``gsw_sp_from_c`` and
``gsw_hill_ratio_at_sp2``: PSS-78 with the Hill et al. 1986 extension below
SP = 2) is restated here from the public description.

Pinning: the reference ships no test or golden vector for this call and gsw
itself cannot be run here.  External anchors, asserted in
tests/test_oracle_units.py: the UNESCO 1983 check value (R=1.888091, t68=40,
p=10000 -> S=40.00000) and the six-point example of the GSW documentation for
gsw_SP_from_C (matched to the last bit).  The Hill extension below SP = 2 has
no published vector: that branch stays PARITY UNPINNED.

Works on python floats and numpy arrays alike (used (a) as the ``gsw`` stand-in
when the real reference is run under oracle/ref_shim.py and (b) by the numpy
restatement oracle/axctd_oracle.py).  Never imported by the product package.
"""
import numpy as np

_A = (0.0080, -0.1692, 25.3851, 14.0941, -7.0261, 2.7081)
_B = (0.0005, -0.0056, -0.0066, -0.0375, 0.0636, -0.0144)
_C = (0.6766097, 2.00564e-2, 1.104259e-4, -6.9698e-7, 1.0031e-9)
_D = (3.426e-2, 4.464e-4, 4.215e-1, -3.107e-3)
_E = (2.070e-5, -6.370e-10, 3.989e-15)
_K = 0.0162
_G = (2.641463563366498e-1, 2.007883247811176e-4, -4.107694432853053e-6,
      8.401670882091225e-8, -1.711392021989210e-9, 3.374193893377380e-11,
      -5.923731174730784e-13, 8.057771569962299e-15, -7.054313817447962e-17,
      2.859992717347235e-19)
C3515_INV = 0.023302418791070513  # 1 / 42.9140 mS/cm


def _sp_poly(rtx, ft68):
    a0, a1, a2, a3, a4, a5 = _A
    b0, b1, b2, b3, b4, b5 = _B
    return (a0 + (a1 + (a2 + (a3 + (a4 + a5 * rtx) * rtx) * rtx) * rtx) * rtx
            + ft68 * (b0 + (b1 + (b2 + (b3 + (b4 + b5 * rtx) * rtx) * rtx) * rtx) * rtx))


def _dsp_drtx(rtx, ft68):
    a0, a1, a2, a3, a4, a5 = _A
    b0, b1, b2, b3, b4, b5 = _B
    return (a1 + (2 * a2 + (3 * a3 + (4 * a4 + 5 * a5 * rtx) * rtx) * rtx) * rtx
            + ft68 * (b1 + (2 * b2 + (3 * b3 + (4 * b4 + 5 * b5 * rtx) * rtx) * rtx) * rtx))


def hill_ratio_at_sp2(t):
    """GSW-C gsw_hill_ratio_at_sp2: ratio that makes the Hill extension
    continuous with PSS-78 at SP = 2 (one modified Newton step on Rtx)."""
    g = _G
    t68 = t * 1.00024
    ft68 = (t68 - 15.0) / (1.0 + _K * (t68 - 15.0))
    rtx0 = g[0] + t68 * (g[1] + t68 * (g[2] + t68 * (g[3] + t68 * (g[4] + t68 * (
        g[5] + t68 * (g[6] + t68 * (g[7] + t68 * (g[8] + t68 * g[9]))))))))
    dsp = _dsp_drtx(rtx0, ft68)
    sp_est = _sp_poly(rtx0, ft68)
    rtx = rtx0 - (sp_est - 2.0) / dsp
    rtxm = 0.5 * (rtx + rtx0)
    dsp = _dsp_drtx(rtxm, ft68)
    rtx = rtx0 - (sp_est - 2.0) / dsp
    x = 400.0 * rtx * rtx
    sqrty = 10.0 * rtx
    part1 = 1.0 + x * (1.5 + x)
    part2 = 1.0 + sqrty * (1.0 + sqrty * (1.0 + sqrty))
    return 2.0 / (2.0 - _A[0] / part1 - _B[0] * ft68 / part2)


def SP_from_C(C, t, p):
    """Practical salinity from conductivity [mS/cm], in-situ temperature
    [deg C, ITS-90] and pressure [dbar]; NaN where Rt < 0 or SP < 0."""
    scalar = np.ndim(C) == 0 and np.ndim(t) == 0 and np.ndim(p) == 0
    C = np.asarray(C, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    with np.errstate(all="ignore"):
        t68 = t * 1.00024
        ft68 = (t68 - 15.0) / (1.0 + _K * (t68 - 15.0))
        r = C3515_INV * C
        c0, c1, c2, c3, c4 = _C
        d1, d2, d3, d4 = _D
        e1, e2, e3 = _E
        rt_lc = c0 + (c1 + (c2 + (c3 + c4 * t68) * t68) * t68) * t68
        rp = 1.0 + (p * (e1 + e2 * p + e3 * p * p)) / (
            1.0 + d1 * t68 + d2 * t68 * t68 + (d3 + d4 * t68) * r)
        rt = r / (rp * rt_lc)
        rt = np.where(rt < 0.0, np.nan, rt)
        rtx = np.sqrt(rt)
        sp = _sp_poly(rtx, ft68)
        x = 400.0 * rt
        sqrty = 10.0 * rtx
        part1 = 1.0 + x * (1.5 + x)
        part2 = 1.0 + sqrty * (1.0 + sqrty * (1.0 + sqrty))
        sp_hill = hill_ratio_at_sp2(t) * (sp - _A[0] / part1 - _B[0] * ft68 / part2)
        sp = np.where(sp < 2.0, sp_hill, sp)
        sp = np.where(sp < 0.0, np.nan, sp)
    if scalar:
        return float(sp)
    return sp
