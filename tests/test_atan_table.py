"""The table behind atan2_pos (ccrs_device.cuh) is generated, never edited: csrc/ccrs_atan_tab.inc must be what
tools/gen_atan_table.py writes, and its entries must satisfy theta_i = 2 atan(i / 64), sin^2 + cos^2 = 1."""
import math
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "camera-intrinsic-calibration-rs_b200", "csrc", "ccrs_atan_tab.inc")


def _rows():
    rows = []
    for line in open(INC):
        m = re.match(r"\s*\{([^}]*)\}", line)
        if m:
            rows.append([float.fromhex(t.strip()) if "x" in t else float(t) for t in m.group(1).split(",")])
    return rows


def test_table_entries():
    rows = _rows()
    assert len(rows) == 65
    for i, (th, s, c, pad) in enumerate(rows):
        assert pad == 0.0
        assert abs(th - 2.0 * math.atan(i / 64.0)) <= 2e-16 * max(th, 1e-300) + 1e-300
        assert abs(s - math.sin(th)) <= 1.2e-16 and abs(c - math.cos(th)) <= 1.2e-16
        assert abs(s * s + c * c - 1.0) < 4e-16
    assert rows[0][:3] == [0.0, 0.0, 1.0] and abs(rows[64][0] - math.pi / 2) < 1e-15


def test_table_is_what_the_generator_writes(tmp_path):
    import importlib.util, shutil
    mp = importlib.util.find_spec("mpmath")
    if mp is None:
        import pytest
        pytest.skip("mpmath not installed")
    import mpmath
    mpmath.mp.dps = 50
    for i, (th, s, c, _) in enumerate(_rows()):
        t = float(2 * mpmath.atan(mpmath.mpf(i) / 64))
        assert th == t and s == float(mpmath.sin(mpmath.mpf(t))) and c == float(mpmath.cos(mpmath.mpf(t)))


def test_algorithm_of_atan2_pos_on_a_dense_grid():
    """numpy restatement of atan2_pos (ccrs_device.cuh): FP32 key t = r / (rho + |z|), table entry i = round(64 t),
    theta = theta_i + asin((r cos_i - |z| sin_i) / rho) with the odd series to x^9, pi - theta for z < 0 — against
    numpy's arctan2 over 2 million (r, z) pairs from 1e-7 rad to 179.99 degrees: the error bound the kernel comment states.
    The key is also perturbed by +-1 table entry: an FP32 approximation (sqrt.approx / rcp.approx) that lands in the
    neighbouring entry still gives a correct angle (the series then sees |x| <= 3/128)."""
    import numpy as np
    tab = np.array(_rows())
    rng = np.random.default_rng(0)
    th = np.concatenate([10.0 ** rng.uniform(-7, 0, 500_000), rng.uniform(0.0, np.pi - 1e-4, 1_500_000)])
    rho = 10.0 ** rng.uniform(-2, 2, th.size)
    r, z = rho * np.sin(th), rho * np.cos(th)
    ref = np.arctan2(r, z)
    r2, rho2 = r * r, r * r + z * z
    irho = 1.0 / np.sqrt(rho2)
    za = np.abs(z)
    tf = np.sqrt(r2.astype(np.float32)) / (np.sqrt(rho2.astype(np.float32)) + za.astype(np.float32))
    for shift in (0, 1, -1):
        i = np.clip(np.rint(tf * np.float32(64.0)).astype(np.int64) + shift, 0, 64)
        xs = (r * tab[i, 2] - za * tab[i, 1]) * irho
        x2 = xs * xs
        p = x2 * (35.0 / 1152.0) + 5.0 / 112.0
        p = x2 * p + 3.0 / 40.0
        p = x2 * p + 1.0 / 6.0
        thp = tab[i, 0] + (xs * x2 * p + xs)
        out = np.where(z < 0.0, np.pi - thp, thp)
        err = np.abs(out - ref)
        assert err.max() < 1.4e-15, (shift, err.max())        # three ulp of pi: table rounding, the series, pi - theta
        big = ref >= 1.0 / 64
        assert (err[big] / ref[big]).max() < 3e-14, shift
        if shift == 0:
            small = ref < 1.0 / 128                      # entry 0 is exact: full relative precision at small angles
            assert (err[small] / ref[small]).max() < 1e-15
