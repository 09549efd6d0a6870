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
