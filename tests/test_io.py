"""Result files in the reference's formats (camera-intrinsic-calibration-rs_b200/io.py). The EUCM model file is pinned
by the reference's own data/eucm.json (copied here verbatim as a fixture: /root/reference is not present on the GPU box)."""
import json

import numpy as np

EUCM_JSON = """{
  "EUCM": {
    "fx": 190.89618687183938,
    "fy": 190.87022285882367,
    "cx": 254.9375370481962,
    "cy": 256.86414483060787,
    "alpha": 0.6283550447635853,
    "beta": 1.0458678747533083,
    "width": 512,
    "height": 512
  }
}"""


def test_model_json_matches_reference_fixture(pkg, tmp_path):
    p = tmp_path / "eucm.json"
    p.write_text(EUCM_JSON)
    cam = pkg.io.model_from_json(str(p))
    assert cam.model == "eucm" and (cam.width, cam.height) == (512, 512)
    assert cam.params.tolist() == [190.89618687183938, 190.87022285882367, 254.9375370481962, 256.86414483060787,
                                   0.6283550447635853, 1.0458678747533083]
    # the synthetic ground truth of the whole test-suite is this file scaled x2
    gt = np.array(pkg.synth.GT_PARAMS["eucm"])
    assert np.allclose(gt[:4], 2 * cam.params[:4], atol=0.02) and np.array_equal(gt[4:], cam.params[4:])
    out = tmp_path / "out.json"
    pkg.io.model_to_json(str(out), cam)
    assert out.read_text() == EUCM_JSON                     # byte-identical round trip (field order, indentation)


def test_all_variants_round_trip(pkg, tmp_path):
    for m in ("ucm", "eucm", "eucmt", "kb4", "opencv5", "ftheta"):
        cam = pkg.GenericModel(m, np.array(pkg.synth.GT_PARAMS[m], dtype=np.float64), 1024, 1024)
        f = tmp_path / f"{m}.json"
        pkg.io.model_to_json(str(f), cam)
        back = pkg.io.model_from_json(str(f))
        assert back.model == m and np.array_equal(back.params, cam.params) and back.width == 1024
        assert list(json.loads(f.read_text())) == [pkg.io.VARIANT[m]]


def test_poses_and_report(pkg, tmp_path):
    rt = {7: pkg.RvecTvec((0.1, -0.2, 0.3), (1.0, 2.0, 3.0)), 2: pkg.RvecTvec((0.0, 0.0, 0.0), (0.5, 0.25, 0.125))}
    f = tmp_path / "cam0_poses.json"
    pkg.io.poses_to_json(str(f), rt)
    d = json.loads(f.read_text())
    assert list(d) == ["2", "7"] and d["7"] == {"rvec": [0.1, -0.2, 0.3], "tvec": [1.0, 2.0, 3.0]}   # BTreeMap order
    assert pkg.io.poses_from_json(str(f)) == rt
    r = tmp_path / "report.txt"
    pkg.io.write_report(str(r), True, [(0.123456, 0.1), (1.0, 0.987654)])
    assert r.read_text() == ("Calibrate with extrinsics: true\n\ncam0:\n    average reprojection error: 0.12346 px\n"
                             "    median  reprojection error: 0.10000 px\n\ncam1:\n    average reprojection error: 1.00000 px\n"
                             "    median  reprojection error: 0.98765 px\n\n")


def test_floats_are_printed_like_serde_json(pkg, tmp_path):
    """ryu / serde_json layout of the shortest round-trip digits: no padded or signed exponents, plain
    decimals down to 1e-5; the files still parse to the same values."""
    import json
    f = pkg.io.ryu_float
    assert f(1e-5) == "0.00001" and f(5e-5) == "0.00005" and f(-2.5e-5) == "-0.000025"
    assert f(1e-7) == "1e-7" and f(1.5e16) == "1.5e16" and f(1e16) == "1e16" and f(-3.25e-9) == "-3.25e-9"
    assert f(0.1) == "0.1" and f(1.0) == "1.0" and f(123456.789) == "123456.789" and f(1e15) == "1000000000000000.0"
    for x in (1e-5, 5e-5, 1e-7, 1.5e16, 0.1, 2.0 / 3.0, 1e-300, 1.7976931348623157e308):
        assert float(f(x)) == x
    cam = pkg.GenericModel("kb4", [400.0, 401.5, 512.0, 500.25, 1e-5, -3.5e-7, 2e-9, 0.0], 1024, 1024)
    p = tmp_path / "cam0.json"
    pkg.io.model_to_json(str(p), cam)
    txt = p.read_text()
    assert "e-0" not in txt and "e+" not in txt and '"k1": 0.00001' in txt and '"k2": -3.5e-7' in txt
    assert json.loads(txt) == pkg.io.model_to_dict(cam)
    back = pkg.io.model_from_json(str(p))
    assert list(back.params) == list(cam.params)


def test_shard_frames_refuses_more_ranks_than_frames(pkg):
    import numpy as np, pytest
    with pytest.raises(ValueError):
        pkg.dist.shard_frames(np.array([0, 10, 20, 30]), 0, 4)
    assert pkg.dist.shard_frames(np.array([0, 10, 20, 30]), 2, 3) == (2, 3)


def test_shard_frames_never_hands_out_an_empty_shard(pkg):
    import numpy as np
    fo = np.array([0, 1000, 1001, 1002, 1003, 1004])          # one frame holds nearly everything
    cuts = [pkg.dist.shard_frames(fo, r, 4) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 5 and all(hi > lo for lo, hi in cuts)
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(3))
