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
