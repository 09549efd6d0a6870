"""Result files in the reference's formats (host-side I/O, like src/io.rs and camera-intrinsic-model's model_to_json):

  * cam{i}.json        — `model_to_json`: {"<Variant>": {fx, fy, cx, cy, <distortion...>, width, height}}; the EUCM layout is
                          pinned by the reference's data/eucm.json, the field names of the other variants follow the
                          parameter names of SURVEY App. A (the crate is not vendored: unpinned);
  * cam{i}_poses.json  — `object_to_json(BTreeMap<usize, RvecTvec>)` (src/bin/camera_calibration.rs:278-293,
                          src/types.rs:13-17): {"<frame>": {"rvec": [x, y, z], "tvec": [x, y, z]}} with ascending keys;
  * report.txt         — `write_report` (src/io.rs:21-31).
"""
from __future__ import annotations

import json
from typing import Dict, Sequence, Tuple

import numpy as np

from .calib import GenericModel, RvecTvec

VARIANT = {"ucm": "UCM", "eucm": "EUCM", "eucmt": "EUCMT", "kb4": "KannalaBrandt4", "opencv5": "OpenCVModel5", "ftheta": "Ftheta"}
FIELDS = {"ucm": ["alpha"], "eucm": ["alpha", "beta"], "eucmt": ["alpha", "beta", "t1", "t2"], "kb4": ["k1", "k2", "k3", "k4"],
          "opencv5": ["k1", "k2", "p1", "p2", "k3"], "ftheta": ["k1", "k2", "k3", "k4"]}
_BY_VARIANT = {v: k for k, v in VARIANT.items()}


def ryu_float(x: float) -> str:
    """A float as serde_json (ryu) prints it: the shortest digits that round-trip — what Python's repr finds too — laid
    out by ryu's rules: plain decimals for 1e-5 <= |x| < 1e16 (Python switches to an exponent below 1e-4), an exponent
    without padding or sign otherwise (1e-7, 1.5e16, not 1e-07 / 1.5e+16)."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return "null"                                   # serde_json writes non-finite floats as null
    r = repr(x)
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        e = int(exp)
        if -5 <= e < 0:                                 # Python: 5e-05; ryu: 0.00005
            digits = mant.replace("-", "").replace(".", "")
            return ("-" if x < 0 else "") + "0." + "0" * (-e - 1) + digits
        return f"{mant}e{e}"
    return r


def _dumps(obj, indent=2, level=0) -> str:
    """json.dumps(obj, indent=2) with floats printed by ryu_float (serde_json::to_string_pretty)."""
    pad, pad_in = " " * (indent * level), " " * (indent * (level + 1))
    if isinstance(obj, dict):
        if not obj:
            return "{}"
        return "{\n" + ",\n".join(f"{pad_in}{json.dumps(str(k))}: {_dumps(v, indent, level + 1)}" for k, v in obj.items()) + "\n" + pad + "}"
    if isinstance(obj, (list, tuple)):
        if not obj:
            return "[]"
        return "[\n" + ",\n".join(pad_in + _dumps(v, indent, level + 1) for v in obj) + "\n" + pad + "]"
    if isinstance(obj, bool) or obj is None or isinstance(obj, (int, str)):
        return json.dumps(obj)
    return ryu_float(obj)


def model_to_dict(cam: GenericModel) -> dict:
    names = ["fx", "fy", "cx", "cy"] + FIELDS[cam.model]
    body = {n: float(v) for n, v in zip(names, np.asarray(cam.params, dtype=np.float64))}
    body["width"] = int(cam.width); body["height"] = int(cam.height)
    return {VARIANT[cam.model]: body}


def model_from_dict(d: dict) -> GenericModel:
    (variant, body), = d.items()
    model = _BY_VARIANT[variant]
    names = ["fx", "fy", "cx", "cy"] + FIELDS[model]
    return GenericModel(model, np.array([body[n] for n in names], dtype=np.float64), int(body["width"]), int(body["height"]))


def model_to_json(path: str, cam: GenericModel) -> None:
    with open(path, "w") as f:
        f.write(_dumps(model_to_dict(cam)))     # serde_json::to_string_pretty: two-space indent, ryu floats


def model_from_json(path: str) -> GenericModel:
    with open(path) as f:
        return model_from_dict(json.load(f))


def poses_to_json(path: str, rtvecs: Dict[int, RvecTvec]) -> None:
    ordered = {str(k): {"rvec": [float(x) for x in rtvecs[k].rvec], "tvec": [float(x) for x in rtvecs[k].tvec]} for k in sorted(rtvecs)}
    with open(path, "w") as f:
        f.write(_dumps(ordered))


def poses_from_json(path: str) -> Dict[int, RvecTvec]:
    with open(path) as f:
        d = json.load(f)
    return {int(k): RvecTvec(tuple(v["rvec"]), tuple(v["tvec"])) for k, v in d.items()}


def write_report(path: str, with_extrinsic: bool, rep_rms: Sequence[Tuple[float, float]]) -> None:
    """src/io.rs:21-31 (the pairs are what `validation` returns per camera; the labels are the reference's)."""
    s = f"Calibrate with extrinsics: {'true' if with_extrinsic else 'false'}\n\n"
    for i, (avg_rep, med_rep) in enumerate(rep_rms):
        s += f"cam{i}:\n    average reprojection error: {avg_rep:.5f} px\n    median  reprojection error: {med_rep:.5f} px\n\n"
    with open(path, "w") as f:
        f.write(s)
