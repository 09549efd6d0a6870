"""B200-native linearisation path of powei-lin/camera-intrinsic-calibration-rs.

The product is libccrs_b200.so (CUDA, sm_100a) behind include/ccrs_b200.h; this package is the thin
Python host mirror of the reference interface used by tests and bench. The directory name carries a
hyphen, so import it with importlib.import_module("camera-intrinsic-calibration-rs_b200") or through
the root-level alias module `ccrs_b200`.
"""
from . import _abi
from ._abi import CcrsError, Options, Summary, default_options, LIB_PATH, SYMBOLS
from .calib import (MODELS, FeaturePoint, FrameFeature, GenericModel, JointProblem, Problem, RvecTvec,
                    calib_all_camera_with_extrinsics, calib_camera, comm_unique_id, convert_model, init_poses, init_ucm, initial_poses,
                    measure_fp64_peak, model_bounds, pack_frames, validation)
from . import synth
from . import io
from . import models
from . import dist

__all__ = ["CcrsError", "Options", "Summary", "default_options", "LIB_PATH", "SYMBOLS", "MODELS", "FeaturePoint",
           "FrameFeature", "GenericModel", "JointProblem", "calib_all_camera_with_extrinsics", "Problem", "RvecTvec", "calib_camera", "comm_unique_id", "measure_fp64_peak",
           "model_bounds", "pack_frames", "validation", "convert_model", "init_poses", "initial_poses", "init_ucm", "synth", "models", "io", "dist"]
