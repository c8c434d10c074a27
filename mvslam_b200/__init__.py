"""mvslam_b200 — B200 (sm_100a) implementation of mvSLAM's two-view front-end hot path.

The product is the C-ABI shared library ``libmvslam_b200.so`` (sources in ``csrc/``, interface in
``include/mvslam_b200.h``); this package is the thin Python host-side mirror used by the tests and
the benchmark.  There is no CPU fallback: everything here fails loudly when the CUDA library or a
CUDA device is missing.
"""
from .capi import (  # noqa: F401
    Context, Comm, MvsError, PairResult, MatchParams, RansacParams, OrbParams, PnpParams, PNP_RESULT_DTYPE, BaParams, BA_OBS_DTYPE, BA_RESULT_DTYPE, pnp_sample_table, MATCH_DTYPE, RESULT_DTYPE, KEYPOINT_DTYPE,
    SCORE_ALGEBRAIC, SCORE_SAMPSON, OK, E_BAD_ARG, E_TOO_FEW_POINTS, E_NO_MODEL, E_TOO_FEW_INLIERS,
    E_NO_CHEIRALITY, E_CUDA, E_CAPACITY, E_UNSUPPORTED, STAGES, lib_path, load_library, sample_table,
)
from .vision import VisualFeature, match_visual_features, sfm_solve, sfm_triangulate, image_pairs, pnp_solve  # noqa: F401
