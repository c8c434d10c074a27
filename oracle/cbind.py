"""ctypes binding of the C oracle (oracle/mvs_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package (mvslam_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmvs_oracle.so")

OK, E_BAD_ARG, E_TOO_FEW_POINTS, E_NO_MODEL, E_TOO_FEW_INLIERS, E_NO_CHEIRALITY = range(6)
SCORE_ALGEBRAIC, SCORE_SAMPSON = 0, 1
SOLVER_REFERENCE, SOLVER_FAST = 0, 1
_SOLVERS = {"reference": SOLVER_REFERENCE, "fast": SOLVER_FAST, SOLVER_REFERENCE: SOLVER_REFERENCE, SOLVER_FAST: SOLVER_FAST}
# solver used by every geometry wrapper below unless the call names one ("reference" = literal A^T A + cv::SVDecomp
# restatement, "fast" = Householder/round-robin-Jacobi/fma contract); tests switch it with set_default_solver()
DEFAULT_SOLVER = "reference"


def set_default_solver(name):
    global DEFAULT_SOLVER
    assert name in ("reference", "fast")
    DEFAULT_SOLVER = name


def _use(solver):
    lib().orc_set_solver(_SOLVERS[DEFAULT_SOLVER if solver is None else solver])

MATCH_DTYPE = np.dtype([("query", np.int32), ("train", np.int32), ("distance", np.float32)])


class PairResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_matches", C.c_int32), ("n_inliers", C.c_int32),
        ("best_hypothesis", C.c_int32), ("n_points", C.c_int32), ("candidate", C.c_int32),
        ("residual", C.c_double), ("F", C.c_double * 9), ("E", C.c_double * 9),
        ("R1to2", C.c_double * 9), ("t1to2", C.c_double * 3),
        ("R2in1", C.c_double * 9), ("t2in1", C.c_double * 3),
    ]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("mvs_oracle.c", "mvs_oracle.h", "pnp_oracle.c", "pnp_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_svd.restype = C.c_int
        _lib.orc_match_hamming.restype = C.c_int
        _lib.orc_match_l2.restype = C.c_int
        _lib.orc_filter_matches.restype = C.c_int
        _lib.orc_count_inliers.restype = C.c_int
        _lib.orc_triangulate_points.restype = C.c_int
        _lib.orc_sfm_triangulate.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def knn2_hamming(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nb = q.shape
    idx = np.empty((nq, 2), np.int32); dist = np.empty((nq, 2), np.int32)
    lib().orc_knn2_hamming(_p(q), nq, _p(t), t.shape[0], nb, _p(idx), _p(dist))
    return idx, dist


def knn2_l2(q, t):
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    nq, dim = q.shape
    idx = np.empty((nq, 2), np.int32); dist = np.empty((nq, 2), np.float32)
    lib().orc_knn2_l2(_p(q), nq, _p(t), t.shape[0], dim, _p(idx), _p(dist))
    return idx, dist


def match_hamming(q, t, ratio=0.7, max_dist=-1.0, cross_check=False):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    out = np.empty(max(q.shape[0], 1), MATCH_DTYPE)
    n = lib().orc_match_hamming(_p(q), q.shape[0], _p(t), t.shape[0], q.shape[1],
                                C.c_double(ratio), C.c_double(max_dist), int(cross_check), _p(out))
    return out[:n].copy()


def match_l2(q, t, ratio=0.7, max_dist=-1.0, cross_check=False):
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    out = np.empty(max(q.shape[0], 1), MATCH_DTYPE)
    n = lib().orc_match_l2(_p(q), q.shape[0], _p(t), t.shape[0], q.shape[1],
                           C.c_double(ratio), C.c_double(max_dist), int(cross_check), _p(out))
    return out[:n].copy()


def svd(A, solver=None):
    _use(solver)
    A = _f64(A); n = A.shape[0]
    U = np.empty((n, n)); w = np.empty(n); Vt = np.empty((n, n))
    sweeps = lib().orc_svd(n, _p(A), _p(U), _p(w), _p(Vt))
    return U, w, Vt, sweeps


def so3_rectify(R):
    out = np.empty((3, 3)); lib().orc_so3_rectify(_p(_f64(R)), _p(out)); return out


def se3_inverse(R, t):
    Ro = np.empty((3, 3)); to = np.empty(3)
    lib().orc_se3_inverse(_p(_f64(R)), _p(_f64(t)), _p(Ro), _p(to)); return Ro, to


def normalize_points(K, xy):
    xy = _f64(xy); out = np.empty((xy.shape[0], 3))
    lib().orc_normalize_points(_p(_f64(K)), _p(xy), xy.shape[0], _p(out)); return out


def find_fundamental_matrix(p1s, p2s, solver=None):
    _use(solver)
    F = np.empty((3, 3))
    lib().orc_find_fundamental_matrix(_p(_f64(p1s)), _p(_f64(p2s)), _p(F)); return F


def sample_table(seed, pair_id, n_points, H):
    out = np.empty((H, 8), np.uint32)
    lib().orc_sample_table(C.c_uint64(seed), C.c_uint64(pair_id), C.c_uint32(n_points), H, _p(out))
    return out


def count_inliers(p1, p2, F, max_error_sq, mode=SCORE_ALGEBRAIC, solver=None):
    _use(solver)
    p1 = _f64(p1); p2 = _f64(p2); n = p1.shape[0]
    mask = np.empty(n, np.uint8); res = C.c_double()
    cnt = lib().orc_count_inliers(_p(p1), _p(p2), n, _p(_f64(F)), C.c_double(max_error_sq), mode,
                                  _p(mask), C.byref(res))
    return cnt, res.value, mask


def ransac_fundamental(p1, p2, samples, max_error_sq, mode=SCORE_ALGEBRAIC, want_all=False, solver=None):
    _use(solver)
    p1 = _f64(p1); p2 = _f64(p2); n = p1.shape[0]
    samples = np.ascontiguousarray(samples, np.uint32); H = samples.shape[0]
    F = np.zeros((3, 3)); mask = np.zeros(max(n, 1), np.uint8)
    cnt = C.c_int(); res = C.c_double(); bh = C.c_int()
    allc = np.zeros(H, np.int32) if want_all else None
    allF = np.zeros((H, 9)) if want_all else None
    st = lib().orc_ransac_fundamental(_p(p1), _p(p2), n, _p(samples), H, C.c_double(max_error_sq), mode,
                                      _p(F), _p(mask), C.byref(cnt), C.byref(res), C.byref(bh),
                                      _p(allc), _p(allF))
    out = dict(status=st, F=F, mask=mask[:n], count=cnt.value, residual=res.value, best_h=bh.value)
    if want_all:
        out["all_counts"] = allc; out["all_F"] = allF.reshape(H, 3, 3)
    return out


def project_essential(F, solver=None):
    _use(solver)
    E = np.empty((3, 3)); lib().orc_project_essential(_p(_f64(F)), _p(E)); return E


def decompose_essential(E, solver=None):
    _use(solver)
    Ra = np.empty((3, 3)); Rb = np.empty((3, 3)); t = np.empty(3)
    lib().orc_decompose_essential(_p(_f64(E)), _p(Ra), _p(Rb), _p(t)); return Ra, Rb, t


def triangulate_points(R, t, p1, p2, mask=None, solver=None):
    _use(solver)
    p1 = _f64(p1); p2 = _f64(p2); n = p1.shape[0]
    pts = np.empty((max(n, 1), 3)); idx = np.empty(max(n, 1), np.uint64)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    m = lib().orc_triangulate_points(_p(_f64(R)), _p(_f64(t)), _p(p1), _p(p2), _p(mask), n, _p(pts), _p(idx))
    return pts[:m].copy(), idx[:m].copy()


def _result_dict(r):
    d = {k: getattr(r, k) for k in ("status", "n_matches", "n_inliers", "best_hypothesis", "n_points",
                                    "candidate", "residual")}
    for k in ("F", "E", "R1to2", "R2in1"):
        d[k] = np.array(getattr(r, k)).reshape(3, 3)
    for k in ("t1to2", "t2in1"):
        d[k] = np.array(getattr(r, k))
    return d


def sfm_solve(xy1, xy2, K, samples=None, H=1, seed=0, pair_id=0, mode=SCORE_ALGEBRAIC, max_error_sq=0.0, solver=None):
    _use(solver)
    xy1 = _f64(xy1); xy2 = _f64(xy2); n = xy1.shape[0]
    if samples is not None:
        samples = np.ascontiguousarray(samples, np.uint32); H = samples.shape[0]
    res = PairResult(); mask = np.zeros(max(n, 1), np.uint8)
    pts = np.empty((max(n, 1), 3)); idx = np.empty(max(n, 1), np.uint64)
    lib().orc_sfm_solve(_p(xy1), _p(xy2), n, _p(_f64(K)), _p(samples), H, C.c_uint64(seed),
                        C.c_uint64(pair_id), mode, C.c_double(max_error_sq), C.byref(res), _p(mask), _p(pts), _p(idx))
    d = _result_dict(res)
    d["mask"] = mask[:n]; d["points"] = pts[:res.n_points].copy(); d["indexes"] = idx[:res.n_points].copy()
    return d


def sfm_triangulate(xy1, xy2, K, R1, t1, R2, t2, solver=None):
    _use(solver)
    xy1 = _f64(xy1); xy2 = _f64(xy2); n = xy1.shape[0]
    pts = np.empty((max(n, 1), 3)); idx = np.empty(max(n, 1), np.uint64)
    m = lib().orc_sfm_triangulate(_p(xy1), _p(xy2), n, _p(_f64(K)), _p(_f64(R1)), _p(_f64(t1)),
                                  _p(_f64(R2)), _p(_f64(t2)), _p(pts), _p(idx))
    return pts[:m].copy(), idx[:m].copy()


def image_pair(desc1, kp1, desc2, kp2, K, ratio=0.7, max_dist=-1.0, cross_check=False, H=1, seed=0,
               pair_id=0, mode=SCORE_ALGEBRAIC, max_error_sq=0.0, solver=None):
    _use(solver)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    kp1 = np.ascontiguousarray(kp1, np.float32); kp2 = np.ascontiguousarray(kp2, np.float32)
    n1, n2 = desc1.shape[0], desc2.shape[0]
    res = PairResult(); cap = max(n2, 1)
    matches = np.empty(cap, MATCH_DTYPE); mask = np.zeros(cap, np.uint8)
    pts = np.empty((cap, 3)); idx = np.empty(cap, np.uint64)
    lib().orc_image_pair(_p(desc1), _p(kp1), n1, _p(desc2), _p(kp2), n2, desc1.shape[1], _p(_f64(K)),
                         C.c_double(ratio), C.c_double(max_dist), int(cross_check), H, C.c_uint64(seed),
                         C.c_uint64(pair_id), mode, C.c_double(max_error_sq), C.byref(res), _p(matches), _p(mask), _p(pts), _p(idx))
    d = _result_dict(res)
    d["matches"] = matches[:res.n_matches].copy(); d["mask"] = mask[:res.n_matches].copy()
    d["points"] = pts[:res.n_points].copy(); d["indexes"] = idx[:res.n_points].copy()
    return d


def pair_batch(descs, kps, pairs, K, ratio=0.7, max_dist=-1.0, cross_check=False, H=1, seed=0,
               mode=SCORE_ALGEBRAIC, threads=0, max_error_sq=0.0, solver=None):
    _use(solver)
    descs = [np.ascontiguousarray(d, np.uint8) for d in descs]
    kps = [np.ascontiguousarray(k, np.float32) for k in kps]
    nf = len(descs)
    dptr = (C.c_void_p * nf)(*[d.ctypes.data for d in descs])
    kptr = (C.c_void_p * nf)(*[k.ctypes.data for k in kps])
    counts = np.array([d.shape[0] for d in descs], np.int32)
    pairs = np.ascontiguousarray(pairs, np.int32); npairs = pairs.shape[0]
    res = (PairResult * npairs)()
    lib().orc_pair_batch(dptr, kptr, _p(counts), nf, _p(pairs), npairs, descs[0].shape[1], _p(_f64(K)),
                         C.c_double(ratio), C.c_double(max_dist), int(cross_check), H, C.c_uint64(seed), mode,
                         C.c_double(max_error_sq), threads, res)
    return [_result_dict(r) for r in res]


def max_threads():
    return lib().orc_max_threads()


# ---------------------------------------------------------------- pnp_solve (oracle/pnp_oracle.c)
def pnp_sample_table(seed, problem_id, n_points, H):
    out = np.empty((H, 4), np.uint32)
    lib().orc_pnp_sample_table(C.c_uint64(seed), C.c_uint64(problem_id), C.c_uint32(n_points), int(H), _p(out))
    return out


def solve_quartic(c):
    c = np.ascontiguousarray(c, np.float64); r = np.zeros(4)
    n = lib().orc_solve_quartic(_p(c), _p(r))
    return r[:n]


def p3p(bearings, world):
    f = np.ascontiguousarray(bearings, np.float64).reshape(3, 3); X = np.ascontiguousarray(world, np.float64).reshape(3, 3)
    R = np.zeros((4, 3, 3)); t = np.zeros((4, 3))
    n = lib().orc_p3p(_p(f), _p(X), _p(R), _p(t))
    return R[:n], t[:n]


def pnp_hypotheses(world, image, K, samples):
    """(valid[H], R[H,3,3], t[H,3]) world->camera, one P3P hypothesis per sample row."""
    w = np.ascontiguousarray(world, np.float64); im = np.ascontiguousarray(image, np.float64)
    K = np.ascontiguousarray(K, np.float64); s = np.ascontiguousarray(samples, np.uint32)
    H = s.shape[0]
    R = np.zeros((H, 3, 3)); t = np.zeros((H, 3)); ok = np.zeros(H, bool)
    L = lib()
    for h in range(H):
        ok[h] = bool(L.orc_pnp_hypothesis(_p(w), _p(im), _p(s[h]), _p(K), _p(R[h]), _p(t[h])))
    return ok, R, t


def pnp_solve(world, image, K, samples=None, H=100, seed=0, problem_id=0, reproj_error=0.05, refine_iters=10):
    w = np.ascontiguousarray(world, np.float64); im = np.ascontiguousarray(image, np.float64)
    K = np.ascontiguousarray(K, np.float64); n = w.shape[0]
    if samples is not None:
        samples = np.ascontiguousarray(samples, np.uint32); H = samples.shape[0]
    R = np.zeros((3, 3)); t = np.zeros(3); mask = np.zeros(max(n, 1), np.uint8)
    Rp = np.zeros((3, 3)); tp = np.zeros(3); counts = np.zeros(H, np.int32)
    ni = C.c_int(0); bh = C.c_int(-1)
    st = lib().orc_pnp_solve(_p(w), _p(im), n, _p(K), _p(samples), int(H), C.c_uint64(seed), C.c_uint64(problem_id),
                             C.c_double(reproj_error), int(refine_iters), _p(R), _p(t), _p(mask), C.byref(ni), C.byref(bh),
                             _p(Rp), _p(tp), _p(counts))
    return dict(status=st, R=R, t=t, mask=mask[:n], n_inliers=ni.value, best_h=bh.value, R_p3p=Rp, t_p3p=tp, all_counts=counts)
