"""ctypes binding of include/mvslam_b200.h (the C ABI of libmvslam_b200.so)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

(OK, E_BAD_ARG, E_TOO_FEW_POINTS, E_NO_MODEL, E_TOO_FEW_INLIERS, E_NO_CHEIRALITY, E_CUDA, E_CAPACITY,
 E_UNSUPPORTED) = range(9)
SCORE_ALGEBRAIC, SCORE_SAMPSON = 0, 1
SOLVER_REFERENCE, SOLVER_FAST = 0, 1
_SOLVERS = {"reference": SOLVER_REFERENCE, "fast": SOLVER_FAST, SOLVER_REFERENCE: SOLVER_REFERENCE, SOLVER_FAST: SOLVER_FAST}
# solver the geometry calls below use unless they name one (include/mvslam_b200.h, MVS_SOLVER_*)
DEFAULT_SOLVER = "reference"


def set_default_solver(name):
    global DEFAULT_SOLVER
    assert name in ("reference", "fast")
    DEFAULT_SOLVER = name


def _solver(solver):
    return _SOLVERS[DEFAULT_SOLVER if solver is None else solver]
STAGES = ("knn", "match_finalize", "hypotheses", "score", "select", "triangulate", "finalize", "l2",
          "orb_pyramid", "orb_fast", "orb_harris", "orb_select", "orb_blur", "orb_describe", "pnp", "ba")

MATCH_DTYPE = np.dtype([("query", np.int32), ("train", np.int32), ("distance", np.float32)])
RESULT_DTYPE = np.dtype([
    ("status", np.int32), ("n_matches", np.int32), ("n_inliers", np.int32), ("best_hypothesis", np.int32),
    ("n_points", np.int32), ("candidate", np.int32), ("residual", np.float64), ("F", np.float64, (3, 3)),
    ("E", np.float64, (3, 3)), ("R1to2", np.float64, (3, 3)), ("t1to2", np.float64, (3,)),
    ("R2in1", np.float64, (3, 3)), ("t2in1", np.float64, (3,)), ("match_inlier_ssd", np.uint64)])
KEYPOINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("size", np.float32), ("angle", np.float32),
                           ("response", np.float32), ("octave", np.int32)])


class MatchParams(C.Structure):
    _fields_ = [("ratio", C.c_double), ("max_dist", C.c_double), ("cross_check", C.c_int32), ("bounded", C.c_int32)]


class RansacParams(C.Structure):
    _fields_ = [("n_hypotheses", C.c_int32), ("score_mode", C.c_int32), ("max_error_sq", C.c_double),
                ("seed", C.c_uint64), ("min_inliers", C.c_int32), ("solver", C.c_int32),
                ("pair_id_base", C.c_uint64)]


class PairResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_matches", C.c_int32), ("n_inliers", C.c_int32),
                ("best_hypothesis", C.c_int32), ("n_points", C.c_int32), ("candidate", C.c_int32),
                ("residual", C.c_double), ("F", C.c_double * 9), ("E", C.c_double * 9),
                ("R1to2", C.c_double * 9), ("t1to2", C.c_double * 3), ("R2in1", C.c_double * 9),
                ("t2in1", C.c_double * 3), ("match_inlier_ssd", C.c_uint64)]


assert C.sizeof(PairResult) == RESULT_DTYPE.itemsize == 376


PNP_RESULT_DTYPE = np.dtype([
    ("status", np.int32), ("n_points", np.int32), ("n_inliers", np.int32), ("best_hypothesis", np.int32),
    ("R_c2w", np.float64, (3, 3)), ("t_c2w", np.float64, (3,)), ("R_w2c_p3p", np.float64, (3, 3)), ("t_w2c_p3p", np.float64, (3,))])


class PnpParams(C.Structure):
    _fields_ = [("n_hypotheses", C.c_int32), ("refine_iterations", C.c_int32), ("reprojection_error", C.c_double),
                ("seed", C.c_uint64), ("problem_id_base", C.c_uint64), ("min_inliers", C.c_int32), ("reserved", C.c_int32)]


assert PNP_RESULT_DTYPE.itemsize == 208 and C.sizeof(PnpParams) == 40


BA_OBS_DTYPE = np.dtype([("frame", np.int32), ("point", np.int32), ("uv", np.float64, (2,)), ("cov", np.float64, (3,))])
BA_RESULT_DTYPE = np.dtype([("status", np.int32), ("iterations", np.int32), ("initial_error", np.float64), ("final_error", np.float64)])


class BaParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("reserved", C.c_int32), ("lambda_initial", C.c_double),
                ("relative_tolerance", C.c_double), ("absolute_tolerance", C.c_double)]


assert BA_OBS_DTYPE.itemsize == 48 and BA_RESULT_DTYPE.itemsize == 24


class OrbParams(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("reserved", C.c_int32 * 3)]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * len(STAGES)), ("launches", C.c_uint64 * len(STAGES))]


class MvsError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"mvslam_b200 status {status}: {msg}")
        self.status = status


def lib_path():
    return os.path.join(_HERE, "libmvslam_b200.so")


_lib = None


def load_library():
    """Load libmvslam_b200.so.  Raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). mvslam_b200 has no CPU fallback.")
    L = C.CDLL(p)
    L.mvs_status_string.restype = C.c_char_p
    L.mvs_last_error.restype = C.c_char_p
    L.mvs_last_error.argtypes = [C.c_void_p]
    L.mvs_kernel_launches.restype = C.c_uint64
    L.mvs_kernel_launches.argtypes = [C.c_void_p]
    L.mvs_debug_guard_check.restype = C.c_int
    L.mvs_debug_guard_check.argtypes = [C.c_void_p]
    L.mvs_debug_guard_poke.restype = C.c_int
    L.mvs_debug_guard_poke.argtypes = [C.c_void_p]
    L.mvs_destroy.restype = None
    L.mvs_destroy.argtypes = [C.c_void_p]
    L.mvs_sample_table.restype = None
    L.mvs_pnp_sample_table.restype = None
    if L.mvs_abi_version() != 3:
        raise ImportError("libmvslam_b200.so ABI version mismatch")
    _lib = L
    return L


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))          # raw address (e.g. torch pinned tensor .data_ptr())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Comm:
    """mvs_comm: the NCCL communicator of the C-ABI multi-GPU driver (mvs_pair_batch_sharded).  The 128-byte id is made on
    one rank (Comm.unique_id()) and handed to the others by any means (a file, torch.distributed, MPI ...)."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        st = load_library().mvs_comm_unique_id(buf)
        if st != OK:
            raise MvsError(st, "mvs_comm_unique_id (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def __init__(self, ctx, uid, rank, world):
        self._L = load_library()
        self._ctx = ctx
        self._h = C.c_void_p()
        self.rank, self.world = rank, world
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        ctx._check(self._L.mvs_comm_create(C.byref(self._h), ctx._h, buf, int(rank), int(world)))

    def close(self):
        if self._h:
            self._L.mvs_comm_destroy(self._h)
            self._h = C.c_void_p()

    def pair_batch_sharded(self, pairs, K, ratio=0.7, max_dist=-1.0, cross_check=False, H=1, seed=0, mode=SCORE_ALGEBRAIC,
                           max_error_sq=0.0, solver=None, root=0, clouds=True, capacity_per_pair=None, results_out=None):
        """Every rank passes the full pair list.  Returns on root (results, dict(point_offsets, points, indexes, match_offsets,
        matches)) and (None, None) elsewhere.  results_out: a caller-owned array of n RESULT_DTYPE records (e.g. a view of pinned
        memory) for the root's records instead of a fresh pageable one."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2); n = pairs.shape[0]
        mp = MatchParams(ratio, max_dist, int(cross_check), 0)
        rp = RansacParams(H, mode, max_error_sq, seed, 0, _solver(solver), 0)
        is_root = self.rank == root
        res = (results_out if results_out is not None else np.zeros(n, RESULT_DTYPE)) if is_root else None
        cap = int(capacity_per_pair if capacity_per_pair is not None else self._ctx._frame_counts.max())
        po = np.zeros(n + 1, np.int64) if clouds else None
        mo = np.zeros(n + 1, np.int64) if clouds else None
        pts = idx = mat = None
        if clouds and is_root:
            pts = np.zeros((n * cap, 3)); idx = np.zeros(n * cap, np.uint64); mat = np.zeros(n * cap, MATCH_DTYPE)
        st = self._L.mvs_pair_batch_sharded(self._ctx._h, self._h, _p(pairs), C.c_int64(n), _p(_f64(K)), C.byref(mp), C.byref(rp),
                                            int(root), _p(res), _p(po), _p(pts), _p(idx), C.c_int64(n * cap if is_root else 0),
                                            _p(mo), _p(mat), C.c_int64(n * cap if is_root else 0))
        self._ctx._check(st)
        if not is_root:
            return None, None
        det = None
        if clouds:
            det = dict(point_offsets=po, points=pts[:po[-1]], indexes=idx[:po[-1]], match_offsets=mo, matches=mat[:mo[-1]])
        return res, det


def pnp_sample_table(seed, problem_id, n_points, H):
    """Host copy of the device 4-point sampler (mvs_pnp_sample_table)."""
    out = np.empty((H, 4), np.uint32)
    load_library().mvs_pnp_sample_table(C.c_uint64(seed), C.c_uint64(problem_id), C.c_uint32(n_points), int(H), _p(out))
    return out


def sample_table(seed, pair_id, n_points, H):
    """Host copy of the device sampler (mvs_sample_table)."""
    out = np.empty((H, 8), np.uint32)
    load_library().mvs_sample_table(C.c_uint64(seed), C.c_uint64(pair_id), C.c_uint32(n_points), int(H), _p(out))
    return out


_LIVE = __import__("weakref").WeakSet()   # open contexts (tests/conftest.py checks their guard bands in MVS_GUARD=1 runs)


class Context:
    """One context per (thread, CUDA device): owns the stream and the HBM workspace."""

    def __init__(self, device=0, stream=None):
        self._L = load_library()
        h = C.c_void_p()
        st = self._L.mvs_create(C.byref(h), int(device))
        if st != OK:
            raise MvsError(st, "mvs_create failed (no CUDA device / not sm_100?): "
                           + self._L.mvs_status_string(st).decode())
        self._h = h
        self.device = device
        self._frame_counts = None
        _LIVE.add(self)
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._L.mvs_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st, ok=(OK,)):
        if st not in ok:
            raise MvsError(st, (self._L.mvs_last_error(self._h) or b"").decode()
                           + " [" + self._L.mvs_status_string(st).decode() + "]")
        return st

    # ---- plumbing
    def set_stream(self, cuda_stream):
        self._check(self._L.mvs_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        self._check(self._L.mvs_synchronize(self._h))

    def profile_enable(self, on=True):
        self._check(self._L.mvs_profile_enable(self._h, int(bool(on))))

    def profile_read(self, reset=True):
        pr = Profile()
        self._check(self._L.mvs_profile_read(self._h, C.byref(pr), int(bool(reset))))
        return {s: (pr.ms[i], int(pr.launches[i])) for i, s in enumerate(STAGES)}

    def kernel_launches(self):
        return int(self._L.mvs_kernel_launches(self._h))

    def debug_guard_check(self):
        """MVS_GUARD=1 runs only: number of workspace buffers written past their end (-1: guard mode is off)."""
        return int(self._L.mvs_debug_guard_check(self._h))

    def debug_guard_poke(self):
        return int(self._L.mvs_debug_guard_poke(self._h))

    # ---- feature extraction
    def orb_extract(self, images, n_features=500, append_frames=False, want=True, device_ptr=None, shape=None, out=None):
        """VisualFeature::extract for a list/array of equal-size 8-bit grayscale images.

        Returns (counts int32[n], keypoints KEYPOINT_DTYPE[total], descriptors uint8[total][32], first_frame).
        device_ptr/shape=(n, h, stride, w): the images already live in device memory (contiguous).
        out=dict(kp=addr, desc=addr, capacity=int): preallocated (e.g. pinned) result buffers; returns counts only."""
        op = OrbParams(int(n_features), (C.c_int32 * 3)(0, 0, 0))
        first = C.c_int32(-1)
        if device_ptr is None:
            imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
            n = len(imgs); h, w = imgs[0].shape; stride = w
            assert all(im.shape == (h, w) for im in imgs), "images must share one size"
            ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        else:
            n, h, stride, w = shape
        counts = np.zeros(n, np.int32)

        def call(kp, desc, cap, append):
            if device_ptr is None:
                return self._L.mvs_orb_extract(self._h, ptrs, n, w, h, stride, C.byref(op), int(append), C.byref(first),
                                               _p(counts), _p(kp), _p(desc), C.c_int64(cap))
            return self._L.mvs_orb_extract_device(self._h, C.c_void_p(int(device_ptr)), n, w, h, stride, C.byref(op),
                                                  int(append), C.byref(first), _p(counts), _p(kp), _p(desc), C.c_int64(cap))
        if out is not None:
            self._check(call(out["kp"], out["desc"], int(out["capacity"]), append_frames))
            kp = desc = None
        elif not want:
            self._check(call(None, None, 0, append_frames))
            kp = desc = None
        else:
            cap = n * (int(n_features) + 64)
            kp = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
            st = call(kp, desc, cap, append_frames)
            if st == E_CAPACITY and counts.sum() > cap:      # ties at a cut-off: retry with the exact size
                if append_frames:
                    raise MvsError(st, "orb_extract: capacity (frames were appended; fetch sizes from counts)")
                cap = int(counts.sum())
                kp = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
                st = call(kp, desc, cap, False)
            self._check(st)
            tot = int(counts.sum())
            kp, desc = kp[:tot], desc[:tot]
        if append_frames:
            prev = self._frame_counts if self._frame_counts is not None else np.zeros(0, np.int32)
            self._frame_counts = np.concatenate([prev, counts])
        return counts, kp, desc, first.value

    # ---- matching
    def knn2_hamming(self, query, train):
        q = np.ascontiguousarray(query, np.uint8); t = np.ascontiguousarray(train, np.uint8)
        idx = np.empty((q.shape[0], 2), np.int32); dist = np.empty((q.shape[0], 2), np.int32)
        self._check(self._L.mvs_knn2_hamming(self._h, _p(q), q.shape[0], _p(t), t.shape[0],
                                             q.shape[1] if q.ndim == 2 else 0, _p(idx), _p(dist)))
        return idx, dist

    def match_hamming(self, query, train, ratio=0.7, max_dist=-1.0, cross_check=False, bounded=False):
        q = np.ascontiguousarray(query, np.uint8); t = np.ascontiguousarray(train, np.uint8)
        out = np.empty(max(q.shape[0], 1), MATCH_DTYPE); n = C.c_int(0)
        mp = MatchParams(ratio, max_dist, int(cross_check), int(bounded))
        self._check(self._L.mvs_match_hamming(self._h, _p(q), q.shape[0], _p(t), t.shape[0],
                                              q.shape[1] if q.ndim == 2 else 0, C.byref(mp), _p(out),
                                              out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    def knn2_l2(self, query, train):
        q = np.ascontiguousarray(query, np.float32); t = np.ascontiguousarray(train, np.float32)
        idx = np.empty((q.shape[0], 2), np.int32); dist = np.empty((q.shape[0], 2), np.float32)
        self._check(self._L.mvs_knn2_l2(self._h, _p(q), q.shape[0], _p(t), t.shape[0], q.shape[1], _p(idx), _p(dist)))
        return idx, dist

    def match_l2(self, query, train, ratio=0.7, max_dist=-1.0, cross_check=False):
        q = np.ascontiguousarray(query, np.float32); t = np.ascontiguousarray(train, np.float32)
        out = np.empty(max(q.shape[0], 1), MATCH_DTYPE); n = C.c_int(0)
        mp = MatchParams(ratio, max_dist, int(cross_check), 0)
        self._check(self._L.mvs_match_l2(self._h, _p(q), q.shape[0], _p(t), t.shape[0], q.shape[1], C.byref(mp),
                                         _p(out), out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    def knn2_l2_ptr(self, q_ptr, nq, t_ptr, nt, dim, idx_ptr, dist_ptr):
        """mvs_knn2_l2 on raw addresses (host or device, e.g. torch CUDA tensors' data_ptr())"""
        self._check(self._L.mvs_knn2_l2(self._h, _p(q_ptr), int(nq), _p(t_ptr), int(nt), int(dim), _p(idx_ptr), _p(dist_ptr)))

    def l2_stats(self):
        out = (C.c_uint64 * 4)()
        self._check(self._L.mvs_l2_stats(self._h, out))
        return dict(fallback_fwd=int(out[0]), fallback_rev=int(out[1]), gemm_us=int(out[2]), total_us=int(out[3]))

    # ---- geometry
    def find_fundamental_matrix(self, p1s, p2s, solver=None):
        p1s = _f64(p1s).reshape(-1, 8, 3); p2s = _f64(p2s).reshape(-1, 8, 3)
        F = np.empty((p1s.shape[0], 3, 3))
        self._check(self._L.mvs_find_fundamental_matrix(self._h, _p(p1s), _p(p2s), p1s.shape[0], _solver(solver), _p(F)))
        return F

    def svd_batch(self, A, solver=None):
        A = _f64(A); cnt, n = A.shape[0], A.shape[1]
        U = np.empty_like(A); Vt = np.empty_like(A); w = np.empty((cnt, n))
        self._check(self._L.mvs_svd_batch(self._h, n, _p(A), cnt, _solver(solver), _p(U), _p(w), _p(Vt)))
        return U, w, Vt

    def ransac_fundamental(self, p1, p2, samples=None, H=1, max_error_sq=1e-3, mode=SCORE_ALGEBRAIC, seed=0,
                           want_all=False, solver=None):
        p1 = _f64(p1); p2 = _f64(p2); n = p1.shape[0]
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.uint32); H = samples.shape[0]
        rp = RansacParams(H, mode, max_error_sq, seed, 0, _solver(solver), 0)
        F = np.zeros((3, 3)); mask = np.zeros(max(n, 1), np.uint8)
        cnt = C.c_int(); res = C.c_double(); bh = C.c_int(-1)
        allc = np.zeros(H, np.int32) if want_all else None
        st = self._L.mvs_ransac_fundamental(self._h, _p(p1), _p(p2), n, _p(samples), C.byref(rp), _p(F), _p(mask),
                                            C.byref(cnt), C.byref(res), C.byref(bh), _p(allc))
        self._check(st, (OK, E_TOO_FEW_POINTS, E_NO_MODEL))
        out = dict(status=st, F=F, mask=mask[:n], count=cnt.value, residual=res.value, best_h=bh.value)
        if want_all:
            out["all_counts"] = allc
        return out

    def sfm_solve(self, xy1, xy2, K, samples=None, H=1, seed=0, pair_id=0, mode=SCORE_ALGEBRAIC, max_error_sq=0.0,
                  solver=None):
        xy1 = _f64(xy1); xy2 = _f64(xy2); n = xy1.shape[0]
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.uint32); H = samples.shape[0]
        rp = RansacParams(H, mode, max_error_sq, seed, 0, _solver(solver), pair_id)
        res = np.zeros(1, RESULT_DTYPE); mask = np.zeros(max(n, 1), np.uint8)
        pts = np.empty((max(n, 1), 3)); idx = np.empty(max(n, 1), np.uint64)
        st = self._L.mvs_sfm_solve(self._h, _p(xy1), _p(xy2), n, _p(_f64(K)), C.byref(rp), _p(samples), _p(res),
                                   _p(mask), _p(pts), _p(idx), max(n, 1))
        self._check(st, (OK, E_TOO_FEW_POINTS, E_NO_MODEL, E_TOO_FEW_INLIERS, E_NO_CHEIRALITY))
        r = res[0]
        d = {k: (r[k].copy() if r[k].ndim else r[k].item()) for k in RESULT_DTYPE.names}
        m = d["n_points"] if st == OK else 0
        d["mask"] = mask[:n]; d["points"] = pts[:m].copy(); d["indexes"] = idx[:m].copy()
        return d

    def sfm_triangulate(self, xy1, xy2, K, R1, t1, R2, t2, solver=None):
        xy1 = _f64(xy1); xy2 = _f64(xy2); n = xy1.shape[0]
        pts = np.empty((max(n, 1), 3)); idx = np.empty(max(n, 1), np.uint64); m = C.c_int(0)
        self._check(self._L.mvs_sfm_triangulate(self._h, _p(xy1), _p(xy2), n, _p(_f64(K)), _p(_f64(R1)), _p(_f64(t1)),
                                                _p(_f64(R2)), _p(_f64(t2)), _solver(solver), _p(pts), _p(idx), max(n, 1), C.byref(m)))
        return pts[:m.value].copy(), idx[:m.value].copy()

    # ---- pnp_solve
    def pnp_solve(self, world, image, K, samples=None, H=100, seed=0, problem_id=0, reproj_error=0.05, refine_iters=10,
                  want_all=False):
        w = _f64(world).reshape(-1, 3); im = _f64(image).reshape(-1, 2); n = w.shape[0]
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.uint32); H = samples.shape[0]
        pp = PnpParams(H, refine_iters, reproj_error, seed, problem_id, 0, 0)
        res = np.zeros(1, PNP_RESULT_DTYPE); mask = np.zeros(max(n, 1), np.uint8)
        allc = np.zeros(H, np.int32) if want_all else None
        st = self._L.mvs_pnp_solve(self._h, _p(w), _p(im), n, _p(_f64(K)), C.byref(pp), _p(samples), _p(res), _p(mask), _p(allc))
        self._check(st, (OK, E_TOO_FEW_POINTS, E_NO_MODEL))
        r = res[0]
        d = {k: (r[k].copy() if r[k].ndim else r[k].item()) for k in PNP_RESULT_DTYPE.names}
        d["mask"] = mask[:n]
        if want_all:
            d["all_counts"] = allc
        return d

    def pnp_solve_batch(self, worlds, images, K, H=100, seed=0, problem_id_base=0, reproj_error=0.05, refine_iters=10, counts=None):
        """worlds/images: lists of per-problem arrays, or (with counts) the already concatenated [total,3] / [total,2]
        arrays the C ABI takes.  Returns (results[PNP_RESULT_DTYPE], masks list)."""
        if counts is None:
            counts = np.array([len(w) for w in worlds], np.int32)
            w = _f64(np.concatenate([np.asarray(x, np.float64).reshape(-1, 3) for x in worlds])) if counts.sum() else np.zeros((0, 3))
            im = _f64(np.concatenate([np.asarray(x, np.float64).reshape(-1, 2) for x in images])) if counts.sum() else np.zeros((0, 2))
        else:
            counts = np.ascontiguousarray(counts, np.int32); w = _f64(worlds); im = _f64(images)
        pp = PnpParams(H, refine_iters, reproj_error, seed, problem_id_base, 0, 0)
        res = np.zeros(len(counts), PNP_RESULT_DTYPE); mask = np.zeros(max(int(counts.sum()), 1), np.uint8)
        self._check(self._L.mvs_pnp_solve_batch(self._h, _p(w), _p(im), _p(counts), len(counts), _p(_f64(K)), C.byref(pp), None,
                                                _p(res), _p(mask)))
        offs = np.concatenate([[0], np.cumsum(counts)])
        return res, [mask[offs[i]:offs[i + 1]] for i in range(len(counts))]

    # ---- bundle adjustment
    @staticmethod
    def ba_pack(problems):
        """List of problem dicts -> the concatenated arrays of the C ABI (see ba_solve_batch)."""
        cat = lambda k, shape: _f64(np.concatenate([np.asarray(p[k], np.float64).reshape(shape) for p in problems]))  # noqa: E731
        return dict(nf=np.array([len(p["pose_R"]) for p in problems], np.int32),
                    npt=np.array([len(p["points"]) for p in problems], np.int32),
                    no=np.array([len(p["obs"]) for p in problems], np.int32),
                    R=cat("pose_R", (-1, 9)), t=cat("pose_t", (-1, 3)), pc=cat("pose_prior_cov", (-1, 36)),
                    X=cat("points", (-1, 3)), xc=cat("point_prior_cov", (-1, 9)),
                    obs=np.ascontiguousarray(np.concatenate([np.asarray(p["obs"], BA_OBS_DTYPE) for p in problems])))

    def ba_solve_packed(self, K, pk, max_iterations=100, lambda_initial=1e-5, relative_tolerance=1e-13, absolute_tolerance=-1.0):
        """One mvs_ba_solve_batch call on packed arrays; returns (results, R, t, pose_cov, points, point_cov) concatenated."""
        Ro = np.empty_like(pk["R"]); to = np.empty_like(pk["t"]); pco = np.empty_like(pk["pc"])
        Xo = np.empty_like(pk["X"]); xco = np.empty_like(pk["xc"])
        res = np.zeros(len(pk["nf"]), BA_RESULT_DTYPE)
        bp = BaParams(max_iterations, 0, lambda_initial, relative_tolerance, absolute_tolerance)
        self._check(self._L.mvs_ba_solve_batch(self._h, len(pk["nf"]), _p(_f64(K)), _p(pk["nf"]), _p(pk["npt"]), _p(pk["no"]),
                                               _p(pk["R"]), _p(pk["t"]), _p(pk["pc"]), _p(pk["X"]), _p(pk["xc"]), _p(pk["obs"]),
                                               C.byref(bp), _p(Ro), _p(to), _p(pco), _p(Xo), _p(xco), _p(res)))
        return res, Ro, to, pco, Xo, xco

    def ba_solve_batch(self, K, problems, **kw):
        """problems: list of dict(pose_R [F,3,3], pose_t [F,3], pose_prior_cov [F,6,6] (NaN rows = none), points [P,3],
        point_prior_cov [P,3,3] (NaN = none), obs BA_OBS_DTYPE[O]).  Returns a list of result dicts."""
        pk = self.ba_pack(problems)
        res, Ro, to, pco, Xo, xco = self.ba_solve_packed(K, pk, **kw)
        out = []
        fo = np.concatenate([[0], np.cumsum(pk["nf"])]); po = np.concatenate([[0], np.cumsum(pk["npt"])])
        for i in range(len(problems)):
            f = slice(fo[i], fo[i + 1]); q = slice(po[i], po[i + 1])
            out.append(dict(status=int(res["status"][i]), iterations=int(res["iterations"][i]),
                            initial_error=float(res["initial_error"][i]), final_error=float(res["final_error"][i]),
                            pose_R=Ro[f].reshape(-1, 3, 3), pose_t=to[f], pose_cov=pco[f].reshape(-1, 6, 6),
                            points=Xo[q], point_cov=xco[q].reshape(-1, 3, 3)))
        return out

    # ---- batched pairs
    def frames_upload(self, descs, kps):
        descs = [np.ascontiguousarray(d, np.uint8) for d in descs]
        kps = [np.ascontiguousarray(k, np.float32) for k in kps]
        nf = len(descs)
        dptr = (C.c_void_p * nf)(*[d.ctypes.data for d in descs])
        kptr = (C.c_void_p * nf)(*[k.ctypes.data for k in kps])
        counts = np.array([d.shape[0] for d in descs], np.int32)
        self._check(self._L.mvs_frames_upload(self._h, nf, dptr, kptr, _p(counts), 32))
        self._frame_counts = counts

    def frames_upload_packed(self, desc_all, kp_all, counts):
        """desc_all / kp_all: addresses (e.g. of pinned torch tensors) or contiguous arrays; does not synchronise."""
        counts = np.ascontiguousarray(counts, np.int32)
        self._check(self._L.mvs_frames_upload_packed(self._h, len(counts), _p(desc_all), _p(kp_all), _p(counts), 32))
        self._frame_counts = counts

    def frames_clear(self):
        self._check(self._L.mvs_frames_clear(self._h))
        self._frame_counts = np.zeros(0, np.int32)

    def frames_append(self, desc, kp):
        desc = np.ascontiguousarray(desc, np.uint8); kp = np.ascontiguousarray(kp, np.float32)
        idx = C.c_int32(-1)
        self._check(self._L.mvs_frames_append(self._h, _p(desc), _p(kp), desc.shape[0], 32, C.byref(idx)))
        prev = self._frame_counts if self._frame_counts is not None else np.zeros(0, np.int32)
        self._frame_counts = np.concatenate([prev, np.array([desc.shape[0]], np.int32)])
        return idx.value

    def pair_batch(self, pairs, K, ratio=0.7, max_dist=-1.0, cross_check=False, H=1, seed=0, mode=SCORE_ALGEBRAIC,
                   max_error_sq=0.0, pair_id_base=0, details=True, out=None, enqueue_only=False, bounded=False, solver=None):
        """Returns (results[RESULT_DTYPE], details dict or None).  `out` may hold preallocated (e.g. pinned)
        buffers: dict(results=addr/array, matches=, mask=, points=, indexes=, capacity=int)."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2); npairs = pairs.shape[0]
        mp = MatchParams(ratio, max_dist, int(cross_check), int(bounded))
        rp = RansacParams(H, mode, max_error_sq, seed, 0, _solver(solver), pair_id_base)
        fn = self._L.mvs_pair_batch_enqueue if enqueue_only else self._L.mvs_pair_batch
        if out is not None:
            st = fn(self._h, _p(pairs), npairs, _p(_f64(K)), C.byref(mp), C.byref(rp), _p(out["results"]),
                    _p(out.get("matches")), _p(out.get("mask")), _p(out.get("points")), _p(out.get("indexes")),
                    int(out.get("capacity", 0)))
            self._check(st)
            return None, None
        res = np.zeros(npairs, RESULT_DTYPE)
        det = None
        if details:
            cap = int(self._frame_counts.max())
            det = dict(matches=np.zeros((npairs, cap), MATCH_DTYPE), mask=np.zeros((npairs, cap), np.uint8),
                       points=np.zeros((npairs, cap, 3)), indexes=np.zeros((npairs, cap), np.uint64), capacity=cap)
        st = fn(self._h, _p(pairs), npairs, _p(_f64(K)), C.byref(mp), C.byref(rp), _p(res),
                _p(det["matches"]) if det else None, _p(det["mask"]) if det else None,
                _p(det["points"]) if det else None, _p(det["indexes"]) if det else None, det["capacity"] if det else 0)
        self._check(st)
        return res, det
