"""Oracle A — numpy + cv2 restatement of mvSLAM's own two-view branch.

TEST INFRASTRUCTURE ONLY (never imported by the product package).

It calls the *same third-party routines* the reference calls — cv2.BFMatcher.knnMatch /
cv2.batchDistance (source/vision/visual-feature.cpp:59-62) and cv2.SVDecomp
(source/math/svd.hpp:65, source/vision/fundamental-matrix.cpp:115,131) — and restates the
reference's Eigen arithmetic around them.  Everything UPSTREAM of an SVD (K^-1, Hartley
normalisation, the rows of A, A^T A, the residuals) is written in the reference's own loop order
with one rounding per operation (numpy elementwise ufuncs / Python floats; never BLAS, whose FMA
kernels round differently): on an ill-conditioned 8-point sample a single different last bit in
A^T A moves the null vector by ~eps*cond(A)^2 (4e-5 on Tsukuba pair 4-5), so only an exact
restatement pins the reference there.  Its job is to pin Oracle B (oracle/mvs_oracle.c, whose
REFERENCE solver restates cv::SVDecomp itself bit for bit and is what the CUDA kernels are checked
against) and to generate the golden fixtures in tests/golden/.
File:line citations are relative to /root/reference.
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

EPSILON = np.finfo(np.float64).eps          # source/system-config.hpp:8
TOLERANCE = EPSILON * 1000                  # :10
INFINITY = np.finfo(np.float64).max / 10    # :14
MAX_ERROR_SQ = 5e-2                         # source/vision/sfm-solve.cpp:18
VF_MATCH_INLIER_MIN = 8                     # :20
RATIO = 0.7                                 # source/vision/visual-feature.cpp:24


# ---------------------------------------------------------------- matching
def knn2_hamming(query, train):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) in array form (visual-feature.cpp:59-62)."""
    dist, idx = cv2.batchDistance(query, train, cv2.CV_32S, normType=cv2.NORM_HAMMING, K=2)
    return idx.astype(np.int32), dist.astype(np.int32)


def knn2_l2(query, train):
    dist, idx = cv2.batchDistance(query, train, cv2.CV_32F, normType=cv2.NORM_L2, K=2)
    return idx.astype(np.int32), dist.astype(np.float32)


def filter_matches(idx, dist, max_dist=-1.0, ratio=RATIO):
    """Lowe ratio + max_dist + sort (visual-feature.cpp:64-78); canonical order (distance, queryIdx)."""
    d1 = dist[:, 0].astype(np.float32).astype(np.float64)
    d2 = dist[:, 1].astype(np.float32).astype(np.float64)
    keep = d1 < ratio * d2
    if max_dist >= 0:
        keep &= d1 <= max_dist
    q = np.nonzero(keep)[0]
    order = np.lexsort((q, d1[q]))
    q = q[order]
    return q.astype(np.int32), idx[q, 0].astype(np.int32), d1[q].astype(np.float32)


def match_visual_features(desc1, desc2, max_dist=-1.0, norm="hamming"):
    """VisualFeature::match_visual_features(vf1, vf2): query = vf2, train = vf1."""
    idx, dist = (knn2_hamming if norm == "hamming" else knn2_l2)(desc2, desc1)
    return filter_matches(idx, dist, max_dist)


# ---------------------------------------------------------------- algebra
def svd(A):
    """SVD<> wrapper (source/math/svd.hpp:59-72): returns U, w (descending), V (= vt^T)."""
    w, u, vt = cv2.SVDecomp(np.array(A, np.float64, order="C"), flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
    return u, w.ravel(), vt.T


def mat3_mul(A, B):
    """Eigen fixed-size coefficient-based product: sum over k left to right, no FMA (SConstruct:70,86)."""
    A = np.asarray(A, np.float64); B = np.asarray(B, np.float64)
    return (A[:, 0:1] * B[0] + A[:, 1:2] * B[1]) + A[:, 2:3] * B[2]


def mat3_vec(A, v):
    A = np.asarray(A, np.float64)
    return (A[:, 0] * v[0] + A[:, 1] * v[1]) + A[:, 2] * v[2]


def inverse3(K):
    """Eigen's fixed 3x3 inverse (camera.cpp:14-18 K_.inverse()): cofactor^T * (1/det)."""
    K = [[float(x) for x in r] for r in np.asarray(K, np.float64)]
    c00 = K[1][1] * K[2][2] - K[1][2] * K[2][1]; c01 = K[1][2] * K[2][0] - K[1][0] * K[2][2]
    c02 = K[1][0] * K[2][1] - K[1][1] * K[2][0]; c10 = K[0][2] * K[2][1] - K[0][1] * K[2][2]
    c11 = K[0][0] * K[2][2] - K[0][2] * K[2][0]; c12 = K[0][1] * K[2][0] - K[0][0] * K[2][1]
    c20 = K[0][1] * K[1][2] - K[0][2] * K[1][1]; c21 = K[0][2] * K[1][0] - K[0][0] * K[1][2]
    c22 = K[0][0] * K[1][1] - K[0][1] * K[1][0]
    idet = 1.0 / (K[0][0] * c00 + K[0][1] * c01 + K[0][2] * c02)
    return np.array([[c00 * idet, c10 * idet, c20 * idet], [c01 * idet, c11 * idet, c21 * idet],
                     [c02 * idet, c12 * idet, c22 * idet]])


def so3_rectify(R):
    """SO3::rectify (source/math/lie-group.hpp:84-96); row 1 is NOT normalised."""
    R = np.asarray(R, np.float64)
    u0 = R[0] / np.sqrt(R[0, 0] * R[0, 0] + R[0, 1] * R[0, 1] + R[0, 2] * R[0, 2])
    u1 = R[1] - (R[1, 0] * u0[0] + R[1, 1] * u0[1] + R[1, 2] * u0[2]) * u0
    u2 = np.cross(u0, u1)
    return np.stack([u0, u1, u2])


def se3_inverse(R, t):
    """SE3::inverse (lie-group.hpp:203-207): RT = SO3(R^T) (rectified again), t' = -(RT t)."""
    RT = so3_rectify(np.asarray(R).T)
    return RT, -mat3_vec(RT, t)


def se3_compose(Ra, ta, Rb, tb):
    """SE3::operator* (lie-group.hpp:220-225)."""
    return so3_rectify(mat3_mul(Ra, Rb)), mat3_vec(Ra, tb) + ta


def rodrigues(v):
    """source/math/lie-group.cpp:16-32"""
    theta = np.linalg.norm(v)
    if theta < EPSILON:
        A = 1.0 - theta ** 2 / 6.0; B = 0.5 - theta ** 2 / 24.0
    else:
        A = np.sin(theta) / theta; B = (1.0 - np.cos(theta)) / theta ** 2
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], float)
    return np.eye(3) + A * K + B * K @ K


def so3_from_rpy(roll, pitch, yaw):
    """SO3(roll,pitch,yaw) (lie-group.hpp:42-56)"""
    Rx = np.array([[1, 0, 0], [0, np.cos(roll), -np.sin(roll)], [0, np.sin(roll), np.cos(roll)]])
    Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def so3_ln(R):
    """SO3::ln (lie-group.hpp:139-161)"""
    c = min(max(0.5 * (np.trace(R) - 1.0), -1.0), 1.0)
    theta = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    A = (1.0 + theta ** 2 / 6.0) * 0.5 if theta < 1e-5 else 0.5 * theta / np.sin(theta)
    return v * A


def se3_ln(R, t):
    """SE3::ln (lie-group.hpp:236-262): translation part first."""
    w = so3_ln(R); theta = np.linalg.norm(w)
    if theta < 1e-5:
        G = 1.0 / 12.0 + theta ** 2 / 720.0
    else:
        A = np.sin(theta) / theta; B = (1.0 - np.cos(theta)) / theta ** 2
        G = (1.0 - 0.5 * A / B) / theta ** 2
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], float)
    return np.concatenate([(np.eye(3) - 0.5 * K + G * K @ K) @ t, w])


# ---------------------------------------------------------------- camera
def normalize_points(K, xy):
    """PinholeCamera::normalize_points (source/vision/camera.cpp:55-79): K^-1 (u, v, 1)."""
    Kinv = inverse3(K)
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    return np.stack([(Kinv[r, 0] * xy[:, 0] + Kinv[r, 1] * xy[:, 1]) + Kinv[r, 2] * 1.0 for r in range(3)], axis=1)


def project_points(K, R_w2c, t_w2c, pts):
    """PinholeCamera::project_points (camera.cpp:24-53) with extrinsics P = (R_w2c, t_w2c)."""
    pc = pts @ R_w2c.T + t_w2c
    pn = np.stack([pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2], np.ones(len(pc))], axis=1)
    return (pn @ K.T)[:, :2]


# ---------------------------------------------------------------- 8-point
def find_normalization_transform(p):
    """source/vision/fundamental-matrix.cpp:18-54 (mean distance -> sqrt(2))."""
    p = np.asarray(p, np.float64)
    inv = 1.0 / len(p)
    mean = np.zeros(3)
    for q in p:                       # :29-35, sequential accumulation
        mean = mean + q
    mean = mean * inv
    c = p - mean
    scale = 0.0
    for q in c:                       # :39-43, Eigen norm() = sqrt(x^2 + y^2 + z^2)
        scale = scale + float(np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]))
    scale = scale * inv
    scale = float(np.sqrt(2.0)) / scale
    T = np.array([[scale, 0, -mean[0] * scale], [0, scale, -mean[1] * scale], [0, 0, 1]])
    return c * scale, T


def find_fundamental_matrix_8point(n1, n2):
    """source/vision/fundamental-matrix.cpp:56-140"""
    x1, y1, x2, y2 = n1[:, 0], n1[:, 1], n2[:, 0], n2[:, 1]
    A = np.stack([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, np.ones(8)], axis=1)
    AtA = np.zeros((9, 9))
    for k in range(8):                # :104-111: AT_A(i,j) += A(k,i)*A(k,j), k outermost per element = same order
        AtA = AtA + np.outer(A[k], A[k])
    _, _, vt = cv2.SVDecomp(AtA, flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
    F = vt[8].reshape(3, 3).copy()
    w, u, vt3 = cv2.SVDecomp(F.copy(), flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
    w = w.ravel().copy(); w[2] = 0
    # u * diag(w) * vt (:134) through cv::gemm's small-matrix path: products summed left to right
    return mat3_mul(mat3_mul(u, np.diag(w)), vt3)


def find_fundamental_matrix(p1s, p2s):
    """source/vision/fundamental-matrix.cpp:204-267"""
    n1, T1 = find_normalization_transform(np.asarray(p1s, np.float64))
    n2, T2 = find_normalization_transform(np.asarray(p2s, np.float64))
    return mat3_mul(mat3_mul(T2.T, find_fundamental_matrix_8point(n1, n2)), T1)


# ---------------------------------------------------------------- RANSAC
def residuals(p1, p2, F, mode="algebraic"):
    p1 = np.asarray(p1, np.float64); p2 = np.asarray(p2, np.float64); F = np.asarray(F, np.float64)
    v = (p2[:, 0:1] * F[0] + p2[:, 1:2] * F[1]) + p2[:, 2:3] * F[2]      # rows: p2^T F, left to right
    r = (v[:, 0] * p1[:, 0] + v[:, 1] * p1[:, 1]) + v[:, 2] * p1[:, 2]
    if mode == "algebraic":
        return np.abs(r)             # estimator-RANSAC.cpp:114-116
    l = (p1[:, 0:1] * F[:, 0] + p1[:, 1:2] * F[:, 1]) + p1[:, 2:3] * F[:, 2]   # rows: F p1
    return r * r / ((l[:, 0] * l[:, 0] + l[:, 1] * l[:, 1]) + (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]))


def count_inliers(p1, p2, F, max_error_sq, mode="algebraic"):
    """source/vision/estimator-RANSAC.cpp:100-129"""
    r = residuals(p1, p2, F, mode)
    mask = r < max_error_sq
    res = 0.0
    for x in r[mask]:                # residual += r in point order (:120)
        res += float(x)
    return int(mask.sum()), res, mask.astype(np.uint8)


def ransac_fundamental(p1, p2, samples, max_error_sq, mode="algebraic"):
    """source/vision/estimator-RANSAC.cpp:16-90 with an explicit sample table (row 0 = reference)."""
    best = dict(count=0, residual=INFINITY, F=None, mask=None, best_h=-1)
    if len(p1) < 8:
        return best
    for h, row in enumerate(samples):
        F = find_fundamental_matrix(p1[row], p2[row])
        cnt, res, mask = count_inliers(p1, p2, F, max_error_sq, mode)
        if cnt > best["count"] or (cnt == best["count"] and res < best["residual"]):
            best = dict(count=cnt, residual=res, F=F, mask=mask, best_h=h)
    return best


# ---------------------------------------------------------------- essential / pose / triangulation
def project_essential(F):
    """source/vision/sfm-solve.cpp:73-87"""
    U, s, V = svd(F)
    v = np.sqrt(s[0] * s[1])
    return mat3_mul(mat3_mul(U, np.diag([v, v, 0.0])), V.T)


def det3(M):
    return (M[0, 0] * (M[1, 1] * M[2, 2] - M[1, 2] * M[2, 1]) - M[0, 1] * (M[1, 0] * M[2, 2] - M[1, 2] * M[2, 0])
            + M[0, 2] * (M[1, 0] * M[2, 1] - M[1, 1] * M[2, 0]))


def decompose_essential(E):
    """source/vision/sfm-solve.cpp:97-127"""
    U, _, V = svd(E)
    if det3(U) < 0:
        U = -U
    if det3(V) < 0:
        V = -V
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], float)
    Z = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 0]], float)
    Ra = mat3_mul(mat3_mul(U, W), V.T)
    Rb = mat3_mul(mat3_mul(U, W.T), V.T)
    S = mat3_mul(mat3_mul(U, Z), U.T)
    return Ra, Rb, np.array([-S[1, 2], S[0, 2], -S[0, 1]])


def triangulate_points(R, t, p1, p2, mask=None):
    """source/vision/sfm-solve.cpp:134-227"""
    P1 = np.eye(4)
    P2 = np.eye(4); P2[:3, :3] = so3_rectify(R); P2[:3, 3] = t
    pts, idx = [], []
    for i in range(len(p1)):
        if mask is not None and mask[i] == 0:
            continue
        x1, x2 = p1[i], p2[i]
        A = np.stack([x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1],
                      x2[0] * P2[2] - P2[0], x2[1] * P2[2] - P2[1]])
        _, _, V = svd(A)
        X = V[:, 3]
        if abs(X[3]) < TOLERANCE:
            continue
        pt = X[:3] * (1.0 / X[3])
        if pt[2] < TOLERANCE:
            continue
        if ((R[2, 0] * pt[0] + R[2, 1] * pt[1]) + R[2, 2] * pt[2]) + t[2] < TOLERANCE:
            continue
        pts.append(pt); idx.append(i)
    return np.array(pts).reshape(-1, 3), np.array(idx, np.uint64)


def recover_pose_and_points(E, p1, p2, mask):
    """source/vision/sfm-solve.cpp:232-284"""
    Ra, Rb, t = decompose_essential(E)
    best = None
    for ci, (R, tt) in enumerate([(Ra, t), (Ra, -t), (Rb, t), (Rb, -t)]):
        pts, idx = triangulate_points(R, tt, p1, p2, mask)
        if len(idx) > (0 if best is None else len(best["indexes"])):
            best = dict(R=R, t=tt, points=pts, indexes=idx, candidate=ci)
    return best


def sfm_solve(xy1, xy2, K, samples=None, mode="algebraic"):
    """source/vision/sfm-solve.cpp:285-368 (own branch of find_essential_matrix, :64-90)."""
    p1 = normalize_points(K, xy1); p2 = normalize_points(K, xy2)
    if samples is None:
        samples = np.arange(8, dtype=np.uint32)[None, :]      # the reference's single sample
    max_error_sq = MAX_ERROR_SQ / K[0, 0] / K[1, 1]
    if len(p1) < 8:
        return None
    r = ransac_fundamental(p1, p2, samples, max_error_sq, mode)
    if r["count"] <= 0:
        return None
    E = project_essential(r["F"])
    if r["count"] < VF_MATCH_INLIER_MIN:
        return None
    rec = recover_pose_and_points(E, p1, p2, r["mask"])
    if rec is None:
        return None
    R2in1, t2in1 = se3_inverse(so3_rectify(rec["R"]), rec["t"])
    return dict(F=r["F"], E=E, mask=r["mask"], n_inliers=r["count"], residual=r["residual"],
                best_h=r["best_h"], R1to2=rec["R"], t1to2=rec["t"], R2in1=R2in1, t2in1=t2in1,
                points=rec["points"], indexes=rec["indexes"], candidate=rec["candidate"])


def sfm_triangulate(xy1, xy2, K, R1, t1, R2, t2):
    """source/vision/sfm-solve.cpp:370-394"""
    Ri, ti = se3_inverse(R2, t2)
    R12, t12 = se3_compose(Ri, ti, R1, t1)
    return triangulate_points(R12, t12, normalize_points(K, xy1), normalize_points(K, xy2), None)


def get_rig_points(kind, R, t, scale):
    """test/unit-test-helper.cpp:42-79"""
    if kind == "cube":
        p = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], float)
    else:
        p = np.array([[1, 0, 0], [0, 0, 0], [0, 2, 0], [1, 0, 3], [0, 0, 3], [0, 2, 3],
                      [0.5, 0, 1.5], [0, 1, 1.5]], float)
    return (scale * p) @ R.T + t
