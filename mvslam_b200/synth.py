"""Seeded synthetic workloads of the named BASELINE shapes.  Input generators only —
no part of the hot path.  numpy on the host; used by tests/ and bench.py."""
import numpy as np

K_S8K = np.array([[700.0, 0, 640.0], [0, 700.0, 360.0], [0, 0, 1.0]])


def _rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def _project(K, R, t, X):
    pc = X @ R.T + t
    return np.stack([K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 2], K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2]], 1), pc[:, 2]


def _flip_bits(desc, p, rng):
    flips = np.packbits(rng.random((desc.shape[0], desc.shape[1] * 8)) < p, axis=1)
    return desc ^ flips


def synthetic_pair(pair_id, n=8192, noise_px=0.5, flip_p=0.05, K=K_S8K, size=(1280, 720), seed_base=0x5EED0000):
    """Config 3 (S8k): one frame pair with n keypoints per image.  Returns
    (desc1[n,32] u8, kp1[n,2] f32, desc2, kp2, truth dict)."""
    rng = np.random.Generator(np.random.PCG64(seed_base + pair_id))
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.uniform(4, 12, n)], 1)
    rv = rng.normal(size=3); rv *= 0.1 * rng.uniform() ** (1 / 3) / np.linalg.norm(rv)
    R = _rodrigues(rv)
    t = rng.normal(size=3); t *= 0.5 / np.linalg.norm(t)
    x1, z1 = _project(K, np.eye(3), np.zeros(3), X)
    x2, z2 = _project(K, R, t, X)
    x1 = x1 + rng.normal(size=x1.shape) * noise_px
    x2 = x2 + rng.normal(size=x2.shape) * noise_px
    w, h = size
    inside = ((x1[:, 0] >= 0) & (x1[:, 0] < w) & (x1[:, 1] >= 0) & (x1[:, 1] < h) & (z1 > 0) &
              (x2[:, 0] >= 0) & (x2[:, 0] < w) & (x2[:, 1] >= 0) & (x2[:, 1] < h) & (z2 > 0))
    out = ~inside
    no = int(out.sum())
    x1[out] = rng.uniform(0, 1, (no, 2)) * [w, h]
    x2[out] = rng.uniform(0, 1, (no, 2)) * [w, h]
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    d2 = _flip_bits(d1, flip_p, rng)
    d2[out] = rng.integers(0, 256, (no, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    truth = dict(R=R, t=t, inlier=inside, perm=perm, X=X)
    return (d1, x1.astype(np.float32), np.ascontiguousarray(d2[perm]), np.ascontiguousarray(x2[perm]).astype(np.float32),
            truth)


def synthetic_window(n_frames=512, n_kp=2048, n_scene=20000, noise_px=0.5, flip_p=0.05, K=K_S8K,
                     size=(1280, 720), seed=0x512):
    """Config 5 (W512): frames rendered from one scene along a smooth seeded trajectory.
    Returns (descs list, kps list)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = np.stack([rng.uniform(-12, 12, n_scene), rng.uniform(-6, 6, n_scene), rng.uniform(4, 16, n_scene)], 1)
    D = rng.integers(0, 256, (n_scene, 32), dtype=np.uint8)
    w, h = size
    descs, kps = [], []
    for f in range(n_frames):
        s = f / max(n_frames - 1, 1)
        rv = np.array([0.05 * np.sin(2 * np.pi * s), 0.3 * (s - 0.5), 0.03 * np.cos(2 * np.pi * s)])
        c = np.array([10.0 * (s - 0.5), 0.5 * np.sin(4 * np.pi * s), 1.0 * s])
        R = _rodrigues(rv)
        x, z = _project(K, R, -R @ c, X)
        vis = np.nonzero((x[:, 0] >= 0) & (x[:, 0] < w) & (x[:, 1] >= 0) & (x[:, 1] < h) & (z > 0.5))[0]
        if len(vis) > n_kp:
            vis = rng.choice(vis, n_kp, replace=False)
        nv = len(vis)
        kp = np.empty((n_kp, 2)); d = np.empty((n_kp, 32), np.uint8)
        kp[:nv] = x[vis] + rng.normal(size=(nv, 2)) * noise_px
        d[:nv] = _flip_bits(D[vis], flip_p, rng)
        kp[nv:] = rng.uniform(0, 1, (n_kp - nv, 2)) * [w, h]
        d[nv:] = rng.integers(0, 256, (n_kp - nv, 32), dtype=np.uint8)
        perm = rng.permutation(n_kp)
        descs.append(np.ascontiguousarray(d[perm])); kps.append(np.ascontiguousarray(kp[perm]).astype(np.float32))
    return descs, kps


def synthetic_l2(nq=32768, nt=32768, dim=64, seed=1234):
    """Config 4 (L2): unit-norm float descriptors; half of the train rows are noisy copies of queries."""
    r1 = np.random.Generator(np.random.PCG64(seed)); r2 = np.random.Generator(np.random.PCG64(seed + 1))
    Q = r1.normal(size=(nq, dim)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    T = r2.normal(size=(nt, dim)).astype(np.float32)
    T /= np.linalg.norm(T, axis=1, keepdims=True)
    m = min(nq, nt) // 2
    perm = r2.permutation(nt)[:m]
    C = Q[:m] + 0.05 * r2.normal(size=(m, dim)).astype(np.float32)
    T[perm] = C / np.linalg.norm(C, axis=1, keepdims=True)
    return np.ascontiguousarray(Q), np.ascontiguousarray(T)


def synthetic_image(seed, width=640, height=480, n_shapes=220):
    """Seeded 8-bit grayscale test image with plenty of corners at several scales: smooth background, overlapping
    rectangles and triangles of random intensity, mild noise.  Input generator for the feature-extraction tests/bench."""
    rng = np.random.Generator(np.random.PCG64(0x0B5E0000 + seed))
    coarse = rng.random((height // 32 + 2, width // 32 + 2)) * 120 + 40
    yy = np.arange(height) / 32.0
    xx = np.arange(width) / 32.0
    y0 = yy.astype(int); x0 = xx.astype(int)
    fy = (yy - y0)[:, None]; fx = (xx - x0)[None, :]
    img = ((1 - fy) * (1 - fx) * coarse[y0][:, x0] + (1 - fy) * fx * coarse[y0][:, x0 + 1]
           + fy * (1 - fx) * coarse[y0 + 1][:, x0] + fy * fx * coarse[y0 + 1][:, x0 + 1])
    Y, X = np.mgrid[0:height, 0:width]
    for _ in range(n_shapes):
        cx, cy = rng.integers(0, width), rng.integers(0, height)
        sz = int(rng.integers(4, 70))
        val = float(rng.integers(0, 256))
        if rng.random() < 0.6:
            a = rng.random() * np.pi
            u = (X - cx) * np.cos(a) + (Y - cy) * np.sin(a)
            v = -(X - cx) * np.sin(a) + (Y - cy) * np.cos(a)
            m = (np.abs(u) < sz) & (np.abs(v) < sz * (0.3 + 0.7 * rng.random()))
        else:
            m = (X - cx >= 0) & (Y - cy >= 0) & ((X - cx) + (Y - cy) * (0.5 + rng.random()) < sz)
        img = np.where(m, 0.75 * val + 0.25 * img, img)
    img = img + rng.normal(0, 2.0, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
