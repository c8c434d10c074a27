"""Host-side model of the arithmetic of the tensor-core matcher (mvslam_b200/csrc/match_hamming_tc.cu), checked against the
CPU oracle without a GPU: the +-8 byte encoding, the signed 16-bit key `128 (128 - hamming) + (63 - c')` (c' = column
inside a thread's 64-column half of a 128-column tile), its widening to `hamming << 22 | trainIdx`, and the split of the
top-2 search into "maximum of four column streams per 64-column block" (epilogue) + "the best's 15 stream-mates" (K2's
refine_second_warp).  The CUDA kernels are tested against the same oracle in tests/test_gpu_parity.py; this file pins the
algebra they rely on."""
import numpy as np
import pytest

from oracle import cbind as orc

BN = 128
HALF = 64


def expand(desc):
    """expand_desc_kernel: bit k of the descriptor -> byte k, 1 -> +8, 0 -> -8."""
    bits = np.unpackbits(desc, axis=1, bitorder="little").astype(np.int32)
    return 16 * bits - 8


def model_knn2(q, t):
    Q, T = expand(q), expand(t)
    nt = t.shape[0]
    g = np.full((q.shape[0], 2), np.iinfo(np.int64).max, np.int64)          # running (best, second) of hamming << 22 | idx
    for t0 in range(0, nt, HALF):                                            # one thread's share of a tile: 64 columns
        tile = T[t0:t0 + HALF]
        acc = Q @ tile.T                                                     # 8 K steps: 64 (256 - 2 hamming)
        assert (acc % 128 == 0).all() and acc.min() >= -16384 and acc.max() <= 16384
        k16 = np.full((q.shape[0], HALF), -32768, np.int64)
        k16[:, :tile.shape[0]] = (acc + (63 - np.arange(tile.shape[0]))[None, :]).astype(np.int16)   # the IMAD of the epilogue
        streams = np.stack([k16[:, r::4].max(axis=1) for r in range(4)], 1)   # VIMNMX3 chains: columns r (mod 4)
        order = np.sort(streams, axis=1)[:, ::-1][:, :2]                      # the two best of the four stream maxima
        wide = ((128 - (order >> 7)) << 22) + t0 + (63 - (order & 127))      # widen_key()
        g = np.sort(np.concatenate([g, wide], 1), axis=1)[:, :2]
    # K2 refine_second_warp: the true second neighbour is g[:, 1] or one of the best's stream-mates
    dist, idx = g >> 22, g & 4194303
    for i in range(q.shape[0]):
        if dist[i, 0] > 256:
            continue
        t1 = int(idx[i, 0])
        mates = [t for t in range((t1 & ~63) + (t1 & 3), min((t1 & ~63) + 64, nt), 4) if t != t1]
        for tm in mates:
            d = int(np.unpackbits(q[i] ^ t[tm]).sum())
            key = (d << 22) + tm
            if key < g[i, 1]:
                g[i, 1] = key
    dist, idx = g >> 22, g & 4194303
    none = dist > 256                                                        # empty lane / nothing found
    return np.where(none, -1, idx), np.where(none, -1, dist)


@pytest.mark.parametrize("nq,nt,rand_bytes", [(5, 2, 32), (40, 127, 32), (33, 129, 1), (64, 300, 2), (17, 1000, 32)])
def test_model_equals_oracle(nq, nt, rand_bytes):
    rng = np.random.default_rng(nq * 131 + nt)
    q = np.zeros((nq, 32), np.uint8); t = np.zeros((nt, 32), np.uint8)
    q[:, :rand_bytes] = rng.integers(0, 256, (nq, rand_bytes)); t[:, :rand_bytes] = rng.integers(0, 256, (nt, rand_bytes))
    t[nt // 2] = t[0]; q[0] = 255 - t[1]; q[1] = t[1]                         # duplicates, distance 256 and distance 0
    im, dm = model_knn2(q, t)
    io, do = orc.knn2_hamming(q, t)
    assert np.array_equal(im, io) and np.array_equal(dm, do)


def test_key_is_monotone_in_distance_then_index():
    """Larger 16-bit key <=> (smaller distance, then smaller column); the widened key orders the same way under MIN."""
    d = np.arange(0, 257)[:, None]; c = np.arange(0, 64)[None, :]
    k16 = 128 * (128 - d) + (63 - c)
    flat = k16.ravel()
    assert len(np.unique(flat)) == flat.size and flat.min() == -16384 and flat.max() == 16447
    wide = ((128 - (k16 >> 7)) << 22) + (63 - (k16 & 127))
    assert np.array_equal(wide, (d << 22) + c)
    assert np.array_equal(np.argsort(-flat, kind="stable"), np.argsort(wide.ravel(), kind="stable"))
