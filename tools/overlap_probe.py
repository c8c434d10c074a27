#!/usr/bin/env python
"""Probe: the default bench batch (1024 Tsukuba pairs) split over C contexts / streams, S sub-batches each, so that the FP64
geometry kernels of one sub-batch can run under the tensor-core matcher of another.  Prints ms per 1024-pair step."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvslam_b200 as mvs  # noqa: E402

f = np.load(os.path.join(ROOT, "tests", "golden", "tsukuba_orb2000.npz"))
descs = [np.ascontiguousarray(f[f"desc{i}"]) for i in range(1, 6)]
kps = [np.ascontiguousarray(f[f"kp{i}"]) for i in range(1, 6)]
K = f["K"]
B = 1024
pairs = np.array([[(i % 4), (i % 4) + 1] for i in range(B)], np.int32)
item = mvs.RESULT_DTYPE.itemsize
res_t = torch.empty(B * item, dtype=torch.uint8).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for C, S in ((1, 1), (2, 1), (2, 2), (2, 4), (3, 2), (4, 2)):
    streams = [torch.cuda.Stream() for _ in range(C)]
    ctxs = [mvs.Context(0, stream=s.cuda_stream) for s in streams]
    for c in ctxs:
        c.frames_upload(descs, kps)
    n_sub = C * S
    bounds = [(B * k // n_sub, B * (k + 1) // n_sub) for k in range(n_sub)]
    main = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event(); ev.record(main)
        for s in streams:
            s.wait_event(ev)
        for k, (lo, hi) in enumerate(bounds):
            ctxs[k % C].pair_batch(pairs[lo:hi], K, max_dist=10.0, H=1024, seed=0, out=dict(results=res_t.data_ptr() + lo * item),
                                   enqueue_only=True, pair_id_base=lo)
        for s in streams:
            e = torch.cuda.Event(); e.record(s); main.wait_event(e)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ref = res_t.numpy().tobytes() if (C, S) == (1, 1) else ref
    same = res_t.numpy().tobytes() == ref
    tot = 0.0
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main); step(); b.record(main)
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    out[f"{C}x{S}"] = dict(ms_per_step=round(tot / 20, 4), pairs_per_s=round(B / (tot / 20 * 1e-3)), identical_records=bool(same))
    for c in ctxs:
        c.close()
print(json.dumps(out))
