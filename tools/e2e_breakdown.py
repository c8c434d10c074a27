import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import mvslam_b200 as mvs
f = np.load("tests/golden/tsukuba_orb2000.npz")
descs = [np.ascontiguousarray(f[f"desc{i}"]) for i in range(1, 6)]; kps = [np.ascontiguousarray(f[f"kp{i}"]) for i in range(1, 6)]
K = f["K"]; n = 1024; cap = 256
pairs = np.array([[(i % 4), (i % 4) + 1] for i in range(n)], np.int32)
pin = lambda a: torch.from_numpy(a).pin_memory()
pd = [pin(d) for d in descs]; pk = [pin(k) for k in kps]
item = mvs.RESULT_DTYPE.itemsize
res_t = torch.empty(n * item, dtype=torch.uint8).pin_memory()
mat_t = torch.empty(n * cap * 12, dtype=torch.uint8).pin_memory(); msk_t = torch.empty(n * cap, dtype=torch.uint8).pin_memory()
pts_t = torch.empty(n * cap * 3, dtype=torch.float64).pin_memory(); idx_t = torch.empty(n * cap, dtype=torch.int64).pin_memory()
stream = torch.cuda.current_stream()
ctx = mvs.Context(0, stream=stream.cuda_stream)
out_rec = dict(results=res_t.data_ptr())
out_all = dict(results=res_t.data_ptr(), matches=mat_t.data_ptr(), mask=msk_t.data_ptr(), points=pts_t.data_ptr(), indexes=idx_t.data_ptr(), capacity=cap)
kw = dict(max_dist=10.0, H=1, solver="reference")
def upload(): ctx.frames_upload([t.numpy() for t in pd], [t.numpy() for t in pk])
def T(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(float(np.median(ts)), 4)
upload()
print("upload only", T(upload))
print("batch records only (sync call)", T(lambda: ctx.pair_batch(pairs, K, out=out_rec, **kw)))
print("batch all details (sync call)", T(lambda: ctx.pair_batch(pairs, K, out=out_all, **kw)))
print("batch records, enqueue+sync", T(lambda: ctx.pair_batch(pairs, K, out=out_rec, enqueue_only=True, **kw)))
print("upload + batch all", T(lambda: (upload(), ctx.pair_batch(pairs, K, out=out_all, **kw))))
ctx.profile_enable(True); ctx.profile_read()
for _ in range(10): ctx.pair_batch(pairs, K, out=out_rec, **kw)
pr = ctx.profile_read(); print("device stages sum", round(sum(v[0] for v in pr.values()) / 10, 4))
