import os, sys, time, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench, mvslam_b200 as mvs
descs, kps, K, pairs, params, cfg = bench.load_workload("seq", 1024, 256)
nk = descs[0].shape[0]
D = torch.from_numpy(np.concatenate(descs)).pin_memory(); P = torch.from_numpy(np.concatenate(kps)).pin_memory()
kw = dict(max_dist=params["max_dist"], H=params["H"], seed=0, mode=params["mode"], solver="fast")
item = mvs.RESULT_DTYPE.itemsize
res_t = torch.empty(1024 * item, dtype=torch.uint8).pin_memory()
for chunk in (1024, 512, 256, 128):
    ctx = mvs.Context(0)
    nf = chunk + 1
    ctx.frames_upload_packed(D.data_ptr(), P.data_ptr(), np.full(nf, nk, np.int32)); ctx.synchronize()
    loc = np.stack([np.arange(nf - 1), np.arange(1, nf)], 1).astype(np.int32)
    for _ in range(3): ctx.pair_batch(loc, K, out=dict(results=res_t.data_ptr()), **kw)
    ctx.profile_enable(True); ctx.profile_read()
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.pair_batch(loc, K, out=dict(results=res_t.data_ptr()), **kw); ts.append((time.perf_counter() - t0) * 1e3)
    pr = ctx.profile_read()
    st = {k: round(v[0] / 10, 4) for k, v in pr.items() if v[0] > 0}
    # upload alone
    us = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.frames_upload_packed(D.data_ptr(), P.data_ptr(), np.full(nf, nk, np.int32)); ctx.synchronize(); us.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps(dict(chunk=chunk, call_ms=round(float(np.median(ts)), 4), stages_sum=round(sum(st.values()), 4), stages=st, upload_ms=round(float(np.median(us)), 4))))
    ctx.close()
