#!/usr/bin/env python
"""Turn the scratch captures in gpurun_out/ into the tracked summaries under profiles/ (run in the build container)."""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(P, exist_ok=True)

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def stall_summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    res, cur, H = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], stalls=collections.Counter(), ops=collections.Counter()); res.append(cur)
        elif r and r[0] == "Address":
            H = {h: i for i, h in enumerate(r)}
        elif cur is not None and H and len(r) > 10:
            for h, i in H.items():
                if h.startswith("stall_") and "Not Issued" not in h:
                    cur["stalls"][h[6:]] += int(r[i] or 0)
            toks = [t for t in r[H["Source"]].split() if not t.startswith("@")]
            if toks:
                cur["ops"][toks[0].split(".")[0]] += int(r[H["Instructions Executed"]] or 0)
    return res


lines = [f"# ncu summaries ({R})", "",
         "Captured on one B200 with `tools/collect_profiles.sh` (`ncu --set full --clock-control none --import-source on`),",
         "after the same command had exited 0 without ncu.  Values are per launch; ncu times are cold-cache and serialised.", ""]
for rep, title in ((f"prof_path_{R}.ncu-rep", "two-view path (bench.py default workload, 1024 Tsukuba pairs, " + ("H=1024)" if R == "r1" else "H=1, REFERENCE solver)")),
                   (f"prof_fast_{R}.ncu-rep", "FAST solver, H=1024 on the same pairs (bench.py --solver fast --hypotheses 1024): hypotheses / score / triangulate"),
                   (f"prof_score_s8k_{R}.ncu-rep", "score_kernel on the S8k workload (64 pairs x 8192 kpts, H=4096, Sampson)"),
                   (f"prof_l2_{R}.ncu-rep", "l2_gemm_topk_kernel, 32768 x 32768 x 64 float descriptors"),
                   (f"prof_orb_{R}.ncu-rep", "feature extraction (tools/orb_bench.py: 256 Tsukuba frames per call, nfeatures 2000)"),
                   (f"prof_pnp_{R}.ncu-rep", "pnp_solve (tools/pnp_bench.py: 1024 problems x 500 points x 100 hypotheses)"),
                   (f"prof_ba_{R}.ncu-rep", "bundle adjustment (tools/ba_bench.py: 1024 two-view problems x 200 points)")):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    H, U, data = raw(path)
    idx = {h: i for i, h in enumerate(H)}
    lines += [f"## {title}", "", "| kernel | " + " | ".join(k.split(".")[0].replace("sm__inst_executed_pipe_", "pipe_") for k in KEEP if k in idx) + " |",
              "|---|" + "---|" * sum(k in idx for k in KEEP)]
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("mvs::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        vals = []
        for k in KEEP:
            if k in idx:
                v = r[idx[k]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(v + (" " + U[idx[k]] if U[idx[k]] not in ("", "%") and k.endswith(".sum") and "bytes" in k or k.startswith("gpu__time") else ""))
        lines.append(f"| {name} | " + " | ".join(vals) + " |")
    lines.append("")
    for k in stall_summary(path):
        tot = sum(k["stalls"].values()) or 1
        top = ", ".join(f"{n} {v / tot:.0%}" for n, v in k["stalls"].most_common(5))
        ops = ", ".join(f"{n} {v:,}" for n, v in k["ops"].most_common(8))
        nm = k["name"].split("(")[0].replace("void ", "").replace("mvs::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        lines += [f"* `{nm}` — warp-stall samples: {top}; executed warp instructions: {ops}"]
    lines.append("")
open(os.path.join(P, f"ncu_summary_{R}.md"), "w").write("\n".join(lines))

# launch list -> share per kernel
lp = os.path.join(G, f"launches_{R}.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, f"ncu_launches_{R}.csv"))
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hi]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) < len(H):
            continue
        d = dict(zip(H, r))
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else v * 1e3 if d["Metric Unit"] == "ms" else v
        k = d["Kernel Name"].split("(")[0].replace("void ", "")
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"ncu_launch_shares_{R}.md"), "w") as f:
        f.write(f"# kernel shares of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras` ({R})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n\n"
                "| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |\n")

# DRAM traffic per launch of every kernel of the default workload (bench.py fills roofline.traffic from it)
pp = os.path.join(G, f"prof_path_{R}.ncu-rep")
if os.path.exists(pp):
    H, U, data = raw(pp)
    idx = {h: i for i, h in enumerate(H)}
    traffic = {}
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("mvs::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("unnamed>::", "").split("<")[0]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        b = sum(float(r[idx[k]].replace(",", "")) * scale.get(U[idx[k]], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[name] = int(b)
    json.dump(dict(workload="tsukuba_vo_2k", source=f"profiles/ncu_summary_{R}.md (ncu --set full, per launch)", dram_bytes=traffic),
              open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)

for f in os.listdir(G):
    if f.endswith(f"_{R}.json"):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
print(open(os.path.join(P, f"ncu_summary_{R}.md")).read()[:6000])
