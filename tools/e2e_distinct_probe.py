#!/usr/bin/env python
"""A/B of the chunking of bench.py's e2e_distinct job (1025 distinct frames -> 1024 consecutive pairs): prints one line per
(chunk size, context count)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import mvslam_b200 as mvs  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for chunk, n_ctx in [(256, 4), ([384, 320, 192, 128], 4), ([320, 320, 256, 128], 4), ([448, 320, 192, 64], 4), ([512, 256, 128, 128], 4),
                     ([192, 256, 256, 192, 128], 5), ([128, 384, 384, 128], 4)]:
    r = bench.e2e_distinct_block(mvs, torch, 0, flush, 1024, chunk, n_ctx)
    print(json.dumps(dict(chunk=chunk, contexts=n_ctx, ms=r["ms_per_step"]["median"], value=r["value"], frac=r["frac_of_device_resident"])))
