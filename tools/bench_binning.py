"""Times the on-GPU binning of C3-sized float32 columns (10M rows) against the host rule.
Usage: python tools/bench_binning.py [rows] [columns]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ydf_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = torch.Generator(device="cuda").manual_seed(1234)
cols = []
for j in range(k):
    t = torch.randn(n, device="cuda", generator=g)
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h.copy_(t)
    cols.append(h.numpy())
torch.cuda.synchronize()
b = ydf_b200.DatasetBuilder(n, k)
b.add_numerical(0, cols[0], 256, 3)      # warm-up (allocates the scratch buffers)
times = []
for j in range(k):
    t0 = time.perf_counter()
    bounds, mean, na, miss = b.add_numerical(j, cols[j], 256, 3)
    times.append(time.perf_counter() - t0)
# pipelined: all columns enqueued (3 in flight), then collected
t0 = time.perf_counter()
for j in range(k):
    b.add_numerical_async(j, cols[j], 256, 3)
for j in range(k):
    bounds, mean, na, miss = b.get_numerical(j)
pipelined = (time.perf_counter() - t0) / k
ds = b.finish()
t0 = time.perf_counter()
wb, wmean = ydf_b200.discretize_boundaries(cols[k - 1], 256, 3)
t1 = time.perf_counter()
enc = ydf_b200.discretize_encode(cols[k - 1], wb, na)
t2 = time.perf_counter()
same = bool(np.array_equal(wb, bounds)) and bool(np.array_equal(ds.get_bins(k - 1), enc))
print(json.dumps({"rows": n, "columns": k, "gpu_s_per_column": float(np.median(times)), "gpu_s_all": times, "gpu_s_per_column_pipelined": pipelined,
                  "host_boundaries_s": t1 - t0, "host_encode_s": t2 - t1, "identical_to_host_rule": same,
                  "h2d_bytes_per_column": 4 * n, "speedup_vs_host_1_thread": (t2 - t0) / float(np.median(times))}))
