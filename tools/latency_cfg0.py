"""BASELINE configs[0]: prime32 Plan N=1024 p=P0, fwd+inv round trip at batch 1 -- latency, not throughput.
Device-resident (two launches on one stream, CUDA events over 1000 round trips) and through the host-slice call."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cntt = importlib.import_module("concrete-ntt_b200")
n, p = 1024, 1062862849
plan = cntt.prime32.Plan.try_new(n, p)
for batch in (1, 16, 256):
    d = torch.randint(0, p, (batch, n), dtype=torch.int32, device="cuda")
    for _ in range(50): plan.fwd(d); plan.inv(d); plan.normalize(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 1000
    e0.record()
    for _ in range(reps): plan.fwd(d); plan.inv(d)
    e1.record(); torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) * 1e3 / reps
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        plan.fwd(d); plan.inv(d)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): plan.fwd(d); plan.inv(d)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100): g.replay()
    e1.record(); torch.cuda.synchronize()
    graph_us = e0.elapsed_time(e1) * 1e3 / 1000
    h = np.random.default_rng(1).integers(0, p, size=(batch, n), dtype=np.uint64).astype(np.uint32)
    for _ in range(20): plan.fwd_inv(h)
    t0 = time.perf_counter()
    for _ in range(200): plan.fwd_inv(h)
    host_us = (time.perf_counter() - t0) * 1e6 / 200
    print("prime32 N=1024 batch=%d fwd+inv: device-resident %.2f us per round trip (launch-bound), CUDA graph %.2f us, host slice (H2D + 2 kernels + D2H) %.1f us" % (batch, dev_us, graph_us, host_us))
