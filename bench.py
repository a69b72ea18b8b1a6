#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 batched-NTT library (contract: see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            our arm  (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): prime64::Plan, N = 2048,
p = Solinas 2^64 - 2^32 + 1, batch 65536 per GPU, one step = fwd pass + inv pass over the whole batch
(2 * 65536 NTTs per GPU).  Batches are independent, so N GPUs each transform their own 65536 polynomials
(weak scaling, no collective on the data path).  `value` = NTTs/s with inputs resident in HBM;
`e2e` = the same job through the host-slice C-ABI call (pinned host buffer -> H2D -> fwd -> inv -> D2H).
One JSON line on stdout (rank 0).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

N_POLY = 2048
BATCH = 65536
SOLINAS_P = 0xFFFFFFFF00000001
WORKLOAD = "prime64 Plan N=2048 p=2^64-2^32+1 (Solinas) batch 65536 per GPU, fwd+inv"
# algorithmic HBM bytes of one NTT launch over the batch: every word read once and written once
ALG_BYTES_PER_NTT = 2 * N_POLY * 8
# dram__bytes_read.sum + dram__bytes_write.sum of one k_ntt_cta<A64S,11> launch over the batch, from the
# ncu --set full capture committed under profiles/ (parsed at run time; None if the summary is absent)
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_kernels", "ncu_ntt64s_2048.txt")


def ncu_traffic(direction):
    """DRAM bytes per launch of k_ntt_cta<A64S,11,...,FWD> at grid 65536 from the committed ncu summary:
    template argument 5 is FWD (1 = forward, 0 = inverse)."""
    try:
        want = "1" if direction == "fwd" else "0"
        cur, tot = None, {}
        for line in open(NCU_SUMMARY):
            if line.startswith("=="):
                args = line.split("<", 1)[1].split(">", 1)[0].split(",")
                cur = args[4].strip()
                tot[cur] = 0.0
            elif "dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line:
                f = line.split()
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[-1]]
                tot[cur] += float(f[-2]) * scale
        return tot.get(want) or None
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled
    in-process every 2 ms (the timed region of the default run is ~50 ms, shorter than nvidia-smi's start-up);
    `nvidia-smi --query-gpu=... -lms` is the fallback when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.lines, self.proc = index, [], [], None
        self.nvml, self.handle, self.stop_flag, self.thread, self.smax = None, None, False, None, None

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._phys_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t_end = time.time() + 5.0
            while not self.lines and time.time() < t_end:   # wait for nvidia-smi's first sample
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((time.time(), sm, [k for k, bit in names if mask & bit], pw))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1.0)
            inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
            reasons = sorted({r for s in inside for r in s[2]})
            power = [s[3] for s in inside if s[3] is not None]
            return {"sm_mhz": float(np.median([s[1] for s in inside])) if inside else None, "sm_max_mhz": self.smax,
                    "reasons": reasons, "power_w_max": max(power) if power else None, "samples": len(inside), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                smax = float(f[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    power.append(float(f[3]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:  # region shorter than the sampling period: use every sample we have
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm), "source": "nvidia-smi"}


def host_threads():
    """every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which is not the host's core count)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(local):
    """Pin this process to the cores of the NUMA node its GPU hangs off (sysfs), BEFORE the pinned host buffers are allocated
    and first touched: eight ranks staging 4 GiB per step through one socket's memory controllers is what held the r01 e2e
    scaling at 1.6x on 8 GPUs.  Returns the node (None when the topology cannot be read; the process is then left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def make_inputs(rank, batch, n, p):
    g = np.random.Generator(np.random.PCG64(0xC0FFEE + 2 + 1000 * rank))
    a = g.integers(0, 2**64, size=(batch, n), dtype=np.uint64)
    a[a >= np.uint64(p)] -= np.uint64(p)       # uniform-enough in [0, p): values >= p (prob 2^-32) folded once
    return a


# ------------------------------------------------------------------------------------------------------------
CONFIG = {"workload": WORKLOAD, "n": N_POLY, "batch_per_gpu": BATCH, "prime": "0xFFFFFFFF00000001",
          "l2": "the 1 GiB batch per GPU is streamed every launch and exceeds the 126 MB L2 (no flush needed)"}


def cpu_plan():
    """The CPU arm's implementation of the path: the AVX-512 / AVX2 port of the reference's SIMD stage loops when the host has
    the ISA (oracle/cntt_simd.c, bit-checked against the scalar restatement in tests/test_oracle_simd.py), else the scalar
    restatement.  Returns (plan, description)."""
    from oracle import oracle as O
    try:
        O.build(native=True)
        native = True
    except Exception:
        native = False
    plan = O.Plan64.try_new(N_POLY, SOLINAS_P, native=native)
    isa = getattr(O, "simd_isa", lambda native=False: "scalar")(native)
    desc = ("C port of concrete-ntt's Solinas path, %s (oracle/%s, gcc -O3 -march=%s, OpenMP over polynomials); the crate itself "
            "cannot be built here (no cargo)" % ({"avx512": "AVX-512 stage loops as in src/prime64/generic_solinas.rs:132-446",
                                                 "avx2": "AVX2 stage loops as in src/prime64/generic_solinas.rs:132-446",
                                                 "scalar": "scalar restatement"}.get(isa, isa),
                                                "cntt_simd.c" if isa != "scalar" else "cntt_oracle.c", "native" if native else "x86-64-v3"))
    return plan, desc, isa


def run_reference(args):
    """The reference's own CPU algorithm for the path, all host threads, bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    plan, desc, isa = cpu_plan()
    sample = 16384      # 256 MiB per step: larger than the host caches, ~10-20 ms per step on 16 AVX-512 cores
    buf = make_inputs(0, sample, N_POLY, SOLINAS_P)
    for _ in range(max(1, args.warmup)):
        plan.fwd_batch(buf, threads)
        plan.inv_batch(buf, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.fwd_batch(buf, threads)
        plan.inv_batch(buf, threads)
    dt = time.perf_counter() - t0
    value = 2.0 * sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "NTTs/sec", "value": value, "unit": "NTT/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": dict(CONFIG), "sample_batch_per_step": sample,
        "cpu_baseline": {"value": value, "unit": "NTT/s", "cores": threads, "kind": "port", "isa": isa,
                         "sample": "%d polynomials of the workload per step (fwd+inv); %s" % (sample, desc)},
        "e2e": {"value": value, "unit": "NTT/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Bounded sample of the workload on the host cores (rank 0, N=1 only): ~10-20 core-seconds."""
    threads = host_threads()
    plan, desc, isa = cpu_plan()
    probe = make_inputs(7, 256, N_POLY, SOLINAS_P)
    plan.fwd_batch(probe, threads)
    t0 = time.perf_counter()
    plan.fwd_batch(probe, threads)
    plan.inv_batch(probe, threads)
    per_ntt_core_s = (time.perf_counter() - t0) * threads / 512.0
    sample = int(min(BATCH, max(8192, 4.0 / max(per_ntt_core_s, 1e-9) / 2)))   # >= 128 MiB: larger than the host caches
    buf = make_inputs(8, sample, N_POLY, SOLINAS_P)
    plan.fwd_batch(buf, threads)          # warm-up pass: page faults, thread pool
    plan.inv_batch(buf, threads)
    passes, t0 = 0, time.perf_counter()
    while passes < 3 or (time.perf_counter() - t0) * threads < 12.0:
        plan.fwd_batch(buf, threads)
        plan.inv_batch(buf, threads)
        passes += 1
    dt = time.perf_counter() - t0
    return {"value": 2.0 * sample * passes / dt, "unit": "NTT/s", "cores": threads, "kind": "port", "isa": isa,
            "sample": "%d of the %d polynomials, fwd+inv %d times after one warm-up pass (%.1f core-s); %s" % (sample, BATCH, passes, dt * threads, desc)}


def cpu_prime32_leg(n=1024, p=1062862849, sample=32768):
    """configs[0]'s transform on the host cores (rank 0, N=1 only): the 16- / 8-lane Shoup port of the reference's prime32 SIMD path
    (oracle/cntt_simd32.c, src/prime32/shoup.rs:305-451), OpenMP over polynomials, ~2 s."""
    from oracle import oracle as O
    try:
        O.build(native=True)
        native = True
    except Exception:
        native = False
    threads = host_threads()
    plan = O.Plan32.try_new(n, p, native=native)
    buf = np.random.default_rng(3).integers(0, p, size=(sample, n), dtype=np.uint32)   # 128 MiB: larger than the host caches
    plan.fwd_batch(buf, threads)
    plan.inv_batch(buf, threads)
    passes, t0 = 0, time.perf_counter()
    while passes < 3 or time.perf_counter() - t0 < 1.5:
        plan.fwd_batch(buf, threads)
        plan.inv_batch(buf, threads)
        passes += 1
    dt = time.perf_counter() - t0
    return {"ntts_per_s": 2.0 * sample * passes / dt, "cores": threads, "kind": "port", "isa": O.simd_isa(native),
            "sample": "%d polynomials, fwd+inv %d times after one warm-up pass" % (sample, passes)}


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None   # host buffers of the e2e leg next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cntt = importlib.import_module("concrete-ntt_b200")
    plan = cntt.prime64.Plan.try_new(N_POLY, SOLINAS_P, device=local)
    assert plan is not None
    peak, peak_src = peaks()

    host = make_inputs(rank, BATCH, N_POLY, SOLINAS_P)
    pinned = torch.empty((BATCH, N_POLY), dtype=torch.int64).pin_memory()
    pinned.numpy().view(np.uint64)[...] = host
    d = pinned.cuda(non_blocking=False)
    d_orig = d.clone()
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    lib = cntt._lib.lib()
    h = plan._h
    dptr = d.data_ptr()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def verify(buf, rounds, what):
        """inv(fwd(x)) = N x, so after `rounds` round trips `rounds` normalisations must give the input back, bit for bit:
        a launch that wrote garbage cannot produce a bench line."""
        for _ in range(rounds):
            plan.normalize(buf)
        torch.cuda.synchronize()
        if not torch.equal(buf, d_orig):
            bad = int((buf != d_orig).sum().item())
            raise RuntimeError("bench.py: %s does not round-trip (%d of %d words differ after %d fwd+inv steps)" % (what, bad, buf.numel(), rounds))
        return "%d x (fwd, inv) then %d x normalize == input, all %d words" % (rounds, rounds, buf.numel())

    W = max(3, args.warmup)
    for _ in range(W):
        st = lib.cntt_prime64_fwd(h, dptr, BATCH, sptr)
        st |= lib.cntt_prime64_inv(h, dptr, BATCH, sptr)
        if st:
            raise RuntimeError("kernel launch failed: " + lib.cntt_last_cuda_error().decode())
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    barrier()
    t_wall0 = time.time()
    for i in range(K):
        ev[i][0].record(stream)
        st = lib.cntt_prime64_fwd(h, dptr, BATCH, sptr)
        ev[i][1].record(stream)
        st |= lib.cntt_prime64_inv(h, dptr, BATCH, sptr)
        ev[i][2].record(stream)
        if st:
            raise RuntimeError("kernel launch failed: " + lib.cntt_last_cuda_error().decode())
    barrier()
    t_wall1 = time.time()
    total_ms = ev[0][0].elapsed_time(ev[K - 1][2])
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    inv_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_ms = max_over_ranks(total_ms)
    value = world * 2.0 * BATCH * K / (total_ms * 1e-3)
    checked = {"device_resident": verify(d, W + K, "the timed device buffer")}

    # ---- end to end through the reference's call shape: plan.fwd(buf); plan.inv(buf) on a HOST slice -- two C-ABI calls per
    #      step, each staging its own H2D and D2H (cntt_prime64_fwd_host, cntt_prime64_inv_host).  The fused
    #      cntt_prime64_fwd_inv_host (one upload + one download for both transforms, an extension) is reported beside it.
    hview = pinned.numpy().view(np.uint64)
    nbytes = BATCH * N_POLY * 8
    ke = max(2, min(K, 4))

    def e2e(fn, rounds_per_call):
        hview[...] = host
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            fn()                   # synchronous: returns after the D2H of the last chunk
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        d.copy_(pinned)
        return world * 2.0 * BATCH * ke / dt, verify(d, (ke + 1) * rounds_per_call, "the end-to-end host buffer")

    def two_calls():
        plan.fwd(hview)
        plan.inv(hview)

    e2e_value, checked["e2e"] = e2e(two_calls, 1)
    e2e_fused, checked["e2e_fused"] = e2e(lambda: plan.fwd_inv(hview), 1)
    del d_orig

    def timeit(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / reps)

    # ---- secondary figure of the same metric line: native64 polymul (BASELINE configs[2]), device-resident
    extra = {}
    try:
        npl = cntt.native64.Plan32.try_new(N_POLY, device=local)
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        lhs = torch.randint(-2**63, 2**63 - 1, (BATCH, N_POLY), dtype=torch.int64, device="cuda", generator=g)
        rhs = torch.randint(-2**63, 2**63 - 1, (BATCH, N_POLY), dtype=torch.int64, device="cuda", generator=g)
        prod = torch.empty_like(lhs)
        pm_ms = timeit(lambda: npl.negacyclic_polymul(prod, lhs, rhs), 5, 2)
        extra["native64_polymul_n2048_b65536"] = {
            "polymuls_per_s": world * BATCH / (pm_ms * 1e-3), "ms_per_batch": pm_ms,
            "hbm_frac": (3 * N_POLY * 8 * BATCH / (pm_ms * 1e-3) / 1e9) / peak,
            "butterflies_per_clk_per_sm": 168960.0 * BATCH / (pm_ms * 1e-3) / 148 / 1.965e9,
        }
        del lhs, rhs, prod
    except Exception as e:  # never let the secondary figure kill the headline
        extra["native64_polymul_error"] = repr(e)
    # ---- BASELINE configs[0] shape at batch 65536 ("the same kernel at batch 2^16"), device-resident
    try:
        p0 = 1062862849
        pl32 = cntt.prime32.Plan.try_new(1024, p0, device=local)
        g = torch.Generator(device="cuda").manual_seed(99 + rank)
        d32 = torch.randint(0, p0, (BATCH, 1024), dtype=torch.int32, device="cuda", generator=g)
        res = {"fwd": timeit(lambda: pl32.fwd(d32), 10), "inv": timeit(lambda: pl32.inv(d32), 10)}
        extra["prime32_n1024_p0_b65536"] = {
            "fwd_ntts_per_s": BATCH / (res["fwd"] * 1e-3), "inv_ntts_per_s": BATCH / (res["inv"] * 1e-3),
            "hbm_frac_fwd": 2 * 1024 * 4 * BATCH / (res["fwd"] * 1e-3) / 1e9 / peak,
            "hbm_frac_inv": 2 * 1024 * 4 * BATCH / (res["inv"] * 1e-3) / 1e9 / peak, "per": "GPU"}
        del d32
        if rank == 0 and world == 1:
            extra["prime32_n1024_p0_b65536"]["cpu"] = cpu_prime32_leg()
    except Exception as e:
        extra["prime32_error"] = repr(e)

    # ---- BASELINE configs[3] and configs[4], device-resident: native128 N=4096 batch 8192 per GPU, and native_binary64
    #      N=65536 (extended plan) with the batch of 1024 split over the ranks
    def time_polymul(pplan, shape, binary, reps=5):
        g = torch.Generator(device="cuda").manual_seed(4321 + rank)
        a = torch.randint(-2**63, 2**63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
        b = torch.randint(-2**63, 2**63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
        if binary:
            b &= 1
        out = torch.empty_like(a)
        return timeit(lambda: pplan.negacyclic_polymul(out, a, b), reps, 2)
    try:
        ms = time_polymul(cntt.native128.Plan32.try_new(4096, device=local), (8192, 4096, 2), False)
        extra["native128_polymul_n4096_b8192"] = {"polymuls_per_s": world * 8192 / (ms * 1e-3), "ms_per_batch": ms,
                                                  "hbm_frac": (3 * 4096 * 16 * 8192 / (ms * 1e-3) / 1e9) / peak}
    except Exception as e:
        extra["native128_polymul_error"] = repr(e)
    try:
        per = max(1, 1024 // world)
        ms = time_polymul(cntt.native_binary64.Plan32.try_new_extended(65536, device=local), (per, 65536), True)
        extra["native_binary64_polymul_n65536_b1024_total"] = {"polymuls_per_s": world * per / (ms * 1e-3), "ms_per_batch": ms,
                                                               "batch_per_gpu": per, "scaling": "strong",
                                                               "hbm_frac_per_gpu": (3 * 65536 * 8 * per / (ms * 1e-3) / 1e9) / peak}
    except Exception as e:
        extra["native_binary64_polymul_error"] = repr(e)

    # ---- the metric over N = 2^10 .. 2^16 (BASELINE.json "metric"): prime32 (P0), prime64 Solinas and native64 polymul, device-
    #      resident, 2^27 words per batch (2^26 for the polymul operands), per GPU, with the HBM fraction of each.  Every rank
    #      runs it (weak scaling: the figures are per GPU, max time over ranks).
    sweep = {}
    for logn in range(10, 17):
        n = 1 << logn
        row = {}
        try:
            b32 = (1 << 27) >> logn
            pl = cntt.prime32.Plan.try_new(n, 1062862849, device=local)
            g = torch.Generator(device="cuda").manual_seed(7 + rank)
            x = torch.randint(0, 1062862849, (b32, n), dtype=torch.int32, device="cuda", generator=g)
            f, i = timeit(lambda: pl.fwd(x), 5), timeit(lambda: pl.inv(x), 5)
            row["prime32"] = {"batch": b32, "fwd_ntts_per_s": b32 / (f * 1e-3), "inv_ntts_per_s": b32 / (i * 1e-3),
                              "hbm_frac_fwd": 2 * n * 4 * b32 / (f * 1e-3) / 1e9 / peak, "hbm_frac_inv": 2 * n * 4 * b32 / (i * 1e-3) / 1e9 / peak}
            del x, pl
            pl = cntt.prime64.Plan.try_new(n, SOLINAS_P, device=local)
            x = torch.randint(0, 2**62, (b32, n), dtype=torch.int64, device="cuda", generator=g)
            f, i = timeit(lambda: pl.fwd(x), 5), timeit(lambda: pl.inv(x), 5)
            row["prime64_solinas"] = {"batch": b32, "fwd_ntts_per_s": b32 / (f * 1e-3), "inv_ntts_per_s": b32 / (i * 1e-3),
                                      "hbm_frac_fwd": 2 * n * 8 * b32 / (f * 1e-3) / 1e9 / peak, "hbm_frac_inv": 2 * n * 8 * b32 / (i * 1e-3) / 1e9 / peak}
            del x, pl
            bp = (1 << 26) >> logn
            pp = cntt.native64.Plan32.try_new(n, device=local)
            ext = pp is None
            if ext:   # the reference has no native64 plan at N = 65536 (P1 - 1 = 2^16 odd): extended prime set, DESIGN.md section 8
                pp = cntt.native64.Plan32.try_new_extended(n, device=local)
            a = torch.randint(-2**63, 2**63 - 1, (bp, n), dtype=torch.int64, device="cuda", generator=g)
            b = torch.randint(-2**63, 2**63 - 1, (bp, n), dtype=torch.int64, device="cuda", generator=g)
            o = torch.empty_like(a)
            ms = timeit(lambda: pp.negacyclic_polymul(o, a, b), 3, 2)
            row["native64_polymul"] = {"batch": bp, "polymuls_per_s": bp / (ms * 1e-3), "hbm_frac": 3 * n * 8 * bp / (ms * 1e-3) / 1e9 / peak,
                                       "butterflies_per_clk_per_sm": 15 * (n // 2) * logn * bp / (ms * 1e-3) / 148 / 1.965e9, "extended_primes": ext}
            del a, b, o, pp
        except Exception as e:
            row["error"] = repr(e)
        sweep["n%d" % n] = row
    extra["sweep_per_gpu"] = sweep

    if rank == 0:
        slow_ms, which = (fwd_ms, "fwd") if fwd_ms >= inv_ms else (inv_ms, "inv")
        achieved = ALG_BYTES_PER_NTT * BATCH / (slow_ms * 1e-3) / 1e9
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        bf = 11264.0 * BATCH / (slow_ms * 1e-3) / 148 / (mhz * 1e6)
        line = {
            "metric": "NTTs/sec", "value": value, "unit": "NTT/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": dict(CONFIG),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "NTT/s", "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": 2 * nbytes, "steps": ke,
                    "api": "the reference's call shape, plan.fwd(buf); plan.inv(buf) on a pinned host slice = cntt_prime64_fwd_host + "
                           "cntt_prime64_inv_host (each stages its own chunked, double-buffered H2D and D2H)"},
            "e2e_fused": {"value": e2e_fused, "unit": "NTT/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": ke,
                          "api": "cntt_prime64_fwd_inv_host (extension: both transforms between one upload and one download)"},
            "checked": checked, "numa_node_rank0": numa_node,
            "gpu_launches": 2 * K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(which), "traffic_source": os.path.relpath(NCU_SUMMARY, ROOT) + " (ncu --set full, dram read+write bytes per launch)",
                         "kernel": "k_ntt_cta<A64S,11,4> (%s)" % which,
                         "peak_source": peak_src, "ms_per_launch": {"fwd": fwd_ms, "inv": inv_ms},
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_NTT * BATCH,
                         # integer roofline of the same launch (11 levels x 1024 butterflies per NTT).  Peak = the multiply floor: a
                         # 64 x 64 -> 128-bit product is four 32 x 32 -> 64 products and IMAD.WIDE.U32 issues at 24.4 per clock and SM
                         # (profiles/r01_ubench_int_pipes.txt), i.e. 6.1 butterflies per clock and SM if nothing else cost anything.  The
                         # butterflies as shipped, alone in registers, reach 2.40 (multiply) and 3.25 (shift, four of the eleven levels):
                         # 11 / (7 / 2.40 + 4 / 3.25) = 2.65 for the mix of this transform (profiles/r02_ubench_gold_bf.txt).
                         "int_roofline": {"achieved": bf, "peak": 24.4 / 4, "unit": "butterflies/clk/SM", "frac": bf / (24.4 / 4),
                                          "peak_source": "4-product multiply floor, measured IMAD.WIDE.U32 rate / 4",
                                          "butterflies_alone": {"multiply": 2.40, "shift": 3.25, "mix_of_this_transform": 2.65},
                                          "frac_of_butterflies_alone": bf / 2.65},
                         "note": "integer-bound kernel: ALU and multiply pipe ~70-75 % busy at once (ncu summary under profiles/); "
                                 "the HBM fraction cannot exceed ~0.4 for this prime on CUDA cores, see DESIGN.md section 4"},
            "extra": extra,
        }
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_leg()
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "NTT/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
