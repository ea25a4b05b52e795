#!/usr/bin/env python
"""bench.py -- headline benchmark of the packed-genotype hot path (BASELINE.json metric:
"genotypes/sec in bed_prodVec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one bed_prodVec (X~ . y, binomial center/scale, all rows, all columns) over the resident synthetic
.bed.  Workload (default, `cfg2`) = BASELINE.json configs[1]: 50,000 samples x 500,000 SNPs per GPU; with N
GPUs every rank holds its own 500,000-column shard (weak scaling) and the step ends in one NCCL all-reduce of
the n-vector of partial products.  `--workload cfg5` is the UKBB-shaped 487,000 x 1,100,000 matrix split by
columns over the N ranks (strong scaling; needs N >= 1 with 134 GB, sample-major copy only from N >= 2).

Printed JSON (one line, rank 0): the base contract + `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches`.  `value` has inputs resident in HBM; `e2e` goes through the 9-argument C-ABI call
(bsg_prodvec) with pinned HOST buffers, H2D/D2H inside the timed region.

`--impl reference` times the reference's CPU implementation of the same call on the host cores: the literal
C/OpenMP port in oracle/ (the reference needs R + Rcpp + bigstatsr and cannot be built here), all host
threads, on a bounded column sample of the same synthetic matrix.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 20250924 + 1  # 20250924 + config index
WORKLOADS = {
    "cfg2": dict(n=50_000, m=500_000, scaling="weak",
                 name="configs[1]: bed_prodVec on synthetic 50,000 x 500,000 2-bit .bed per GPU (binomial center/scale)"),
    "cfg5": dict(n=487_000, m=1_100_000, scaling="strong",
                 name="configs[4]: bed_prodVec on UKBB-shaped synthetic 487,000 x 1,100,000 .bed, SNP columns sharded over the GPUs"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.err = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
                nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get_reasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), **({"error": self.err} if self.err else {})}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def host_thread_candidates():
    """Thread counts worth trying for the CPU port: every CPU this process may run on and one per physical core.
    Taken from the OS, not from OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its workers); the oracle's
    loops take the count as their `ncores` argument, like the reference's."""
    try:
        logical = len(os.sched_getaffinity(0))
    except Exception:
        logical = os.cpu_count() or 1
    cands = {max(1, logical)}
    try:
        import psutil

        phys = psutil.cpu_count(logical=False)
        if phys:
            cands.add(max(1, min(logical, phys)))
    except Exception:
        pass
    return sorted(cands)


def cpu_port_rate(n, m_cols, seed, steps=1, warmup=0, threads=None):
    """genotypes/s of the oracle's bed_pMatVec4 port on a column sample of the synthetic matrix."""
    from oracle import ref

    threads = threads or ref.max_threads()
    o = ref.synth_bed(n, m_cols, seed=seed)
    sc = ref.bed_scaleBinom(o, ncores=threads)
    y = np.random.default_rng(seed + 1).normal(size=m_cols)
    ir, ic = o.rows_along(), o.cols_along()
    for _ in range(warmup):
        ref.bed_pMatVec4(o, ir, ic, sc["center"], sc["scale"], y, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.bed_pMatVec4(o, ir, ic, sc["center"], sc["scale"], y, threads)
    dt = time.perf_counter() - t0
    return n * m_cols * steps / dt, dt / steps, threads


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use the host's cores, and libgomp
    # reads the variable once, when the oracle library is loaded below
    os.environ["OMP_NUM_THREADS"] = str(max(host_thread_candidates()))
    from oracle import ref

    n = wl["n"]
    # calibrate warm on 20,000 columns (the first OpenMP regions of a process are slow) with every logical CPU and
    # with one thread per physical core, keep the faster (SMT oversubscription halves this loop's rate on some
    # hosts); then size a step to ~min(4 s, 150 s / (K + W)), at most the whole per-GPU matrix
    rate0, threads = 0.0, 1
    for t in host_thread_candidates():
        r, _, _ = cpu_port_rate(n, 20000, SEED, steps=2, warmup=2, threads=t)
        if r > rate0:
            rate0, threads = r, t
    per_step = max(0.5, min(4.0, 150.0 / max(1, args.steps + args.warmup)))
    m_s = int(max(20000, min(wl["m"], rate0 * per_step / n)))
    rate, sec_step, _ = cpu_port_rate(n, m_s, SEED, steps=args.steps, warmup=args.warmup, threads=threads)
    sample = "first %d of %d SNP columns of the synthetic %d-sample matrix (%.2e genotypes per step)" % (
        m_s, wl["m"], n, float(n) * m_s)
    line = {
        "impl": "reference", "metric": "genotypes/sec in bed_prodVec", "value": rate, "unit": "genotypes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_step * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": wl["name"], "n": n, "m_sample": m_s},
        "cpu_baseline": {"value": rate, "unit": "genotypes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "genotypes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--na-rate", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-svd", action="store_true", help="skip the bed_randomSVD wall-time leg")
    ap.add_argument("--layout", choices=("both", "snp"), default="both",
                    help="both (default): SNP-major + sample-major copies, X.y on k_pmv; snp: the SNP-major copy only "
                         "(the library's own default), X.y on the transposing kernel k_pmvT")
    ap.add_argument("--no-single-copy", action="store_true", help="skip the extra leg timing X.y on the SNP-major copy alone")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        # rank 0 prints exactly one line on stdout: keep NCCL's own "NCCL version ..." banner (NCCL_DEBUG=VERSION,
        # which some launch environments export) out of it; INFO / TRACE requests are respected
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"  # WARN still prints the banner (VERSION < WARN in NCCL's levels)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # whatever NCCL does print stays off stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import bigsnpr_b200 as B
    from bigsnpr_b200 import _lib, build

    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    L = _lib.lib()

    n = wl["n"]
    if wl["scaling"] == "weak":
        m_loc, col0 = wl["m"], rank * wl["m"]
    else:
        per = (wl["m"] + world - 1) // world
        col0 = rank * per
        m_loc = max(0, min(wl["m"], col0 + per) - col0)
    g = B.Bed.synthetic(n, m_loc, seed=SEED, na_rate=args.na_rate, col_offset=col0, device=local,
                        layouts=B.LAYOUT_SNP_MAJOR if args.layout == "snp" else (B.LAYOUT_SNP_MAJOR | B.LAYOUT_SAMPLE_MAJOR))
    layouts = g.layouts
    sc = B.bed_scaleBinom(g)
    view = B.View(g, center=sc["center"], scale=sc["scale"])
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + 17 + rank)
    x = torch.randn(m_loc, dtype=torch.float64, device=dev, generator=gen)
    out = torch.zeros(n, dtype=torch.float64, device=dev)
    # a dedicated (non-default) stream: the library enqueues on the stream it is handed, and the CUDA events
    # that time the region are recorded on the same stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        view.prodvec_dev(x.data_ptr(), out.data_ptr(), stream)
        if world > 1:
            dist.all_reduce(out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---------------- timed region: K steps, inputs resident in HBM ----------------
    sampler = ClockSampler(physical_gpu_index(local))
    L.bsg_set_kernel_timing(1)
    launches0 = L.bsg_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = int(L.bsg_launch_count() - launches0)
    import ctypes as C

    cnt, tot = C.c_int(0), C.c_double(0)
    _lib.check(L.bsg_kernel_time_stats(C.byref(cnt), C.byref(tot)))
    L.bsg_set_kernel_timing(0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    geno_step = float(n) * float(m_loc)
    tg = torch.tensor([geno_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tg)
    value = float(tg.item()) * args.steps / (ms / 1e3)

    # roofline of the dominant kernel (k_pmv): algorithmic bytes = ceil(n/4) * m per launch
    alg_bytes = float((n + 3) // 4) * m_loc
    kern_ms = tot.value / max(cnt.value, 1)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_pmv_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": ("bsg::pmvt::k_pmvT2" if g.has_na else
                           ("bsg::pmv::k_pmv" if (layouts & 2) else "bsg::pmvt::k_pmvT")),
                "kernel_ms": kern_ms, "launches_timed": cnt.value, "algorithmic_bytes_per_launch": alg_bytes,
                "peak_source": peak_src, "kernel_share_of_step": (kern_ms * args.steps / ms) if ms > 0 else None}

    # ---------------- e2e: the 9-argument C-ABI call with pinned host buffers ----------------
    e2e_steps = max(3, min(args.steps, 50))
    xc = torch.empty(m_loc, dtype=torch.float64).pin_memory()
    xc.copy_(x.cpu())
    cen = torch.from_numpy(np.ascontiguousarray(sc["center"])).pin_memory()
    sca = torch.from_numpy(np.ascontiguousarray(sc["scale"])).pin_memory()
    outh = torch.empty(n, dtype=torch.float64).pin_memory()
    pd = lambda tns: C.cast(tns.data_ptr(), _lib.c_dbl_p)  # noqa: E731

    def e2e_step():
        _lib.check(L.bsg_prodvec(g._h, None, n, None, m_loc, pd(cen), pd(sca), pd(xc), pd(outh)))
        if world > 1:
            od = outh.to(dev, non_blocking=True)
            dist.all_reduce(od)
            outh.copy_(od)

    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = float(tg.item()) * e2e_steps / float(te.item())
    # agreement of the two paths (device-resident vs host call) on this rank's data
    chk = float((out.cpu() - outh).abs().max().item()) if world == 1 else None

    # ---------------- extra leg: the same product from the SNP-major copy alone (k_pmvT) ----------------
    single_copy = None
    if (layouts & 2) and not args.no_single_copy:
        _lib.check(L.bsg_set_prodvec_path(1))
        try:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            L.bsg_set_kernel_timing(1)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ksc = max(3, min(args.steps, 20))
            s0.record()
            for _ in range(ksc):
                step()
            s1.record()
            torch.cuda.synchronize()
            c2, t2 = C.c_int(0), C.c_double(0)
            _lib.check(L.bsg_kernel_time_stats(C.byref(c2), C.byref(t2)))
            L.bsg_set_kernel_timing(0)
            kms = t2.value / max(c2.value, 1)
            single_copy = {"kernel": "bsg::pmvt::k_pmvT", "steps": ksc, "ms_per_step": s0.elapsed_time(s1) / ksc,
                           "kernel_ms": kms, "achieved": alg_bytes / (kms / 1e3) / 1e9 if kms > 0 else None,
                           "frac": (alg_bytes / (kms / 1e3) / 1e9 / peak) if kms > 0 else None,
                           "note": "X.y read from the SNP-major copy only (no sample-major copy needed); this rank"}
        finally:
            _lib.check(L.bsg_set_prodvec_path(0))

    # ---------------- second headline metric: bed_randomSVD wall time (not part of `value`) ----------------
    svd_info = None
    if not args.no_svd:
        k = 10 if args.workload == "cfg2" else 20
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if world > 1:
            from bigsnpr_b200.dist import randomsvd_sharded

            sv = randomsvd_sharded(g, int(tg.item() / n), k=k)
        else:
            sv = B.bed_randomSVD(g, k=k)
        tsvd = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tsvd, op=dist.ReduceOp.MAX)
        svd_info = {"k": k, "tol": 1e-4, "wall_s": float(tsvd.item()), "nops": sv["nops"], "niter": sv["niter"],
                    "d_top3": [float(v) for v in sv["d"][:3]],
                    "packed_bytes_read": float(alg_bytes) * (2 * sv["nops"] + k + 1),
                    "note": "bed_randomSVD(fun.scaling = bed_scaleBinom), Lanczos on the device; wall time includes "
                            "the scaling pass, the iteration and the k products for v"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle import ref

            r0, threads = 0.0, 1
            for t in host_thread_candidates():
                rr, _, _ = cpu_port_rate(n, 20000, SEED, steps=2, warmup=2, threads=t)
                if rr > r0:
                    r0, threads = rr, t
            m_s = int(max(20000, min(m_loc, r0 * 12.0 / n)))  # ~12 s of CPU work (capped at the whole matrix)
            r1, sec, _ = cpu_port_rate(n, m_s, SEED, steps=1, warmup=1, threads=threads)
            cpu_baseline = {"value": r1, "unit": "genotypes/s", "cores": threads, "kind": "port",
                            "sample": "bed_pMatVec4 port (oracle/bsg_oracle.c, -O2 -fopenmp) on the first %d of %d "
                                      "columns, %d samples, %.1f s" % (m_s, m_loc, n, sec)}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "error": repr(e)}

    if rank == 0:
        line = {
            "metric": "genotypes/sec in bed_prodVec", "value": value, "unit": "genotypes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["name"], "n": n, "m_total": int(tg.item() / n), "m_per_gpu": m_loc,
                       "na_rate": args.na_rate, "layouts": layouts,
                       "l2": "inputs larger than L2: %.2f GB of packed genotypes per pass per GPU vs 126 MB L2"
                             % (alg_bytes / 1e9),
                       "arithmetic": "exact int8 x uint2 on the integer tensor pipe, 61-bit fixed-point vector, fp64 epilogue",
                       "parallelism": "snp-column shards, 1 NCCL all-reduce of n doubles per step" if world > 1 else "1 GPU"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_val, "unit": "genotypes/s", "h2d_bytes_per_step": 3 * 8 * m_loc,
                    "d2h_bytes_per_step": 8 * n, "steps": e2e_steps, "max_abs_diff_vs_resident": chk},
            "clocks": clocks, "gpu_launches": launches, "svd": svd_info, "single_copy": single_copy,
        }
        print(json.dumps(line), flush=True)
    view.close()
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
