#!/usr/bin/env python
"""bench.py -- WRMF-implicit ALS user-updates/sec at rank 128 (BASELINE.json metric).

A "step" is one USER half-iteration of implicit-feedback WRMF with the CG(3) solver over the whole
synthetic matrix: XtX = I'I + lambda*I, (basis change), per-row CG solve of every user row, and -- when
N > 1 -- the NCCL exchange that gives every rank the updated user factors.
Workload (BASELINE.json configs[2], the configuration the metric is quoted on; fits one GPU):
10M x 1M CSR, 80 nnz/row, rank 128, CG(3), lambda 0.1.  N > 1: the same 10M users row-sharded
(strong scaling), one process per GPU under torchrun.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line on rank 0.  `value`: device-resident throughput (CUDA events on the engine's
stream, max over ranks).  `e2e`: the same step through the stateless C-ABI call with HOST buffers
(pinned; CSR values as R's doubles), H2D and D2H inside the timed region.  `roofline`: the per-row CG
kernel against the measured HBM peak.  `cpu_baseline` / `--impl reference`: the reference's CPU path on
this box's host cores (the only place the oracle is executed, as the thing being timed as a baseline).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITEM_SCALE, ITEM_DECAY = 0.1, 0.5


def _metric_name():
    """BASELINE.json's metric string (the driver compares the line against it); the literal is the fallback."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "WRMF-implicit ALS user-updates/sec at rank=128"


METRIC = _metric_name()

def _wl(n_user, n_item, nnz, k, cg=3, lam=0.1, feedback="implicit", solver=1, col_dist=0, len_dist=0):
    return dict(n_user=n_user, n_item=n_item, nnz=nnz, k=k, cg=cg, lam=lam, feedback=feedback, solver=solver,
                col_dist=col_dist, len_dist=len_dist)


WORKLOADS = {
    "c3": _wl(10_000_000, 1_000_000, 80, 128),                       # BASELINE configs[2]: the metric's configuration
    "c3-small": _wl(1_000_000, 1_000_000, 80, 128),                  # ncu captures
    "c3-tiny": _wl(100_000, 50_000, 80, 128),                        # CPU-side dry runs
    "c4": _wl(10_000_000, 1_000_000, 80, 128, feedback="explicit"),  # BASELINE configs[3]: explicit feedback (MMMF)
    "c2": _wl(1_000_000, 100_000, 50, 64, solver=0),                 # BASELINE configs[1]: implicit, rank 64, Cholesky
    "c3-chol": _wl(1_000_000, 1_000_000, 80, 128, solver=0),         # transform_-shaped: C3's rows solved by Cholesky (a10)
    "c5": _wl(50_000_000, 5_000_000, 100, 256),                      # BASELINE configs[4]: rank 256, 8 GPUs (5e9 nnz: >= 4 ranks)
    "c5-slice": _wl(6_250_000, 5_000_000, 100, 256),                 # one rank's share of C5 at 8 GPUs, as a single-GPU run
    "c5-small": _wl(500_000, 5_000_000, 100, 256),                   # ncu captures
    "c3-k64": _wl(10_000_000, 1_000_000, 80, 64),                    # C3's shape at rank 64 (tile kernel, half-warp per gathered row)
    # robustness points: Zipf(1.0)-like item popularity; the same with log-normal row lengths (sigma 1, mean 80)
    "c3-zipf": _wl(10_000_000, 1_000_000, 80, 128, col_dist=1),
    "c3-ragged": _wl(10_000_000, 1_000_000, 80, 128, col_dist=1, len_dist=1),
    "c3-items-small": _wl(2_000_000, 200_000, 80, 128),              # ncu capture of the item half: 200 k item rows of ~800 entries
    "c3-ragged-small": _wl(1_000_000, 1_000_000, 80, 128, col_dist=1, len_dist=1),
}
# side workload "topk": MatrixFactorizationRecommender$predict's top_product
TOPK = dict(n_user=65536, n_item=1_000_000, rank=128, k=10, nnz=80)
BYTES_PER_ROW = lambda n, k: 4 * n * k + 8 * n + 4 + 4 * k + 4 * k   # 42,628 at n=80, k=128
FLOPS_PER_ROW = lambda n, k, s: (s + 1) * (4 * n * k + 2 * k * k)


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def traffic_per_launch(workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu capture
    recorded in profiles/traffic.json for this workload and GPU count; None when no capture exists."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return int(t[workload][str(world)]["bytes"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.rows, self.device, self.proc = [], device, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def host_csr(L, n_rows, n_item, nnz, seed, pinned=True, offset=0):
    alloc = L.pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dt))
    ptr = alloc((n_rows + 1,), np.int32)
    idx = alloc((n_rows * nnz,), np.int32)
    val = alloc((n_rows * nnz,), np.float64)   # R's @x is double
    L.check(L.lib().b200als_synth_csr_host(n_rows, n_item, nnz, seed, 0, offset, L.vp(ptr), L.vp(idx), None, L.vp(val)))
    return ptr, idx, val


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores.  Nothing of the
    product is imported or loaded here: inputs come from the oracle library's own generator."""
    if rank != 0:
        return
    import oracle
    wl = WORKLOADS[args.workload]
    n_user, n_item, nnz, k, cg, lam = (wl[x] for x in ("n_user", "n_item", "nnz", "k", "cg", "lam"))
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the baseline uses every host core it may run on
    threads = oracle.host_threads()
    sample = min(n_user, args.cpu_rows)
    ptr, idx, val = oracle.synth_csr(sample, n_item, nnz, 42, False, 0, True, threads)
    rng = np.random.default_rng(0)
    # same input distribution as the GPU arm (see main()): trained-like item factors, R-initialised user factors
    X = rng.standard_normal((n_item, k), dtype=np.float32) * (ITEM_SCALE * (1.0 + np.arange(k, dtype=np.float32)) ** -ITEM_DECAY)
    X = np.ascontiguousarray(X, dtype=np.float32)
    Y = (rng.standard_normal((sample, k), dtype=np.float32) / 100)
    G = oracle.gram(X, lam, threads)
    impls = ["oracle"] + (["ref"] if oracle.ref_available() else [])
    # which implementation is "the reference" here: oracle/_ref is the reference's own source but linked against
    # our naive mini_arma (no BLAS), so it understates what Armadillo+OpenBLAS would do; the port uses hand-written
    # AVX2 loops.  Report the FASTER of the two as the reference arm and give both numbers.
    speeds = {}
    for impl in impls:
        n_probe = min(sample, 100_000)
        best_t = 1e30
        for _rep in range(2):   # first pass warms caches / page tables
            Yp = Y[:n_probe].copy()
            t = time.perf_counter()
            oracle.als_implicit(ptr[:n_probe + 1], idx[:ptr[n_probe]], val[:ptr[n_probe]], X, Yp, G, lam, 1, cg, threads, impl=impl)
            best_t = min(best_t, time.perf_counter() - t)
        speeds[impl] = n_probe / best_t
    best = max(speeds, key=speeds.get)
    # bounded: the whole arm stays under ~90 s whatever --steps / --warmup the driver passes
    budget_s = 60.0
    est = sample / speeds[best]
    warmup = max(1, min(args.warmup, 1 if est > 2.0 else args.warmup))
    steps = max(1, min(args.steps, int(budget_s / max(est, 1e-3))))
    Y0 = Y.copy()
    for _ in range(warmup):
        oracle.als_implicit(ptr, idx, val, X, Y, G, lam, 1, cg, threads, impl=best)
    dt = 0.0
    for _ in range(steps):
        Y[:] = Y0                      # like the GPU arm: every step starts from the initialisation (not timed)
        t0 = time.perf_counter()
        oracle.als_implicit(ptr, idx, val, X, Y, G, lam, 1, cg, threads, impl=best)
        dt += time.perf_counter() - t0
    value = sample * steps / dt
    kind = "reference" if best == "ref" else "port"
    sample_desc = ("first %d rows of the %dx%d/%d-nnz CSR against the full item matrix, XtX precomputed (not timed), %d host threads "
                   "(OMP_NUM_THREADS in the environment was %s and is overridden by num_threads()); probe on 100k rows: %s"
                   % (sample, n_user, n_item, nnz, threads, os.environ.get("OMP_NUM_THREADS", "unset"),
                      ", ".join("%s %.0f rows/s" % (("oracle/_ref (reference source + mini_arma)" if i == "ref" else "oracle port (AVX2 loops)"), v) for i, v in speeds.items())))
    out = {"impl": "reference", "metric": METRIC, "value": value,
           "unit": "user-updates/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "%s: %dx%d CSR, %d nnz/row, WRMF implicit rank=%d CG(%d), lambda=%g; CPU sample of %d rows"
                      % (args.workload, n_user, n_item, nnz, k, cg, lam, sample)},
           "cpu_baseline": {"value": value, "unit": "user-updates/s", "cores": threads, "kind": kind, "sample": sample_desc},
           "e2e": {"value": value, "unit": "user-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_topk(args):
    """Side workload: top-10 of 65,536 users x 1 M items at rank 128 with each user's 80 interactions excluded,
    through b200als_top_product (host buffers, copies inside the timed region)."""
    from rsparse_b200 import _lib as L
    from rsparse_b200 import top_product
    import scipy.sparse as sp
    T = TOPK
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((T["n_user"], T["rank"]), dtype=np.float32) / 10)
    y = (rng.standard_normal((T["n_item"], T["rank"]), dtype=np.float32) / 10)
    ptr = np.empty(T["n_user"] + 1, np.int32)
    idx = np.empty(T["n_user"] * T["nnz"], np.int32)
    L.check(L.lib().b200als_synth_csr_host(T["n_user"], T["n_item"], T["nnz"], 42, 0, 0, L.vp(ptr), L.vp(idx), None, None))
    nr = sp.csr_matrix((np.ones(len(idx), np.float32), idx, ptr), shape=(T["n_user"], T["n_item"]))
    top_product(x[:4096], y, T["k"], nr[:4096], [])        # warm-up
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, sc = top_product(x, y, T["k"], nr, [])
    dt = (time.perf_counter() - t0) / args.steps
    flops = 2.0 * T["n_user"] * T["n_item"] * T["rank"]
    # CPU stand-in for the reference's loop on a sample of users: BLAS scores + partial sort (numpy), all cores
    m = 512
    t0 = time.perf_counter()
    s64 = x[:m].astype(np.float64) @ y.astype(np.float64).T
    s64[nr[:m].nonzero()] = -np.inf
    part = np.argpartition(-s64, T["k"], axis=1)[:, :T["k"]]
    cpu_dt = time.perf_counter() - t0
    same = float(np.mean([set(part[u]) == set(ids[u]) for u in range(m)]))
    out = {"metric": "top-k recommendation users/sec (top_product, k=10, rank=128, 1M items)", "value": T["n_user"] / dt,
           "unit": "users/s", "n_gpus": 1, "steps": args.steps, "warmup": 1, "ms_per_step": 1e3 * dt, "higher_is_better": True,
           "dtype": "f64 accumulate of f32 factors", "data": "synthetic",
           "config": {"workload": "topk: %(n_user)d users x %(n_item)d items, rank %(rank)d, k %(k)d, %(nnz)d excluded per user" % T},
           "roofline": {"bound": "fp64", "achieved": flops / dt / 1e12, "peak": 40.0, "unit": "TFLOP/s",
                        "frac": flops / dt / 1e12 / 40.0, "traffic": None, "peak_source": "nominal B200 fp64 (no measured value)"},
           "cpu_baseline": {"value": m / cpu_dt, "unit": "users/s", "cores": os.cpu_count(), "kind": "port",
                            "sample": "numpy: dgemm scores + argpartition on %d users (not the reference's heap loop)" % m,
                            "top_k_sets_equal_frac": same},
           "e2e": {"value": T["n_user"] / dt, "unit": "users/s", "h2d_bytes_per_step": int(x.nbytes + y.nbytes + ptr.nbytes + idx.nbytes),
                   "d2h_bytes_per_step": int(ids.size * 12)}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("B200ALS_WORKLOAD", "c3"), choices=sorted(WORKLOADS) + ["topk"])
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic; CG: 2 resident full-XtX, 3 resident eigenbasis, 10 shared-memory tile kernel only; Cholesky: 0 = tcgen05-Gram rows kernel (rank 128) / warp per system (rank 64), 4 FFMA2-Gram rows kernel")
    ap.add_argument("--stage", type=int, default=0, help="resident-kernel tile staging: 0 default, 1 cp.async.bulk, 2 cp.async")
    ap.add_argument("--ctas", type=int, default=0, help="CTAs per SM the kernel is built for: CG resident 0/3/4, rank-128 row-per-thread Cholesky 0/2/3")
    ap.add_argument("--gram", default="default", choices=["default", "bf16", "ffma", "tf32x3"], help="arithmetic of XtX (BASELINE configs[4]: fp32 vs tensor-core bf16 Gram)")
    ap.add_argument("--fit-iters", type=int, default=0, help="after the timed steps, run this many full ALS iterations (item half + user half, "
                    "R/model_WRMF.R:318-338) from R's initialisation and report the device time of every half-iteration (side workload)")
    ap.add_argument("--warm-eig", action="store_true", help="keep the warm start of the eigen-decomposition between steps (the bench repeats a "
                    "half-iteration against an UNCHANGED fixed matrix, which would make the decomposition free; off by default)")
    ap.add_argument("--half", default="users", choices=["users", "items"], help="which half-iteration is the step (items: the item-major orientation is built on the device(s) first)")
    ap.add_argument("--cpu-rows", type=int, default=2_000_000, help="rows in the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if not args.warm_eig:
        os.environ["B200ALS_EIG_WARM"] = "0"
    rank, world, local_rank = (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                               int(os.environ.get("LOCAL_RANK", "0")))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    from rsparse_b200 import parallel

    if args.workload == "topk":
        return run_topk(args)
    from rsparse_b200 import Session
    from rsparse_b200 import _lib as L
    wl = WORKLOADS[args.workload]
    n_user, n_item, nnz, k, cg, lam = (wl[x] for x in ("n_user", "n_item", "nnz", "k", "cg", "lam"))
    feedback, solver = wl["feedback"], wl["solver"]
    if L.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: libb200als.so has no CPU fallback")
    numa = parallel.bind_to_gpu_numa(local_rank) if world > 1 else "numa: single process, not bound"
    L.check(L.lib().b200als_set_device(local_rank))
    parallel.init_engine_comm()
    begin, end = parallel.shard_range(n_user, rank, world)
    n_local = end - begin
    items_half = (args.half == "items")
    main_metric = (args.workload == "c3" and not items_half and args.kernel == 0)
    if not (feedback == "implicit" and solver == L.CONJUGATE_GRADIENT and k == 128 and not items_half and wl["len_dist"] == 0):
        args.no_e2e = args.no_cpu = True     # side workloads: device-resident number only
    if n_user // world * nnz > 2**31 - 1:
        raise SystemExit("workload %s needs more GPUs: %d nnz per rank exceed the 32-bit row pointers of a shard" % (args.workload, n_user // world * nnz))
    s = Session.synthetic(n_local, begin, n_user, n_item, nnz, 42, k, feedback, solver, cg, True, lam,
                          args.kernel, args.stage, args.ctas, wl["col_dist"], wl["len_dist"],
                          {"default": 0, "bf16": 1, "ffma": 2, "tf32x3": 3}[args.gram])
    if items_half:
        s.build_missing_orientation()     # item-major orientation on the device(s): sharded transpose when N > 1
    # Inputs: the SOLVED side at R's initialisation scale N(0,1)/100 (R/model_WRMF.R:203-215); the FIXED side "trained-like":
    # N(0,1) * 0.1 * (1+f)^-0.5, so XtX has a ~rank:1 spectrum and CG takes all its steps.  (i.i.d. fixed factors
    # make XtX ~ c*I: every row then leaves the CG loop after ONE step through `rsnew < CG_TOL`,
    # wrmf_implicit.hpp:27, which would flatter the throughput by ~1.7x.)
    solved, fixed = (L.ITEMS, L.USERS) if items_half else (L.USERS, L.ITEMS)
    n_solved = n_item if items_half else n_user
    s.randomize_factors(fixed, 1234, ITEM_SCALE, ITEM_DECAY)
    s.randomize_factors(solved, 5678, 0.01, 0.0)

    # Every step starts from freshly initialised factors on the solved side (R's N(0,1)/100), like the first
    # half-iteration of a fit: repeating the half-iteration on its own output converges the rows, after which the
    # reference's `rsnew < CG_TOL` exit (wrmf_implicit.hpp:27) leaves CG after one step and the step gets ~1.7x cheaper.
    # The re-initialisation kernel runs between the timed brackets; each step is bracketed by CUDA events on the
    # engine's stream (b200als_timer_start/_stop) and the K device times are summed.
    for w_ in range(args.warmup):
        s.randomize_factors(solved, 900 + w_, 0.01, 0.0)
        s.half_iteration(solved)
    sampler = ClockSampler(local_rank)
    launches0 = L.lib().b200als_launch_count()
    parts = {"gram_ms": 0.0, "prep_ms": 0.0, "solve_ms": 0.0, "comm_ms": 0.0}
    loss = None
    ms_sum = 0.0
    sampler.start()
    for st_ in range(args.steps):
        s.randomize_factors(solved, 5678 + st_, 0.01, 0.0)
        parallel.barrier()
        L.check(L.lib().b200als_timer_start())
        loss = s.half_iteration(solved)
        ms = C.c_float(0)
        L.check(L.lib().b200als_timer_stop(C.byref(ms)))
        ms_sum += parallel.max_over_ranks(ms.value)
        for kk, v in s.last_timing().items():
            parts[kk] += v
    clocks = sampler.stop()
    exchange = s.exchange_mode()
    parallel.barrier()
    launches = int(L.lib().b200als_launch_count() - launches0) - args.steps   # minus the re-initialisation launches
    ms_total = ms_sum
    ms_per_step = ms_total / args.steps
    value = n_solved * args.steps / (ms_total / 1e3)

    # roofline of the solve kernel(s): algorithmic bytes (4nk + 8n + 4 + 4k + 4k per row, summed over the
    # rows this rank solves) / the solve phase's own device time
    hbm_peak, peak_src = measured_peaks()
    solve_ms = parallel.max_over_ranks(parts["solve_ms"]) / args.steps
    plan = s.row_plan(solved) if solver == L.CONJUGATE_GRADIENT else None
    nnz_local = s.row_plan(solved)["nnz_local"]
    rows_local = (parallel.shard_range(n_item, rank, world)[1] - parallel.shard_range(n_item, rank, world)[0]) if items_half else n_local
    algo_bytes = (4 * k + 8) * nnz_local + rows_local * (4 + 8 * k)
    achieved = algo_bytes / (solve_ms / 1e3) / 1e9
    nnz_avg = nnz_local / max(1, rows_local)
    if solver == L.CHOLESKY:   # compute-bound path: 2nk^2 + 2nk + k^3/3 + 2k^2 flop per row
        fl = 2 * nnz * k * k + 2 * nnz * k + k ** 3 / 3 + 2 * k * k
        roofline = {"bound": "tensor", "kernel": ("als_chol_rows_kernel<128,3,1> (row-per-thread panel Cholesky, per-row Gram on tcgen05 3xTF32)" if (k == 128 and args.kernel != 4)
                                                 else "als_chol_warp64_kernel (warp per system, fp32 FFMA2)" if (k == 64 and args.kernel != 4)
                                                 else "als_chol_rows_kernel (row-per-thread panel Cholesky, fp32 FFMA2 Gram)"),
                    "achieved": n_local * fl / (solve_ms / 1e3) / 1e12, "peak": None, "unit": "TFLOP/s", "frac": None,
                    "traffic": None, "kernel_ms": solve_ms}
    else:
        rows = plan["rows"]
        by = {"als_cg_resident_kernel": rows["resident"], "als_cg_tile_kernel": rows["tile_4cta_double"] + rows["tile_4cta_single"] + rows["tile_2cta_single"] + rows["tile_1cta_double"],
              "als_cg_tile_kernel (thread-block clusters)": rows["cluster2"] + rows["cluster4"] + rows["cluster8"],
              ("als_cg_gram_kernel (per-row Gram on tcgen05)" if (k == 128 and os.environ.get("B200ALS_GRAM_ROWS", "1") != "0") else "als_cg_generic_kernel"): rows["long"]}
        if args.kernel == 1:
            by = {"als_cg_generic_kernel": rows_local}
        dominant = max(by, key=by.get)
        roofline = {"bound": "hbm", "kernel": dominant, "rows_by_kernel": plan["rows"] if args.kernel != 1 else by,
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": traffic_per_launch(args.workload if not items_half else args.workload + "-items", world),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes), "kernel_ms": solve_ms,
                    "fp32_tflops": (cg + 1) * (4 * nnz_local * k + 2 * k * k * rows_local) / (solve_ms / 1e3) / 1e12}
    if solver == L.CHOLESKY:
        # rank 128: the per-row Gram (2nk^2 of the flops) runs on tcgen05 as 3xTF32 => effective tensor peak = TF32 dense / 3
        # = measured bf16 sustained / 2 / 3; rank 64: fp32 FFMA2 kernel => nominal fp32 peak (148 SMs x 128 lanes x 2 x clock)
        try:
            pk16 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
        except Exception:
            pk16 = 1364.4
        if k == 128 and args.kernel != 4:
            pk, src = pk16 / 6.0, "3xTF32 effective = measured bf16 sustained / 2 (TF32 rate) / 3 (split passes)"
        else:
            roofline["bound"] = "fp32"
            pk, src = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal fp32 FMA peak (148 SMs x 128 lanes x 2 flop x 1.965 GHz)"
        roofline["peak"], roofline["frac"], roofline["peak_source"] = pk, roofline["achieved"] / pk, src
        roofline["gram_share_of_flops"] = 2 * nnz * k * k / fl
    # ---- e2e: stateless C-ABI call, host buffers, copies inside the timed region -----------------------
    e2e = None
    if not args.no_e2e:
        from rsparse_b200 import als_implicit
        if world > 1:
            # every rank passes the same item matrix: rank 0 uploads it, the others receive it over NVLink
            os.environ["B200ALS_STATELESS_SHARE_FIXED"] = "1"
        ptr, idx, val = host_csr(L, n_local, n_item, nnz, 42, True, begin)
        Xh = L.pinned_empty((n_item, k), np.float32)
        Yh = L.pinned_empty((n_local, k), np.float32)
        L.check(L.lib().b200als_get_factors(s._h, L.ITEMS, L.vp(Xh)))
        s.randomize_factors(L.USERS, 4242, 0.01, 0.0)     # un-converged warm start (R's initialisation)
        Yfull = s.get_factors(L.USERS)
        Yh[:] = Yfull[begin:end]
        del Yfull
        Y0h = np.array(Yh, copy=True)
        als_implicit(ptr, idx, val, Xh, Yh, lam, L.CONJUGATE_GRADIENT, cg)   # warm-up (allocations, first touch)
        dt = 0.0
        for _ in range(args.e2e_steps):
            Yh[:] = Y0h                 # un-converged warm start for every step (host copy, not timed)
            parallel.barrier()
            t0 = time.perf_counter()
            als_implicit(ptr, idx, val, Xh, Yh, lam, L.CONJUGATE_GRADIENT, cg)
            dt += parallel.max_over_ranks(time.perf_counter() - t0)
        h2d = ptr.nbytes + idx.nbytes + val.nbytes + Xh.nbytes + Yh.nbytes      # rank 0 (the other ranks: without the item matrix when N > 1)
        e2e = {"value": n_user * args.e2e_steps / dt, "unit": "user-updates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(Yh.nbytes), "ms_per_step": 1e3 * dt / args.e2e_steps, "steps": args.e2e_steps,
               "h2d_gbs_per_rank": h2d / (dt / args.e2e_steps) / 1e9, "host_binding_rank0": numa,
               "fixed_matrix": "uploaded by every rank" if world == 1 else "uploaded by rank 0, ncclBroadcast to the others (B200ALS_STATELESS_SHARE_FIXED=1)",
               "api": "b200als_als_implicit_float (stateless, host pointers, CSR values double as in R's dgCMatrix)"}

    # ---- CPU baseline (rank 0, N = 1): the oracle timed on this box's host cores -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        threads = oracle.host_threads()
        sample = min(n_local, args.cpu_rows)
        if args.no_e2e:
            ptr, idx, val = host_csr(L, sample, n_item, nnz, 42, False, 0)
            Xh = s.get_factors(L.ITEMS)
        Yc = (np.random.default_rng(1).standard_normal((sample, k), dtype=np.float32) / 100)
        G = oracle.gram(np.asarray(Xh), lam, threads)
        nn = int(ptr[sample])
        t0 = time.perf_counter()
        oracle.als_implicit(ptr[:sample + 1], idx[:nn], val[:nn], np.asarray(Xh), Yc, G, lam, 1, cg, threads)
        dt = time.perf_counter() - t0
        # evidence that the workload exercises all CG steps: result of CG(3) vs CG(2) on 2000 rows
        m = min(2000, sample)
        Ya, Yb = Yc[:m].copy(), Yc[:m].copy()
        Ya[:] = 0.01
        Yb[:] = 0.01
        oracle.als_implicit(ptr[:m + 1], idx[:ptr[m]], val[:ptr[m]], np.asarray(Xh), Ya, G, lam, 1, cg, threads)
        oracle.als_implicit(ptr[:m + 1], idx[:ptr[m]], val[:ptr[m]], np.asarray(Xh), Yb, G, lam, 1, cg - 1, threads)
        step_effect = float(np.linalg.norm(Ya - Yb) / np.linalg.norm(Ya))
        cpu = {"value": sample / dt, "unit": "user-updates/s", "cores": threads, "kind": "port",
               "last_cg_step_changes_result_by_relF": step_effect,
               "sample": "oracle port (C++/OpenMP restatement of wrmf_implicit.hpp, AVX2) on the first %d rows of the "
                         "same CSR against the full item matrix, %d threads, XtX precomputed" % (sample, threads),
               "seconds": dt}

    # ---- full ALS iterations (item half + user half), device-resident: what a fit costs per iteration ----------
    fit = None
    if args.fit_iters > 0:
        os.environ.pop("B200ALS_EIG_WARM", None)      # a real fit: the eigen-decomposition is warm-started from the previous iteration
        s.build_missing_orientation()
        s.init_factors(7)                             # users N(0,1)/100, items zero (CG): R/model_WRMF.R:203-230
        fit = {"iterations": []}
        for it in range(args.fit_iters):
            row = {}
            for nm, side in (("items", L.ITEMS), ("users", L.USERS)):
                parallel.barrier()
                L.check(L.lib().b200als_timer_start())
                ls = s.half_iteration(side)
                ms = C.c_float(0)
                L.check(L.lib().b200als_timer_stop(C.byref(ms)))
                row[nm] = dict(ms=parallel.max_over_ranks(ms.value), loss=ls, **s.last_timing(), rows=s.row_plan(side)["rows"])
            fit["iterations"].append(row)
        last = fit["iterations"][-1]
        fit["ms_per_iteration_last"] = last["items"]["ms"] + last["users"]["ms"]

    if rank == 0:
        side = args.workload + (" items half" if items_half else "")
        out = {"metric": METRIC if (args.workload in ("c3", "c3-small", "c3-tiny") and not items_half) else "%s (side workload %s)" % (METRIC, side),
               "value": value, "unit": "item-updates/s" if items_half else "user-updates/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "%s: %dx%d CSR, %s nnz/row%s, WRMF %s rank=%d %s lambda=%g, %s half-iteration"
                                      % (args.workload, n_user, n_item, ("%d" % nnz) if not wl["len_dist"] else ("log-normal (sigma 1, mean %d)" % nnz),
                                         ", Zipf(1.0)-like item popularity" if wl["col_dist"] else "", feedback, k,
                                         "CG(%d)" % cg if solver == 1 else "Cholesky", lam, "item" if items_half else "user"),
                          "rows_solved_per_step": n_solved, "mean_nnz_per_solved_row_this_rank": nnz_avg,
                          "parallelism": "rows sharded over %d GPU(s), exchange of updated factors: %s" % (
                              world, {"none": "none (single GPU)", "p2p": "peer-memory pushes over NVLink (copy engines)",
                                      "nccl": "grouped NCCL broadcasts"}[exchange]),
                          "l2": "inputs larger than L2 (CSR %.1f GB + factors %.1f GB per step vs 126 MB L2); no flush needed"
                                % (n_local * nnz * 8 / 1e9, (n_local + n_item) * k * 4 / 1e9),
                          "kernel": args.kernel, "stage": args.stage, "ctas": args.ctas, "gram": args.gram, "loss": loss},
               "step_breakdown_ms": {kk: v / args.steps for kk, v in parts.items()},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if fit is not None:
            out["fit"] = fit
        print(json.dumps(out), flush=True)
    s.close()


if __name__ == "__main__":
    main()
