#!/usr/bin/env python
"""bench.py -- sweeps/sec of the auxiliary-field QMC sweep on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle (restatement of ALF, not ALF.out)

A "step" is ONE SWEEP (up + down pass over all L_trot slices, main.F90:714-887, plus TAU_M: --ltau 1 is the default because
BASELINE.json's headline configuration measures time-displaced Green functions) of every
chain resident on the GPU.  value = chain-sweeps per second summed over all ranks, timed with CUDA events on the
handle's stream, max over ranks.  e2e = the same through alf_b200_sweep_host (host buffers in/out each step).
The oracle is used here only as the CPU baseline (cpu_baseline / --impl reference), never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sweeps/sec on 16x16 Hubbard beta=10"


def metric_name(workload):
    """BASELINE.json's metric for the default workload; the other configurations are named after themselves."""
    return METRIC if workload == "hubbard_16x16_beta10" else "sweeps/sec on " + workload
UNIT = "sweeps/s"
WORKLOADS = {
    # name: (L1, L2, beta, dtau, U, nwrap, default chains per GPU)
    "hubbard_16x16_beta10": (16, 16, 10.0, 0.1, 4.0, 10, 148),      # BASELINE.json configs[2] = the metric's configuration
    "hubbard_8x8_beta10": (8, 8, 10.0, 0.1, 4.0, 10, 296),          # configs[1]
    "hubbard_4x4_beta5": (4, 4, 5.0, 0.1, 4.0, 10, 296),            # configs[0]
    "kondo_12x12_beta20": (12, 12, 20.0, 0.1, None, 5, 148),        # configs[3]: SU(2) Kondo lattice, N_dim = 288, complex, Nwrap = 5
    "z2_matter_12x12": (12, 12, 10.0, 0.1, None, 10, 148),          # configs[4]: Z2 gauge + matter, projective (theta = 10, Ltrot = 300), N_Global_tau = 36; use --ltau 0 --steps 1 --warmup 1
}


def make_model(name):
    from alf_b200.model import hubbard_square, kondo_square
    L1, L2, beta, dtau, U, nwrap, chains = WORKLOADS[name]
    if name.startswith("kondo"):
        return kondo_square(L1, L2, beta=beta, dtau=dtau), nwrap, chains
    if name.startswith("z2_matter"):
        from alf_b200.model import z2_matter_square
        return z2_matter_square(L1, L2, beta=beta, dtau=dtau, projector=True, theta=10.0), nwrap, chains
    return hubbard_square(L1, L2, beta=beta, dtau=dtau, U=U), nwrap, chains


def chain_seed(global_chain):
    """Deterministic stand-in for successive lines of the `seeds` file (Prog/Set_random_mod.F90:79-84)."""
    x = (global_chain + 1) * 2654435761 % 2147483647
    return int(x) or 1


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(workload, ltau, cores, seconds_target, steps, warmup):
    """Times the oracle (one independent chain per host thread, OPENBLAS_NUM_THREADS=1 as Documentation/running.tex:352
    advises).  A step is ONE WHOLE SWEEP of every chain when (steps + warmup) sweeps fit the time target (config 3: about 15-25 s per
    sweep); otherwise a step is a bounded sample: the first `seg` of the sweep's `nseg` stabilisation intervals (they must run in order),
    and the rate is extrapolated by seg / nseg (the intervals of the up pass, the down pass and TAU_M differ in cost, so the
    extrapolation is only approximate -- the line says so).  Returns per-step seconds, seg, nseg and the oracle's precision monitors."""
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle.oracle import Oracle
    model, nwrap, _ = make_model(workload)
    orcs = []
    for c in range(cores):
        o = Oracle(model, nwrap=nwrap); o.ranset(chain_seed(c)); o.fields_set(); orcs.append(o)

    def par(fn):
        th = [threading.Thread(target=fn, args=(o,)) for o in orcs]
        [t.start() for t in th]; [t.join() for t in th]
    par(lambda o: o.init())
    nseg = orcs[0].n_segments(ltau)
    pos = [0]

    def run_segments(k):
        lo = pos[0]
        def work(o):
            for i in range(lo, lo + k):
                o.sweep_segment(i % nseg, ltau)
        par(work); pos[0] = lo + k
    t0 = time.perf_counter(); run_segments(1); t_seg = time.perf_counter() - t0        # calibration on the first interval of the up pass (also a warm-up)
    # TAU_M intervals cost about 3x an up/down interval (CGR2_2 on the 2N x 2N system): estimate of a whole sweep from the calibration
    est_sweep = t_seg * (nseg if not ltau else 2 * nseg / 3 + 3 * nseg / 3)
    if est_sweep * (steps + warmup + 1) <= seconds_target:
        run_segments(nseg - 1)                                                        # finish the calibration sweep
        seg = nseg
    else:
        seg = max(1, min(nseg, int(seconds_target / max(t_seg, 1e-6) / max(1, steps + warmup))))
        if seg < nseg:
            pos[0] = 0                                                                # samples restart at the first interval (intervals run in order)
            for o in orcs:
                o.init()
    for _ in range(warmup):
        run_segments(seg)
        if seg < nseg:
            pos[0] = 0
    times = []
    for _ in range(steps):
        if seg < nseg:
            pos[0] = 0
        t0 = time.perf_counter(); run_segments(seg); times.append(time.perf_counter() - t0)
    prec = None
    try:
        cs = [o.control() for o in orcs]
        prec = {"green_max": max(c["XMAXG"] for c in cs), "green_mean": sum(c["XMEANG"] for c in cs) / max(1.0, sum(c["NCG"] for c in cs)),
                "tau_max": max(c["XMAX_tau"] for c in cs), "tau_mean": sum(c["XMEAN_tau"] for c in cs) / max(1.0, sum(c["NCG_tau"] for c in cs))}
    except Exception:
        pass
    return times, seg, nseg, prec


def _cpu_sample_text(seg, nseg, cores):
    what = "whole sweeps (all %d stabilisation intervals)" % nseg if seg == nseg else \
        "the first %d of %d stabilisation intervals of one sweep per step, rate extrapolated by %d/%d (intervals differ in cost: approximate)" % (seg, nseg, seg, nseg)
    return (f"{what}; {cores} independent chains, 1 per host thread, OPENBLAS_NUM_THREADS=1; oracle-CPU = C++ restatement of ALF (not ALF.out) built "
            "-O3 -ffast-math like ALF's GNU build, same LAPACK/BLAS calls (scipy OpenBLAS); it computes in complex f64 as ALF does; for real-valued models (Hubbard Mz) the GPU arm "
            "uses real f64 (SURVEY F6), which is worth 2-4x of the ratio")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = args.cpu_cores or (os.cpu_count() or 1)
    times, seg, nseg, prec = cpu_sample(args.workload, args.ltau, cores, 240.0, args.steps, args.warmup)
    tot = sum(times)
    value = cores * len(times) * (seg / nseg) / tot
    sample = _cpu_sample_text(seg, nseg, cores)
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (complex f64, as ALF)",
            "data": "synthetic", "config": {"workload": args.workload, "ltau": args.ltau, "chains": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "oracle-CPU: C++ restatement of ALF calling the same LAPACK/BLAS routines (scipy OpenBLAS); gfortran is absent so ALF.out cannot be built"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "precision": prec}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    # stdout carries exactly ONE JSON line: whatever libraries print there meanwhile (NCCL's version banner at 8 ranks) goes to stderr
    sys.stdout.flush(); saved_stdout = os.dup(1); os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from alf_b200 import build as _b
    if rank == 0:
        _b.build()
    if world > 1:
        dist.barrier()
    from alf_b200.api import AlfB200, fp64_peak
    from alf_b200.parallel import reduce_bins, init_comm

    model, nwrap, chains_default = make_model(args.workload)
    C = args.chains or chains_default                  # chains per handle
    H = max(1, args.handles)                           # handles (= CUDA streams, each driven by one host thread) per GPU
    gs, streams = [], []
    for k in range(H):
        gk = AlfB200(model, n_chains=C, nwrap=nwrap, device=local)
        gk.set_seeds([chain_seed((rank * H + k) * C + c) for c in range(C)])
        gk.fields_set(); gk.init_sweep()
        gs.append(gk); streams.append(torch.cuda.ExternalStream(gk.stream_ptr(), device=torch.device("cuda", local)))
    g = gs[0]
    init_comm(gs, world, rank)                         # NCCL communicators of the C-ABI's bin reduction (one per handle)
    N, L, M, F = model.Ndim, model.Ltrot, model.n_opv, model.N_FL

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def on_all(fn):
        """fn(k) for every handle, one host thread per handle (the C-ABI is re-entrant across handles; ctypes drops the GIL)."""
        if H == 1:
            fn(0); return
        errs = []
        def wrap(k):
            try:
                torch.cuda.set_device(local); fn(k)
            except Exception as e:      # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=wrap, args=(k,)) for k in range(H)]
        [t.start() for t in th]; [t.join() for t in th]
        if errs:
            raise errs[0]

    def timed(fn):
        """Device time of fn over all handles: first start event to last end event (CUDA events on each handle's stream)."""
        barrier()
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(H)]; ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(H)]
        for k in range(H):
            ev0[k].record(streams[k])                  # all streams idle here
        def body(k):
            fn(k); ev1[k].record(streams[k])           # fn returns after its stream has drained
        on_all(body)
        barrier()
        return max_over_ranks(max(ev0[a].elapsed_time(ev1[b]) for a in range(H) for b in range(H)))

    obsert_on = bool(args.ltau and args.obs_tau and model.latt is not None and not getattr(model, "Projector", False))
    if obsert_on:      # BASELINE configs[2] is "with time-displaced Green/spin observables": the device-side ObserT is part of every timed sweep
        for gk in gs:
            gk.obs_tau_enable(True)
    on_all(lambda k: gs[k].sweep(args.warmup, args.ltau) if args.warmup > 0 else None)
    # ---- timed region 1: device-resident sweeps
    for gk in gs:
        gk.kernel_timing(1 << 0)                 # CUDA events around the dominant kernel (k_wrapgr_fast) only
    c0 = [gk.control() for gk in gs]
    clocks = ClockSampler(local); clocks.start()
    ms = timed(lambda k: [gs[k].sweep(1, args.ltau) for _ in range(args.steps)])
    clk = clocks.stop()
    stats_all = [gk.kernel_stats() for gk in gs]; c1 = [gk.control() for gk in gs]
    stats = {key: (sum(st[key][0] for st in stats_all), sum(st[key][1] for st in stats_all)) for key in stats_all[0]}
    for gk in gs:
        gk.kernel_timing(0)
    value = world * H * C * args.steps / (ms * 1e-3)

    # ---- timed region 2: end to end through the C-ABI with host buffers (fields up, sweep, fields + observables + control down)
    f_in = [gk.get_fields() for gk in gs]; f_out = [np.empty_like(x) for x in f_in]
    obs = [np.zeros(max(16, gk.obs_size())) for gk in gs]; ctl = [np.zeros(16) for _ in gs]
    red = reduce_bins(gs, world, 0, rank)              # untimed: the first NCCL collective of a communicator sets up its connections
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        on_all(lambda k: gs[k].sweep_host(1, args.ltau, f_in[k], f_out[k], obs[k], ctl[k]))
        red = reduce_bins(gs, world, 0, rank)          # alf_b200_reduce_bins: NCCL reduction of ALL bin accumulators on the device (replaces MPI_REDUCE, observables_mod.F90:425-438), then read on rank 0
        f_in, f_out = f_out, f_in
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * H * C * args.steps / e2e_s
    h2d = world * H * C * L * M                    # int8 per field through the pinned staging buffers, all ranks
    d2h = world * H * (C * L * M + 8 * (len(obs[0]) + 16))
    if red is not None and red.get("tau") is not None:      # rank 0 additionally reads the reduced time-displaced lattice bins (Green, SpinZ, SpinXY, Den + backgrounds)
        d2h += H * sum(int(np.asarray(x).nbytes) for x in red["tau"][:2])

    # ---- un-timed extra passes (rank-local, after both timed regions): per-category device time + algorithmic FP64 flops of the
    # dense kernels on ONE handle running alone (CUDA events around every launch perturb the step, so they are kept out of the timed
    # regions), and the same sweep without the time-displaced part for reference
    g.kernel_timing(0xff)
    g.sweep(1, args.ltau)
    cat_stats = g.kernel_stats(); cat_flops = g.kernel_flops()
    g.kernel_timing(0)
    without_obsert = None
    if obsert_on:           # the same sweep with the device-side ObserT switched off (TAU_M propagates and stabilises, nothing is measured)
        for gk in gs:
            gk.obs_tau_enable(False)
        on_all(lambda k: gs[k].sweep(1, args.ltau))
        ot_ms = timed(lambda k: gs[k].sweep(1, args.ltau))
        without_obsert = {"value": world * H * C / (ot_ms * 1e-3), "unit": UNIT, "ms_per_step": ot_ms, "steps": 1,
                          "note": "sweep + TAU_M without ObserT (no Hop_mod_Symm of GT0, G0T, G00, GTT, no Predefined_Obs_tau_* at the time points)"}
    eq_only = None
    if args.ltau:
        eq_ms = timed(lambda k: gs[k].sweep(1, 0))
        eq_only = {"value": world * H * C / (eq_ms * 1e-3), "unit": UNIT, "ms_per_step": eq_ms, "steps": 1, "note": "same chains, sweep without TAU_M (ltau = 0)"}

    # ---- roofline of the dominant kernel and CPU baseline (rank 0, N = 1 only for the latter)
    upd_ms, upd_n = stats["update"]
    acc = sum(b["ACC_up"] - a["ACC_up"] for a, b in zip(c0, c1))
    nprop = sum(b["NC_up"] - a["NC_up"] for a, b in zip(c0, c1))
    w = 16 if g.is_complex else 8
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dfma, dmma = fp64_peak(local)
        traffic = None; traffic_src = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture of this kernel
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_update_kernel.json")))
            if tj.get("workload") == args.workload and tj.get("chains") == C:
                traffic = float(tj["dram_bytes_per_launch"]); traffic_src = tj.get("source")
        except Exception:
            pass
        fp64_kernels = {}
        for cat in ("gemm", "qrp", "formq", "trsm"):
            cms, cn = cat_stats[cat]
            if cn and cms > 0:
                tf = cat_flops[cat] / (cms * 1e-3) / 1e12
                fp64_kernels[cat] = {"launches": cn, "ms": cms, "tflops": tf, "frac_of_dmma_peak": tf / dmma if dmma else None}
        # Dominant kernel = the slice kernel (k_wrapgr_fast / k_wrapgr).  Its algorithm keeps accepted flips as delayed rank-1 factors in shared
        # memory and rewrites G in HBM once per KD accepts ("flush"): algorithmic bytes per flush and chain = F * 2 * w * N^2 (G read + written).
        # The flush count is measured on the device (control["flushes"]).  DESIGN.md section 3 states both figures.
        flushes = sum(b["flushes"] - a["flushes"] for a, b in zip(c0, c1))
        alg_bytes_per_launch = (flushes * F * 2.0 * w * N * N) / max(upd_n, 1)
        ref_alg_bytes_per_launch = (acc * F * 2.0 * w * N * N) / max(upd_n, 1)    # the REFERENCE algorithm (one ZGERU pass over G per accepted flip and flavor)
        alg_flops_per_launch = (acc * F * 2.0 * (4 if g.is_complex else 1) * N * N) / max(upd_n, 1)
        avg_s = upd_ms * 1e-3 / max(upd_n, 1)
        achieved = alg_bytes_per_launch / avg_s / 1e9 if avg_s > 0 else 0.0
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_update_kernel.json")))
            if tj.get("workload") == args.workload and tj.get("chains") == C:
                traffic = float(tj["dram_bytes_per_launch"]); traffic_src = tj.get("source")
        except Exception:
            pass
        roof = {"kernel": "k_wrapgr_fast (delayed-update slice kernel)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches": upd_n, "avg_launch_ms": 1e3 * avg_s, "share_of_step": upd_ms / (H * ms) if ms > 0 else None,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "flushes_per_launch_and_chain": flushes / max(upd_n, 1) / C,
                "reference_algorithm_bytes_per_launch": ref_alg_bytes_per_launch,
                "note": "algorithmic bytes = (rank-KD rewrites of G counted on the device) x F x 2 w N^2; the reference algorithm (one rank-1 pass over G per accepted flip, SURVEY 8d) would move reference_algorithm_bytes_per_launch; the kernel is bound by the latency of the sequential Metropolis decisions, not by HBM",
                "fp64": {"achieved_tflops": alg_flops_per_launch / avg_s / 1e12 if avg_s > 0 else 0.0, "peak_tflops_dfma_measured": dfma, "peak_tflops_dmma_measured": dmma}}
        # one roofline entry per dense FP64 kernel category (north_star: FP64 pipe utilisation of the wrap / QR kernels): algorithmic flops counted by the
        # library per launch (alf_b200_get_kernel_flops), CUDA-event time of the launches (one handle alone), DMMA peak measured in this run
        rooflines = [roof]
        names = {"gemm": "k_gemm (ZGEMM/ZTRMM)", "qrp": "k_qrp_* (ZGEQP3)", "formq": "k_apply_q* (ZUNGQR/ZUNMQR)", "trsm": "k_trsm_blk (ZTRSM)"}
        dense_fl = dense_ms = 0.0
        for cat, kv in fp64_kernels.items():
            rooflines.append({"kernel": names[cat], "bound": "fp64-tensor (DMMA m8n8k4)", "achieved": kv["tflops"], "peak": dmma, "unit": "TFLOP/s", "frac": kv["frac_of_dmma_peak"],
                              "launches": kv["launches"], "ms_per_sweep": kv["ms"], "traffic": None})
            dense_fl += cat_flops[cat]; dense_ms += kv["ms"]
        dense_weighted = {"tflops": dense_fl / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else None, "frac_of_dmma_peak": (dense_fl / (dense_ms * 1e-3) / 1e12 / dmma) if dense_ms > 0 and dmma else None}
        cpu = None; cpu_prec = None
        if world == 1 and not args.no_cpu_baseline:
            cores = args.cpu_cores or (os.cpu_count() or 1)
            times, seg, nseg, cpu_prec = cpu_sample(args.workload, args.ltau, cores, 30.0, 1, 0)
            cpu = {"value": cores * (seg / nseg) / times[0], "unit": UNIT, "cores": cores, "kind": "port", "sample": _cpu_sample_text(seg, nseg, cores)}
        nl = sum(v[1] for v in stats.values())
        line = {"metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128" if g.is_complex else "f64", "data": "synthetic",
                "config": {"workload": args.workload, "N_dim": N, "L_trot": L, "N_FL": F, "nwrap": nwrap, "ltau": args.ltau, "chains_per_gpu": H * C, "chains": world * H * C,
                           "handles_per_gpu": H, "chains_per_handle": C,
                           "parallelism": f"chains sharded over {world} GPU(s) x {H} handle(s) (one CUDA stream and host thread each), no data-path collective",
                           "l2": "working set (G + UDV storage of all chains) exceeds L2"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": nl, "kernel_launches": {k: v[1] for k, v in stats.items()},
                "acceptance": acc / max(nprop, 1), "precision_green_max": max(c["XMAXG"] for c in c1),
                "precision": {"green_max": max(c["XMAXG"] for c in c1), "green_mean": sum(c["XMEANG"] for c in c1) / max(1.0, sum(c["NCG"] for c in c1)),
                              "tau_max": max(c["XMAX_tau"] for c in c1), "tau_mean": sum(c["XMEAN_tau"] for c in c1) / max(1.0, sum(c["NCG_tau"] for c in c1)),
                              "phase_max": max(c["XMAXP"] for c in c1), "oracle_cpu_sample": cpu_prec,
                              "note": "Control_PrecisionG / _tau / P accumulated over all sweeps since init (mean = sum / count, ALF's 'Precision Green Mean'); stabilization.tex:225 recommends mean <= 1e-8"},
                "clocks": clk, "roofline": roof, "rooflines": rooflines, "fp64_dense_flop_weighted": dense_weighted, "cpu_baseline": cpu,
                "fp64_kernels": fp64_kernels, "fp64_peak_measured": {"dfma_tflops": dfma, "dmma_tflops": dmma},
                "breakdown_ms_per_sweep": {k: round(v[0], 3) for k, v in cat_stats.items()}, "equal_time_only": eq_only, "device_obsert_in_value": obsert_on, "without_device_obsert": without_obsert}
    for gk in gs:
        gk.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.stdout.flush(); os.dup2(saved_stdout, 1); os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hubbard_16x16_beta10", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0, help="chains per handle (default: per workload)")
    ap.add_argument("--handles", type=int, default=2, help="handles per GPU, each with its own CUDA stream and host thread (measured: 2 x 148 chains "
                    "overlap the latency-bound kernels of one handle with the throughput-bound ones of the other, +7 %% over 1 x 148)")
    ap.add_argument("--ltau", type=int, default=1, help="1: the sweep includes TAU_M (BASELINE configs[2]: time-displaced Green functions); 0: equal-time only")
    ap.add_argument("--cpu-cores", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--obs-tau", type=int, default=1, help="also time one sweep with the device-side ObserT on (reported as with_device_obsert; not part of value)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
