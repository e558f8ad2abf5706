#!/usr/bin/env python
"""bench.py — spin-updates/s of the IsingModel.jl spin-update hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at every N (weak scaling, replicas sharded, no data-path collective — SURVEY §8e): BASELINE.json
configs[1] = "SK dense Gaussian J, N=1024, 4096 replicas, Glauber sweeps with annealing schedule" per GPU.
One *step* = one full annealing run of --sweeps sequential sweeps (geometric schedule T 2.0 -> 0.05, one
temperature per sweep) of all 4096 replicas from the same random initial spins = 4096*1024*sweeps updates.

  value     : device-timed (CUDA events on the launching stream, per step, max over ranks), inputs resident in
              HBM, noise drawn by the in-kernel Philox RNG.
  e2e       : the same run through the public host API (isingmodel.jl_b200: SpinSystem / GlauberDynamics /
              SamplingHelper.run_), timed on the host with pinned buffers: H2D of the initial spins and the
              schedule, the sweeps, D2H of the final spins, energies and flip counts.
  roofline  : the sweep kernel against the measured HBM bandwidth, algorithmic bytes = (accepted flips) x N x 8 B
              (incremental-field formulation: a J row is consumed per accepted flip), per-chain
              accounting.  J rows are staged once per CTA in shared memory and shared by its chains, so the
              algorithmic rate can exceed what HBM alone could deliver; `traffic` is the DRAM traffic ncu saw.
  cpu_baseline / --impl reference : the CPU oracle (C restatement of the reference's algorithm: a full
              Float64 row dot per update; Julia is not installed here), on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_SITES, REPLICAS = 1024, 4096
T0, TF = 2.0, 0.05
SEED_J, SEED_S = 2, 3
METRIC, UNIT = "spin-updates/sec", "updates/s"


def workload(sweeps):
    from isingmodel_jl_b200 import synth
    J = synth.sk_J(N_SITES, SEED_J)
    h = np.zeros(N_SITES)
    T = synth.geometric_schedule(T0, TF, sweeps)
    return J, h, T


def config(args, extra=None):
    c = {"workload": "C2: Sherrington-Kirkpatrick dense Gaussian J, N=1024, 4096 replicas/GPU, Glauber "
                     f"sequential sweeps, geometric annealing T {T0}->{TF} over {args.sweeps} sweeps",
         "n_sites": N_SITES, "replicas_per_gpu": REPLICAS, "sweeps_per_step": args.sweeps,
         "updates_per_step_per_gpu": N_SITES * REPLICAS * args.sweeps, "sharding": "replicas (no collective)",
         "l2": "flushed between timed steps (256 MiB write); J (8 MiB) is L2/smem-resident by design"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.t_lo = self.t_hi = None

    def mark(self, lo=None, hi=None):
        """Bounds (time.time()) of the timed region: only samples taken inside it are reported."""
        if lo is not None:
            self.t_lo = lo
        if hi is not None:
            self.t_hi = hi

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if (self.t_lo is None or t >= self.t_lo) and (self.t_hi is None or t <= self.t_hi + 0.05)]
        scope = "timed region"
        if len(inside) < 3:  # a very short timed region: fall back to every sample under load (warm-up included)
            inside, scope = [r for _, r in self.rows], "warm-up + timed region"
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "scope": scope}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_rate(sweeps_total, seconds_target, threads=None):
    """Times the CPU oracle (one chain per thread, full Float64 row dot per update) on a bounded sample of
    the same workload; returns (updates/s, description, threads)."""
    import oracle
    from isingmodel_jl_b200 import synth
    threads = threads or oracle.num_threads()
    J, h, T = workload(sweeps_total)
    R = threads
    S0 = synth.spins(SEED_S, R, N_SITES)

    def run(nsw):
        nsteps = nsw * N_SITES
        fl = synth.logistic(5, (R, nsteps))
        t0 = time.perf_counter()
        oracle.ssf_run_batch(oracle.GLAUBER, J, h, S0, nsteps, fluct=fl, fluct_per_replica=True, T=T[:nsw],
                             steps_per_T=N_SITES, nthreads=threads)
        return time.perf_counter() - t0, nsteps * R

    dt, n = run(2)  # calibration (also warms the caches)
    nsw = int(max(2, min(sweeps_total, round(2 * seconds_target / max(dt, 1e-6)))))
    reps = int(max(1, round(seconds_target / max(dt * nsw / 2, 1e-6))))
    dt, n = 0.0, 0
    for _ in range(reps):
        d, m = run(nsw)
        dt, n = dt + d, n + m
    return n / dt, (f"{R} chains (one per thread) x first {nsw} sweeps of the C2 schedule x {reps} repeats "
                    f"({n} updates, {dt:.1f} s wall)"), threads


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rates = []
    desc, threads = "", 1
    for i in range(args.warmup + args.steps):
        rate, desc, threads = cpu_oracle_rate(args.sweeps, args.ref_seconds)
        if i >= args.warmup:
            rates.append(rate)
    v = float(np.mean(rates))
    upd = N_SITES * REPLICAS * args.sweeps
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * upd / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args, {"note": "CPU restatement of the reference algorithm (Julia unavailable); "
                                            "ms_per_step extrapolated from the bounded sample"}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0



# ---------------------------------------------------------------------------------------------- SCA workloads
def sca_workload(which):
    """(W, h_visible, b_hidden, R, T schedule factory, description) of BASELINE.json configs[2] / configs[3]."""
    from isingmodel_jl_b200 import synth
    if which == "c4":
        W, h, b = synth.bipartite_W(784, 512, 4, 0.1)
        return W, h, b, 16384, (lambda n: np.ones(n)), "C4: bipartite 784x512 block Gibbs (SCA), 16384 chains/GPU, T=1"
    N = 4096
    J = synth.sk_J(N, 3)
    # pinning parameter q = eigmax(J)/2 (demo.jl:82); W = (J + qI)/2, biases h/2 (h = 0 here)
    q = 0.5 * float(np.linalg.eigvalsh(J)[-1])
    W = 0.5 * (J + q * np.eye(N))
    z = np.zeros(N)
    return W, z, z, 8192, (lambda n: np.linspace(1.0, 0.05, n)), \
        f"C3: dense Gaussian J N=4096 MultiSpinFlip SCA (bipartite embedding W=(J+qI)/2, q={q:.4f}), 8192 replicas/GPU, linear annealing T 1->0.05"


def sca_cpu_rate(W, h, b, T, seconds_target):
    import oracle
    from isingmodel_jl_b200 import synth
    threads = oracle.num_threads()
    nv, nh = W.shape
    S0, T0 = synth.spins(7, threads, nv), synth.spins(8, threads, nh)

    def run(n):
        Fv, Fh = synth.logistic(9, (n, nv), 1), synth.logistic(9, (n, nh), 2)
        t0 = time.perf_counter()
        oracle.bip_run_batch(oracle.SCA, W, h, b, S0, T0, n, Fv, Fh, T[:n] if len(T) >= n else np.resize(T, n), nthreads=threads)
        return time.perf_counter() - t0, n * (nv + nh) * threads

    dt, cnt = run(1)
    n = int(max(1, min(200, round(seconds_target / max(dt, 1e-6)))))
    dt, cnt = run(n)
    return cnt / dt, f"{threads} chains (one per thread) x {n} SCA steps ({cnt} updates, {dt:.1f} s wall)", threads


def sca_main(args):
    which = args.workload
    prec_name = args.prec or "bf16x3"
    nst = args.sca_steps or (20 if which == "c3" else 200)
    W, h, b, R, sched, desc = sca_workload(which)
    nv, nh = W.shape
    T = sched(nst)
    upd_step = (nv + nh) * R * nst
    cfg = {"workload": desc + f", {nst} SCA steps per bench step", "nv": nv, "nh": nh, "chains_per_gpu": R,
           "sca_steps_per_step": nst, "updates_per_step_per_gpu": upd_step, "coupling_storage": prec_name,
           "sharding": "replicas (no collective)",
           "l2": "flushed between timed steps (256 MiB write); spin matrices exceed nothing by design: W is L2-resident"}
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        rates = []
        for i in range(args.warmup + args.steps):
            v, sd, cores = sca_cpu_rate(W, h, b, T, args.ref_seconds)
            if i >= args.warmup:
                rates.append(v)
        v = float(np.mean(rates))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * upd_step / v,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sd},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0

    import torch
    import torch.distributed as dist
    from isingmodel_jl_b200 import _lib, synth, sharding, SpinSystems, OnBipartiteGraph, SamplingHelper
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = _lib.context(local)
    ctx.set_stream(stream.cuda_stream)
    prec = {"bf16x3": _lib.PREC_BF16X3, "bf16x2": _lib.PREC_BF16X2, "bf16x1": _lib.PREC_BF16X1, "f64": _lib.PREC_F64,
            "fp16x2": _lib.PREC_FP16X2, "fp16x1": _lib.PREC_FP16X1}[prec_name]
    P = {"bf16x3": 3, "bf16x2": 2, "bf16x1": 1, "f64": 1, "fp16x2": 2, "fp16x1": 1}[prec_name]
    pv = torch.empty((R, nv), dtype=torch.int8).pin_memory().numpy()
    ph = torch.empty((R, nh), dtype=torch.int8).pin_memory().numpy()
    pv[:] = synth.spins(11 + 1000 * rank, R, nv)
    ph[:] = synth.spins(12 + 1000 * rank, R, nh)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ss = SpinSystems.SpinSystemOnBipartiteGraph(pv, ph, W, h, b, device=local, prec=prec)
    ua = OnBipartiteGraph.StochasticCellularAutomata(ss, float(T[0]))
    ens = ss._ensemble()

    def device_step(k):
        ens.set_spins(pv)
        ens.set_hidden(ph)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ens.bip_run(_lib.BIP_SCA, nst, seed=777 + rank, step_offset=k * nst, T=T)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ens.last_stats()

    def e2e_step(k):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ss.spinConfiguration = pv
        ss.hiddenLayer = ph
        SamplingHelper.run_(ua, nst, seed=777 + rank, step_offset=k * nst, temperatures=T)
        st = ens.last_stats()
        S, Tm = ss.spinConfiguration, ss.hiddenLayer
        E = SpinSystems.calcEnergy(ua)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt, pv.nbytes + ph.nbytes + st["h2d_bytes"], S.nbytes + Tm.nbytes + E.nbytes + st["d2h_bytes"], float(E.mean())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        device_step(k)
    barrier()
    sampler.mark(lo=time.time())
    ms, kms, launches = [], [], 0
    for k in range(args.steps):
        m, st = device_step(args.warmup + k)
        ms.append(m)
        kms.append(st["kernel_ms"])
        launches += st["launches"]
    barrier()
    sampler.mark(hi=time.time())
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sharding.max_over_ranks(sum(ms) / 1e3)
    for k in range(min(args.warmup, 1)):
        e2e_step(k)
    barrier()
    e2e_t = 0.0
    for k in range(args.steps):
        dt, h2d, d2h, Emean = e2e_step(args.warmup + k)
        e2e_t += dt
    barrier()
    t_e2e = sharding.max_over_ranks(e2e_t)
    total = upd_step * args.steps * world
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak_burst = None
        if os.path.exists(peaks_path):
            pk = json.load(open(peaks_path))
            peak, src = float(pk["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
            peak_burst = float(pk.get("bf16_tflops", 0.0)) or None
        else:
            peak, src = 1400.0, "B200_PROFILING.md fallback ~1.4 PFLOP/s sustained (of fallback)"
        n_half = 2 * nst
        kern_s = float(np.mean(kms)) / 1e3 / n_half          # average half-step (one GEMM + sample launch)
        alg = 2.0 * nv * nh * R                               # algorithmic flops per half-step launch
        ach = alg / kern_s / 1e12
        line = {"metric": METRIC, "value": total / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64" if prec_name == "f64" else prec_name[:4],
                "data": "synthetic", "config": cfg,
                "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                             "traffic": None, "peak_source": src, "kernel": "isb::bip_tc_kernel",
                             "kernel_ms": 1e3 * kern_s, "split_passes": P, "executed_TFLOPs": ach * P,
                             "executed_frac": ach * P / peak,
                             # the kernels are timed back to back inside a long step, so the sustained cuBLAS rate is the
                             # denominator; the burst rate (a GEMM timed alone) is given beside it
                             "peak_burst": peak_burst, "frac_of_burst": (ach / peak_burst) if peak_burst else None,
                             "accounting": "2 x N_out x N_in x R flop per half-step launch (one bf16 pass); executed = passes x algorithmic"},
                "e2e": {"value": total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "mean_final_energy": Emean},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            v, sd, cores = sca_cpu_rate(W, h, b, T, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sd}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0



# ---------------------------------------------------------------------------------------------- C5: row-sharded SCA
def c5_main(args):
    """BASELINE config 5: dense J with N = 8192 x GPUs (65536 on 8), rows of W = (J + qI)/2 sharded across the GPUs,
    R replicas, SCA annealing, one all-gather of the freshly sampled spin blocks per half-step (NCCL over NVLink)."""
    import torch
    import torch.distributed as dist
    from isingmodel_jl_b200 import _lib, synth, sharding, rowshard
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "config 5 (N=65536: J is 32 GiB in Float64) is GPU-only; see --workload c3 for the CPU arm of the same algorithm"}))
        return 0
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    prec_name = args.prec or "bf16x3"
    prec = {"bf16x3": _lib.PREC_BF16X3, "bf16x2": _lib.PREC_BF16X2, "bf16x1": _lib.PREC_BF16X1}[prec_name]
    P = {"bf16x3": 3, "bf16x2": 2, "bf16x1": 1}[prec_name]
    nb, R = args.c5_n_per_gpu, args.c5_replicas
    n = nb * world
    nst = args.sca_steps or 10
    sca = rowshard.RowShardedSCA(n, R, seed=5, q=1.0, prec=prec, device=local,
                                 exchange=os.environ.get("ISB_C5_EXCHANGE") or None)
    S0 = synth.spins(21, R, n)   # the same initial configuration on every rank
    T = np.linspace(1.0, 0.05, nst)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        sca.set_spins(S0)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sca.run(nst, T, seed=31, step_offset=k * nst)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        step(k)
    barrier()
    sampler.mark(lo=time.time())
    l0 = sca.launches
    ms = [step(args.warmup + k) for k in range(args.steps)]
    launches = sca.launches - l0
    barrier()
    sampler.mark(hi=time.time())
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sharding.max_over_ranks(sum(ms) / 1e3)
    # e2e: host spins in (pinned) -> run -> host spins out
    pin = torch.empty((R, n), dtype=torch.int8).pin_memory().numpy()
    pin[:] = S0
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        sca.set_spins(pin)
        sca.run(nst, T, seed=31, step_offset=(args.warmup + k) * nst)
        out = sca.get_spins()
    torch.cuda.synchronize()
    t_e2e = sharding.max_over_ranks(time.perf_counter() - t0)
    upd_step = 2 * n * R * nst
    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(pk))["bf16_tflops_sustained"]) if os.path.exists(pk) else 1400.0
        half_s = t_dev / args.steps / (2 * nst)
        flops = 2.0 * nb * n * R            # per GPU per half-step (one bf16 pass)
        gather = (world - 1) * R * nb * 2   # bytes received per GPU per half-step
        t_mma, t_link = flops * P / (peak * 1e12), gather / 770e9
        ach = flops / half_s / 1e12
        print(json.dumps({
            "metric": METRIC, "value": upd_step * args.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C5: dense SK J N={n} row-sharded over {world} GPU(s) ({nb} rows each), {R} replicas, "
                                   f"SCA annealing T 1->0.05, {nst} steps per bench step, all-gather of spins per half-step",
                       "n": n, "rows_per_gpu": nb, "replicas": R, "coupling_storage": prec_name,
                       "collective": {"pipelined": "ncclAllGather per half-step, hidden under the other replica group's GEMM",
                                      "nccl": "ncclAllGather per half-step (torch.distributed)",
                                      "copy": "copy-engine pushes into symmetric memory + barrier, hidden under the other replica group's GEMM",
                                      "fused": "peer stores fused into the sampling epilogue (symmetric memory) + barrier",
                                      "local": "none (1 GPU)"}[sca.exchange],
                       "l2": "flushed between timed steps; W block (>= 1 GiB) exceeds L2"},
            "roofline": {"bound": "tensor" if t_mma >= t_link else "nvlink", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "kernel": "isb::bip_tc_kernel",
                         "kernel_ms": 1e3 * half_s, "split_passes": P, "executed_frac": ach * P / peak,
                         "fused_target_ms": 1e3 * max(t_mma, t_link), "frac_of_fused_target": max(t_mma, t_link) / half_s,
                         "all_gather_bytes_per_half_step": gather},
            "e2e": {"value": upd_step * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(pin.nbytes),
                    "d2h_bytes_per_step": int(out.nbytes)},
            "gpu_launches": int(launches), "clocks": clocks}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0



# ---------------------------------------------------------------------------------------------- C1: 32x32 lattice
def c1_main(args):
    """BASELINE config 1: 2-D ferromagnetic 32x32 periodic lattice, Metropolis at T = 2.269, sequential sweeps, built
    from a sparse J exactly as the reference's tests / demo build theirs (neighbour-list kernel); 4096 replicas."""
    import scipy.sparse as sp
    from isingmodel_jl_b200 import synth
    N, R, T0 = 1024, 4096, 2.269
    sweeps = args.sweeps if args.sweeps != 1000 else 10000
    J = synth.lattice_J(32)
    upd_step = N * R * sweeps
    cfg = {"workload": f"C1: 32x32 periodic ferromagnet (sparse J, 4 neighbours), Metropolis at T={T0}, {sweeps} sequential "
                       f"sweeps, {R} replicas/GPU", "n_sites": N, "replicas_per_gpu": R, "sweeps_per_step": sweeps,
           "updates_per_step_per_gpu": upd_step, "sharding": "replicas (no collective)", "l2": "flushed between timed steps"}
    rank = int(os.environ.get("RANK", "0"))

    def cpu_rate(seconds):
        import oracle
        threads = oracle.num_threads()
        S0 = synth.spins(SEED_S, threads, N)

        def run(nsw):
            fl = synth.exponential(5, (threads, nsw * N))
            t0 = time.perf_counter()
            oracle.ssf_run_batch(oracle.METROPOLIS, J, np.zeros(N), S0, nsw * N, fluct=fl, fluct_per_replica=True,
                                 T=np.array([T0]), steps_per_T=nsw * N, nthreads=threads)
            return time.perf_counter() - t0

        dt = run(4)  # calibration
        nsw = int(max(4, min(4000, round(4 * seconds / max(dt, 1e-6)))))
        dt = run(nsw)
        return threads * nsw * N / dt, (f"{threads} chains (one per thread) x {nsw} sweeps ({dt:.1f} s wall), dense-row dot "
                                        "per update as the reference does"), threads

    if args.impl == "reference":
        if rank == 0:
            v, sd, cores = cpu_rate(args.ref_seconds)
            print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * upd_step / v,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                              "data": "synthetic", "config": cfg,
                              "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sd},
                              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}))
        return 0
    import torch
    import torch.distributed as dist
    from isingmodel_jl_b200 import _lib, sharding, SpinSystems, SingleSpinFlip, SamplingHelper
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = _lib.context(local)
    ctx.set_stream(stream.cuda_stream)
    pin = torch.empty((R, N), dtype=torch.int8).pin_memory().numpy()
    pin[:] = synth.spins(SEED_S + 1000 * rank, R, N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ss = SpinSystems.SpinSystem(pin, sp.csc_matrix(J), np.zeros(N), device=local)
    ua = SingleSpinFlip.MetropolisMethod(ss, T0)
    ens = ss._ensemble()
    nsteps = sweeps * N

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(k):
        ens.set_spins(pin)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ens.ssf_run(_lib.RULE_METROPOLIS, nsteps, seed=99 + rank, step_offset=k * nsteps, T=np.array([T0]), steps_per_T=nsteps)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ens.last_stats()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        device_step(k)
    barrier()
    sampler.mark(lo=time.time())
    ms, kms, flips, launches = [], [], [], 0
    for k in range(args.steps):
        m, st = device_step(args.warmup + k)
        ms.append(m)
        kms.append(st["kernel_ms"])
        flips.append(st["flips"])
        launches += st["launches"]
    barrier()
    sampler.mark(hi=time.time())
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sharding.max_over_ranks(sum(ms) / 1e3)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        ss.spinConfiguration = pin
        SamplingHelper.run_(ua, nsteps, order="sequential", seed=99 + rank, step_offset=(args.warmup + k) * nsteps,
                            temperatures=np.array([T0]), steps_per_T=nsteps)
        S = ss.spinConfiguration
        E = SpinSystems.calcEnergy(ua)
    torch.cuda.synchronize()
    t_e2e = sharding.max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6650.0
        kern_s = float(np.mean(kms)) / 1e3
        alg = float(np.mean(flips)) * 4 * 12.0   # accepted flips x 4 neighbours x (8 B coupling + 4 B index)
        line = {"metric": METRIC, "value": upd_step * args.steps * world / t_dev, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "roofline": {"bound": "hbm", "achieved": alg / kern_s / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / kern_s / 1e9 / peak, "traffic": None, "kernel": "isb::ssf_sparse_kernel",
                             "kernel_ms": 1e3 * kern_s, "accept_rate": float(np.mean(flips)) / upd_step,
                             "accounting": "accepted flips x 4 neighbours x 12 B; the neighbour-list kernel is bound by "
                                           "per-window noise / decision arithmetic and shared-memory latency, not by bytes"},
                "e2e": {"value": upd_step * args.steps * world / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(pin.nbytes),
                        "d2h_bytes_per_step": int(S.nbytes + E.nbytes), "mean_final_energy": float(E.mean())},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            v, sd, cores = cpu_rate(args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sd}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweeps", type=int, default=1000, help="annealing sweeps per step (SURVEY §8d C2: 1000)")
    ap.add_argument("--ref-seconds", type=float, default=8.0, help="CPU seconds per reference-arm step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU seconds of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prec", default=None,
                    help="c2: f64 | f32; c3/c4/c5: bf16x3 | bf16x2 | bf16x1 (c3/c4 also f64, fp16x2, fp16x1)")
    ap.add_argument("--c5-n-per-gpu", type=int, default=8192, help="c5: rows of J per GPU (N = this x GPUs; 8 GPUs -> 65536)")
    ap.add_argument("--c5-replicas", type=int, default=1024)
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c2 (default, the headline): SK N=1024 single-spin Glauber annealing; c3: dense N=4096 "
                         "MultiSpinFlip SCA, 8192 replicas; c4: bipartite 784x512 block Gibbs, 16384 chains")
    ap.add_argument("--sca-steps", type=int, default=None, help="SCA steps per bench step (c3: 20, c4: 200)")
    args = ap.parse_args()
    if args.workload == "c5":
        return c5_main(args)
    if args.workload == "c1":
        return c1_main(args)
    if args.workload != "c2":
        return sca_main(args)
    args.prec = args.prec or "f64"
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import isingmodel_jl_b200 as pkg
    from isingmodel_jl_b200 import _lib, synth, sharding, SpinSystems, SingleSpinFlip, SamplingHelper

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = _lib.context(local)
    ctx.set_stream(stream.cuda_stream)

    sweeps = args.sweeps
    nsteps = sweeps * N_SITES
    J, h, T = workload(sweeps)
    prec = _lib.PREC_F64 if args.prec == "f64" else _lib.PREC_F32
    bJ = 8 if args.prec == "f64" else 4
    # this rank's replicas: global replica ids [rank*R, (rank+1)*R) -> distinct initial spins and noise streams
    S0 = synth.spins(SEED_S + 1000 * rank, REPLICAS, N_SITES)
    pin = torch.empty((REPLICAS, N_SITES), dtype=torch.int8).pin_memory()
    S0p = pin.numpy()
    S0p[:] = S0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- public-API objects (also used for the device-resident arm: same ensemble)
    ss = SpinSystems.SpinSystem(S0p, J, h, device=local, prec=prec)
    ua = SingleSpinFlip.GlauberDynamics(ss, T0)
    ens = ss._ensemble()

    def device_step(k):
        ens.set_spins(S0p)                      # untimed reset to the initial configuration
        flush.zero_()                           # L2 flush
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ens.ssf_run(_lib.RULE_GLAUBER, nsteps, order=_lib.ORDER_SEQUENTIAL, seed=12345 + rank, step_offset=k * nsteps,
                    T=T, steps_per_T=N_SITES)
        e1.record(stream)
        torch.cuda.synchronize()
        st = ens.last_stats()
        return e0.elapsed_time(e1), st

    def e2e_step(k):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ss.spinConfiguration = S0p              # H2D (pinned)
        out = SamplingHelper.run_(ua, nsteps, order="sequential", seed=12345 + rank, step_offset=k * nsteps,
                                  temperatures=T, steps_per_T=N_SITES)
        st = ens.last_stats()
        S = ss.spinConfiguration                # D2H
        E = SpinSystems.calcEnergy(ua)          # D2H
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = S0p.nbytes + st["h2d_bytes"]
        d2h = S.nbytes + E.nbytes + st["d2h_bytes"]
        return dt, h2d, d2h, float(E.mean()), int(out["flips"].sum())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for k in range(args.warmup):
        device_step(k)
    barrier()
    sampler.mark(lo=time.time())
    ms, kms, flips, launches = [], [], [], 0
    for k in range(args.steps):
        m, st = device_step(args.warmup + k)
        ms.append(m)
        kms.append(st["kernel_ms"])
        flips.append(st["flips"])
        launches += st["launches"]
    barrier()
    sampler.mark(hi=time.time())
    clocks = sampler.stop() if rank == 0 else None
    t_dev = sharding.max_over_ranks(sum(ms) / 1e3)

    for k in range(min(args.warmup, 1)):
        e2e_step(k)
    barrier()
    e2e_t, h2d, d2h, Emean, fl2 = 0.0, 0, 0, 0.0, 0
    for k in range(args.steps):
        dt, h2d, d2h, Emean, fl2 = e2e_step(args.warmup + k)
        e2e_t += dt
    barrier()
    t_e2e = sharding.max_over_ranks(e2e_t)

    upd_step = N_SITES * REPLICAS * sweeps
    total_updates = upd_step * args.steps * world
    value = total_updates / t_dev
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        kern_s = float(np.mean(kms)) / 1e3
        alg_bytes = float(np.mean(flips)) * N_SITES * bJ
        achieved = alg_bytes / kern_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("ssf_kernel_dram_bytes_per_launch")
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        smem_peak = 128.0 * 148 * sm_mhz * 1e6 / 1e9  # GB/s: 128 B/clk/SM
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.prec, "data": "synthetic",
            "config": config(args),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "isb::ssf_kernel",
                         "kernel_ms": 1e3 * kern_s,
                         "accounting": f"accepted flips ({np.mean(flips):.4g}/launch) x N x {bJ} B per launch, per chain",
                         "accept_rate": float(np.mean(flips)) / upd_step,
                         "attempts_accounting_GBps": upd_step * N_SITES * bJ / kern_s / 1e9,
                         "smem_GBps": achieved, "smem_peak_GBps": smem_peak, "smem_frac": achieved / smem_peak},
            "e2e": {"value": total_updates / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "mean_final_energy": Emean},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, desc, cores = cpu_oracle_rate(sweeps, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
