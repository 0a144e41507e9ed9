#!/usr/bin/env python
"""bench.py -- BASELINE.json's configurations on B200.  Headline (default, `--config c2`):

    coordinate-descent restart-sweeps/s on Boolean least squares n=1000 / m_rows=1500 (dense P0, 1000 constraints
    x_i^2 = 1), 1024 SDR samples + COORD_DESCENT per GPU  ("configs[1]").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl reference] [--no-extras] [--no-cpu]

One STEP = one pass of the hot path over one batch:
  c2  SDR randomized rounding of 1024 draws per GPU (x = mu + z F, eval) -> improve_coord_descent of every draw (own MT19937 stream,
      reference defaults) -> best pick.  Weak scaling: 1024 restarts per GPU.
  c3  MAXCUT G(2000, 0.1): the same pipeline on 256 restarts SPLIT over the GPUs (strong scaling).
  c4  beamforming N=128, 32 constraints: improve_admm for the 16-value rho sweep, the rho values split over the GPUs.
  c5  circle packing, 200 circles (N=401, 20 701 constraints): suggest(RANDOM) + improve_coord_descent, 4096 restarts split over the GPUs.
Unit of work for c2/c3/c5 = one restart-sweep = n coordinate steps of one restart (one outer iteration of qcqp.py:110 / :160),
only EXECUTED steps counted; for c4 one ADMM iteration of one run.  N > 1 is launched by torchrun (one rank per GPU): no
collective on the data path, ONE all-gather per step -- inside the timed region -- carrying every rank's best (f0, maxviol, x).
The default run adds short measurements of c3, c4, c5 under "other_configs" (skip with --no-extras).
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

UNIT_CD = "restart-sweeps/s"
CONFIGS = {
    "c2": dict(kind="sdr_cd", gen="boolean_least_squares", gargs=dict(n=1000, m=1500, seed=1), restarts=1024, scaling="weak",
               metric="coord_descent_restart_sweeps_per_sec", unit=UNIT_CD,
               workload="boolean_least_squares n=1000 m_rows=1500 (dense P0, 1000 x_i^2==1 constraints), 1024 SDR samples + COORD_DESCENT per GPU"),
    "c3": dict(kind="sdr_cd", gen="maxcut", gargs=dict(n=2000, p=0.1, seed=1), restarts=256, scaling="strong",
               metric="coord_descent_restart_sweeps_per_sec", unit=UNIT_CD,
               workload="maxcut G(n=2000, p=0.1) (CSR objective, 2000 x_i^2==1 constraints), suggest(SDR) + COORD_DESCENT, 256 restarts split over the GPUs"),
    "c4": dict(kind="admm", gen="beamforming", gargs=dict(n=64, m=24, l=8, seed=1), restarts=16, scaling="strong",
               metric="admm_iterations_per_sec", unit="admm-iterations/s",
               workload="secondary_user_beamforming N=128 real variables / 32 dense rank-2 constraints, improve(ADMM), 16-value rho sweep split over the GPUs"),
    "c5": dict(kind="cd", gen="circle_packing", gargs=dict(ncirc=200), restarts=4096, scaling="strong",
               metric="coord_descent_restart_sweeps_per_sec", unit=UNIT_CD,
               workload="circle_packing 200 circles (N=401, 20701 constraints), suggest(RANDOM) + COORD_DESCENT, 4096 restarts split over the GPUs"),
}
STATIC_CD = dict(num_iters=1000, viol_tol=1e-2, tol=1e-4, rng="MT19937 stream per restart (np.random.seed(1000 + r))")
STATS_DT = np.dtype([("s1", "<i8"), ("s2", "<i8"), ("u1", "<i8"), ("u2", "<i8"), ("w1", "<i4"), ("w2", "<i4"),
                     ("status", "<i4"), ("ran2", "<i4"), ("skip", "<i8")])
ADMM_DT = np.dtype([("p1", "<i4"), ("p2", "<i4"), ("calls", "<i8"), ("status", "<i4"), ("pad", "<i4")])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md, the clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# problem instances (pure NumPy; loaded WITHOUT importing the qcqp_b200 package so that the CPU arms do not map the product .so)
# ------------------------------------------------------------------------------------------------------------
def build_problem(cfg, want_sdr=True):
    """(forms, info, X*) of a configuration.  X* is the host solve of the SDP relaxation (qcqp_b200/relax.py), outside every timed
    region -- the reference caches it too (qcqp.py:390-395)."""
    from oracle import ref_python as rp
    pb, relax, fm = rp.plain_module("problems"), rp.plain_module("relax"), rp.plain_module("forms")
    forms, info = getattr(pb, cfg["gen"])(**cfg["gargs"])
    Xstar = None
    if cfg["kind"] == "sdr_cd" and want_sdr:
        Xstar, _bound = relax.solve_sdr(fm.QCQPForm.from_tuples(forms))
    return forms, info, Xstar


def sdr_factor_np(Xstar, eps=1e-8):
    """(mu, F): qcqp.py:394-395 and the factor np.random.multivariate_normal builds from Sigma (row-broadcast mu*mu.T kept)."""
    Xs = np.asarray(Xstar, dtype=np.float64)
    n = Xs.shape[0] - 1
    mu = Xs[:-1, -1].copy()
    Sigma = Xs[:-1, :-1] - mu * mu.T + eps * np.eye(n)
    _u, s, vt = np.linalg.svd(Sigma)
    return mu, np.ascontiguousarray(np.sqrt(s)[:, None] * vt)


def shard(total, rank, world):
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def starts_for(cfg, forms, Xstar, lo, hi, rank, weak):
    """Per-rank inputs of a CD configuration: the standard normals of its draws (sdr_cd) or its suggest(RANDOM) points (cd), and the
    np.random.seed value of every restart.  Strong scaling slices ONE batch (the result does not depend on the number of GPUs)."""
    n = int(np.asarray(forms[0][1]).size)
    R = hi - lo
    if cfg["kind"] == "sdr_cd":
        if weak:
            Z = np.random.RandomState(2 + rank).standard_normal((R, n))
            seeds = 1000 + rank * R + np.arange(R)
        else:
            Z = np.random.RandomState(2).standard_normal((cfg["restarts"], n))[lo:hi]
            seeds = 1000 + np.arange(lo, hi)
        return np.ascontiguousarray(Z), seeds.astype(np.uint32)
    X0 = np.stack([np.random.RandomState(int(r)).randn(n) for r in range(lo, hi)]) if R else np.zeros((0, n))   # qcqp.py:382
    return np.ascontiguousarray(X0), np.arange(lo, hi).astype(np.uint32)


# ------------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified Python reference (baseline/_ref) and the C port (oracle/), never the product library
# ------------------------------------------------------------------------------------------------------------
def port_sample(cfg, forms, Xstar, restarts, threads, fast=True, num_iters=1000):
    """`restarts` restarts of the configuration through oracle/qcqp_oracle.c: (units of work, seconds)."""
    from oracle import oracle as orc
    P = orc.Problem(forms)
    n = P.n
    t0 = time.perf_counter()
    if cfg["kind"] == "admm":
        rhos = np.sqrt(32) * 2.0 ** (np.arange(-8, 8) / 2.0)
        np.random.seed(4)
        X0 = 2 * np.random.randn(1, n)
        P.compute_eig()
        t0 = time.perf_counter()
        _X, _f, _v, st = P.improve_admm_batch(X0, rhos[:restarts], nthreads=threads)
        return float(sum(s.iters_p1 + s.iters_p2 for s in st)), time.perf_counter() - t0
    Z, seeds = starts_for(cfg, forms, Xstar, 0, restarts, 0, False)
    if cfg["kind"] == "sdr_cd":
        mu, F = sdr_factor_np(Xstar)
        X0, _f, _v = P.sdr_sample_eval(mu, F, Z, nthreads=threads)
    else:
        X0 = Z
    rngs = (orc.RngState * restarts)()
    for r in range(restarts):
        rngs[r] = orc.RngState.from_seed(int(seeds[r]))
    _X, _f0, _mv, st = P.improve_cd_batch(X0, rngs, fast=fast, nthreads=threads, num_iters=num_iters)
    dt = time.perf_counter() - t0
    return sum(s.steps_p1 + s.steps_p2 for s in st) / float(n), dt


PORT_SAMPLE = {"c2": 1024, "c3": 64, "c4": 16, "c5": 64}     # restarts per step of the C-port arm (c2: the whole batch)


def reference_windows(cfg, procs, steps_per_proc, pool=None):
    """The UNMODIFIED Python reference (baseline/_ref/qcqp, else /root/reference) on `procs` worker processes: each runs a window of
    `steps_per_proc` coordinate steps of coord_descent_phase2 (Boolean-type problems) on its own restart.  restart-sweeps/s."""
    from oracle import ref_python as rp
    own = pool is None
    if own:
        pool = rp.ReferencePool(cfg["gen"], cfg["gargs"], procs=procs)
    steps, wall, _secs = pool.window(steps_per_proc)
    if own:
        pool.close()
    n = cfg["gargs"].get("n", 1)
    return steps / float(n) / wall, wall, steps


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores.  For the Boolean-type CD
    configurations (c2, c3) the line's value is the unmodified Python reference, timed over bounded windows of its own loop (one
    worker process per core); the C port of the same algorithm is reported beside it.  c4 / c5: the C port (kind 'port')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    from oracle import oracle as orc
    from oracle import ref_python as rp
    cores = os.cpu_count() or 1
    forms, _info, Xstar = build_problem(cfg)
    line = {"impl": "reference", "metric": cfg["metric"], "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": static_config(cfg), "gpu_launches": 0}
    pthreads = orc.lib().orc_max_threads()
    pr = PORT_SAMPLE[args.config]
    for _ in range(min(args.warmup, 1)):
        port_sample(cfg, forms, Xstar, min(pr, 2 * pthreads), pthreads)
    w_port, s_port = 0.0, 0.0
    for _ in range(max(1, min(args.steps, 3))):
        wk, dt = port_sample(cfg, forms, Xstar, pr, pthreads)
        w_port += wk; s_port += dt
    port = {"value": w_port / s_port, "unit": cfg["unit"], "cores": pthreads, "kind": "port",
            "sample": "%d restarts per step, complete runs, oracle/qcqp_oracle.c in cached-f mode (an algorithmic improvement over the "
                      "reference's O(n nnz) get_onevar_func), %d threads" % (pr, pthreads)}
    if cfg["kind"] == "sdr_cd" and rp.ref_root() is not None:
        K = 8 if args.config == "c2" else 2                # coordinate steps per worker and window: ~1 s (c2) / ~2.5 s (c3) of CPU each
        pool = rp.ReferencePool(cfg["gen"], cfg["gargs"], procs=cores)
        for _ in range(min(args.warmup, 1)):
            pool.window(1)
        tot_steps, tot_wall = 0, 0.0
        for _ in range(args.steps):
            _v, wall, st = reference_windows(cfg, cores, K, pool)
            tot_steps += st; tot_wall += wall
        pool.close()
        n = cfg["gargs"]["n"]
        value = tot_steps / float(n) / tot_wall
        line.update(value=value, ms_per_step=1e3 * tot_wall / args.steps,
                    cpu_baseline={"value": value, "unit": cfg["unit"], "cores": cores, "kind": "reference",
                                  "sample": "unmodified cvxgrp/qcqp (baseline/_ref) behind a stub cvxpy: per step every one of %d worker processes "
                                            "runs %d coordinate steps of coord_descent_phase2 (qcqp.py:152-178) on its own restart of the "
                                            "configuration; restart-sweeps/s = steps / n / wall of the slowest worker, i.e. extrapolated from "
                                            "windows (a whole sweep costs ~%d s per core)" % (cores, K, int(cores / max(value, 1e-12))),
                                  "reference_root": rp.ref_root(), "c_port": port})
    else:
        line.update(value=port["value"], ms_per_step=1e3 * s_port / max(1, min(args.steps, 3)), cpu_baseline=port)
    line["e2e"] = {"value": line["value"], "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))


def static_config(cfg):
    """The part of `config` that both arms print identically."""
    c = {"workload": cfg["workload"], "restarts": cfg["restarts"], "restarts_are": "per GPU" if cfg["scaling"] == "weak" else "total, split over the GPUs"}
    c.update(cfg["gargs"])
    if cfg["kind"] != "admm":
        c.update(STATIC_CD)
    else:
        c.update(num_iters=1000, tol=1e-2, viol_lim=1e4, rhos="sqrt(32) 2^(k/2), k = -8..7", start="2 randn(128) after seed(4)")
    return c


# ------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from qcqp_b200 import _lib
        self.lib = _lib
        self.L = _lib.load()
        self.flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=self.dev)    # > 126 MB L2
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def sum_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]


class CrossRankBest:
    """The only collective, enqueued on the launching stream so that it sits INSIDE the CUDA-event bracket: every rank contributes its
    best (f0, maxviol, x); one all-gather; the `better` pick over the N entries by the library's own best kernel; no host sync."""

    def __init__(self, env, n):
        t = env.torch
        self.env, self.n = env, n
        self.send = t.zeros(n + 2, dtype=t.float64, device=env.dev)
        self.recv = t.zeros(env.world * (n + 2), dtype=t.float64, device=env.dev)
        self.win = t.zeros(1, dtype=t.int32, device=env.dev)
        self.xwin = t.zeros(n, dtype=t.float64, device=env.dev)

    def __call__(self, d_f, d_v, d_X, d_best):
        env, t = self.env, self.env.torch
        idx = d_best.to(t.long)
        self.send[0:1].copy_(d_f.index_select(0, idx)); self.send[1:2].copy_(d_v.index_select(0, idx))
        self.send[2:].copy_(d_X.index_select(0, idx)[0])
        if env.world == 1:
            self.xwin.copy_(self.send[2:])
            return
        env.dist.all_gather_into_tensor(self.recv, self.send)
        tab = self.recv.view(env.world, self.n + 2)
        fcol, vcol = tab[:, 0].contiguous(), tab[:, 1].contiguous()
        env.lib.check(env.L.qcqp_best_device(fcol.data_ptr(), vcol.data_ptr(), env.world, 1e-4, self.win.data_ptr(), None, None, env.stream))
        self.xwin.copy_(tab.index_select(0, self.win.to(t.long))[0, 2:])


def timed_steps(env, step, steps, warmup, n_events, min_warmup=3):
    """W untimed + K timed steps; per step CUDA events on the launching stream, L2 flushed between iterations outside the brackets."""
    t = env.torch
    for _ in range(max(warmup, min_warmup)):
        env.flush.zero_(); step(None)
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if env.rank == 0:
        sampler.start()
    rows = []
    env.barrier()
    wall0 = time.perf_counter()
    for _ in range(steps):
        env.flush.zero_()
        ev = [t.cuda.Event(enable_timing=True) for _ in range(n_events)]
        step(ev)
        t.cuda.synchronize()
        rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(n_events - 1)] + [ev[0].elapsed_time(ev[-1])])
    env.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if env.rank == 0 else None
    return np.mean(np.array(rows), axis=0), wall, clocks


def probes(env):
    g, dm, df = C.c_double(0), C.c_double(0), C.c_double(0)
    env.lib.check(env.L.qcqp_probe_l2_bandwidth(48 << 20, 20, C.byref(g)))
    env.lib.check(env.L.qcqp_probe_fp64_peaks(C.byref(dm), C.byref(df)))
    return {"l2_read_GBps": g.value, "dmma_f64_TFLOPs": dm.value, "dfma_f64_TFLOPs": df.value,
            "how": "qcqp_probe_l2_bandwidth (48 MiB L2-resident buffer, 16-byte ld.global.cg, CUDA events) and qcqp_probe_fp64_peaks "
                   "(register-resident mma.sync.m8n8k4.f64 / DFMA chains), measured in this process before the timed region"}


def run_cd_config(env, key, steps, warmup, cpu_leg, light=False):
    """c2 / c3 (SDR draws + coordinate descent through qcqp_sdr_sample_eval_device + qcqp_cd_improve_device) and c5 (coordinate
    descent from suggest(RANDOM) points).  Returns the JSON line as a dict (rank 0) or None."""
    t, L, lib = env.torch, env.L, env.lib
    from qcqp_b200 import engine
    cfg = CONFIGS[key]
    weak = cfg["scaling"] == "weak"
    forms, _info, Xstar = build_problem(cfg)
    pack = engine.Pack(forms)
    n = pack.n
    lo, hi = (0, cfg["restarts"]) if weak else shard(cfg["restarts"], env.rank, env.world)
    R = hi - lo
    A, seeds = starts_for(cfg, forms, Xstar, lo, hi, env.rank, weak)
    sdr = cfg["kind"] == "sdr_cd"
    if sdr:
        mu, F = sdr_factor_np(Xstar)
        d_mu, d_F = t.from_numpy(mu).to(env.dev), t.from_numpy(F).to(env.dev)
    rng_host = engine.rng_states(seeds=[int(s) for s in seeds])
    d_rng0 = t.from_numpy(engine.rng_states_as_tensor_bytes(rng_host)).to(env.dev) if R else t.zeros(0, dtype=t.uint8, device=env.dev)
    d_rng = t.empty_like(d_rng0)
    d_A = t.from_numpy(A).to(env.dev)                       # standard normals (sdr_cd) or start points (cd), resident for `value`
    d_X0 = t.empty((R, n), dtype=t.float64, device=env.dev) if sdr else d_A
    d_X = t.empty((R, n), dtype=t.float64, device=env.dev)
    d_f = t.empty(R, dtype=t.float64, device=env.dev); d_v = t.empty_like(d_f); d_fs = t.empty_like(d_f); d_vs = t.empty_like(d_f)
    d_stats = t.zeros(R * C.sizeof(lib.CdStats), dtype=t.uint8, device=env.dev)
    d_best = t.zeros(1, dtype=t.int32, device=env.dev)
    prm = lib.CdParams(1000, 1e-2, 1e-4, 1, 0, 0)
    pick = CrossRankBest(env, n)

    def step(ev):
        d_rng.copy_(d_rng0)
        if ev: ev[0].record()
        if sdr:
            lib.check(L.qcqp_sdr_sample_eval_device(pack.handle, d_mu.data_ptr(), d_F.data_ptr(), d_A.data_ptr(), 0, R, d_X0.data_ptr(),
                                                     d_fs.data_ptr(), d_vs.data_ptr(), env.stream))
        if ev: ev[1].record()
        lib.check(L.qcqp_cd_improve_device(pack.handle, C.byref(prm), d_X0.data_ptr(), R, d_rng.data_ptr(), d_X.data_ptr(), d_f.data_ptr(),
                                            d_v.data_ptr(), d_stats.data_ptr(), env.stream))
        if ev: ev[2].record()
        lib.check(L.qcqp_best_device(d_f.data_ptr(), d_v.data_ptr(), R, 1e-4, d_best.data_ptr(), None, None, env.stream))
        pick(d_f, d_v, d_X, d_best)                         # the collective is inside the bracket
        if ev: ev[3].record()

    cd_parts = []
    part_buf = (C.c_double * 4)(); part_cnt = C.c_int32(0)

    def step_and_parts(ev):
        step(ev)
        if ev:
            t.cuda.synchronize()
            lib.check(L.qcqp_cd_get_timing(pack.handle, part_buf, C.byref(part_cnt)))
            if part_cnt.value == 4:
                cd_parts.append([part_buf[i] for i in range(4)])

    peaks = probes(env) if env.rank == 0 else None
    ms, wall, clocks = timed_steps(env, step_and_parts, steps, warmup, 4, 1 if light else 3)      # [sdr, cd, best+collective, total]
    ctr = (C.c_uint64 * 4)(); cc = C.c_int32(0)
    lib.check(L.qcqp_cd_get_counters(pack.handle, ctr, C.byref(cc)))
    st = np.frombuffer(d_stats.cpu().numpy().tobytes(), dtype=STATS_DT)
    info = pack.info
    sw1, sw2 = float(st["s1"].sum()) / n, float(st["s2"].sum()) / n
    parts = np.mean(np.array(cd_parts), axis=0) if cd_parts else None
    t_step, t_cd = env.max_over_ranks([ms[3] * 1e-3, ms[1] * 1e-3])
    t_dom = env.max_over_ranks([(parts[2] if parts is not None else ms[1]) * 1e-3])[0]
    tot_sweeps, tot_restarts = env.sum_over_ranks([sw1 + sw2, R])
    value = tot_sweeps / t_step

    # ---- e2e: the public host-buffer API, pinned host buffers, every copy inside the timed region, the same single collective ----
    from qcqp_b200.dist import local_best, global_best
    Ap = t.from_numpy(A).pin_memory().numpy() if R else A
    out_e = (t.empty((R, n), dtype=t.float64).pin_memory().numpy(), t.empty(R, dtype=t.float64).pin_memory().numpy(),
             t.empty(R, dtype=t.float64).pin_memory().numpy())
    e2e_t, e2e_sweeps = [], 0.0
    if sdr:
        h2d = Ap.nbytes + seeds.nbytes
        d2h = R * n * 8 + 2 * R * 8 + R * C.sizeof(lib.CdStats) + 4
    else:
        h2d = Ap.nbytes + R * C.sizeof(lib.RngState)
        d2h = R * n * 8 + 2 * R * 8 + R * C.sizeof(lib.CdStats) + R * C.sizeof(lib.RngState)
    # Two passes over the same loop.  single: every call uploads its own inputs first (one isolated request).  pipelined (the steady
    # state of a caller that runs batch after batch, SDR configurations): before the call on this step's batch the upload of the
    # NEXT step's standard normals is started with qcqp_sdr_prefetch, so it runs beside this step's kernels; every step still
    # issues one 8 MB upload from pinned memory and reads its results back inside its timed region.
    e2e_single = None
    rng_src = None if sdr else engine.rng_states(seeds=[int(s) for s in seeds])   # np.random.seed(seed_r) streams, built once (host work, untimed)
    for mode in (("single", "pipelined") if (sdr and R) else ("single",)):
        e2e_t = []
        if mode == "pipelined":
            pack.sdr_prefetch(Ap)                      # the first step's inputs (untimed warm-up step below consumes them)
        for it in range(1 + (1 if light else min(steps, 3))):
            env.barrier()
            t0 = time.perf_counter()
            if sdr:
                if mode == "pipelined":
                    pack.sdr_prefetch(Ap)              # next step's inputs: asynchronous, overlaps this step's kernels
                res = pack.sdr_cd_pipeline(seeds, mu=mu if (it == 0 and mode == "single") else None, F=F if (it == 0 and mode == "single") else None,
                                           Z=Ap, out=out_e)
                fh, vh, sth, Xh = res["f0"], res["maxviol"], res["stats"], res["X"]
            else:
                rng_e = type(rng_src)()                       # this step's MT19937 states: a fresh copy of the seeded streams
                C.memmove(rng_e, rng_src, C.sizeof(rng_src))
                Xh, fh, vh, sth = pack.cd_improve(Ap, rng_e)
            b, f, i = local_best(fh, vh)
            gb = global_best(b, f, lo + i if i >= 0 else -1, device=env.dev, x=Xh[i] if i >= 0 else np.zeros(n))
            env.barrier()
            dt = time.perf_counter() - t0
            if it > 0:
                e2e_t.append(dt)
                e2e_sweeps = sum(s.steps_p1 + s.steps_p2 for s in sth) / float(n)
        if mode == "single":
            e2e_single = env.max_over_ranks([float(np.mean(e2e_t))])[0]
    e2e_time = env.max_over_ranks([float(np.mean(e2e_t))])[0]
    e2e_sweeps = env.sum_over_ranks([e2e_sweeps])[0]
    agree = bool(np.allclose(fh, d_f.cpu().numpy(), rtol=1e-9, atol=0)) if R else True
    agree = agree and (env.world > 1 or gb[2] == int(d_best.item()))
    if env.world > 1:
        agree = agree and bool(np.array_equal(gb[3], pick.xwin.cpu().numpy()))       # host pick == device pick of the winner's point

    line = None
    if env.rank == 0:
        hbm, peak_src = measured_peaks()
        hbm_peak = float(hbm.get("hbm_gbs", 6650.0))
        alg_bytes = sw1 * info.bytes_per_sweep_phase1 + sw2 * info.bytes_per_sweep_phase2
        lpc2 = parts is not None and cc.value == 4 and ctr[3] > 0
        if lpc2:
            # the dominant kernel: cd_lpc2_kernel (phase 2).  Its bound is the L2 -> SM traffic it requests (counted by the kernel itself:
            # one row of P0 per accepted move, one 32 x 32 block + constants per pass), against the L2 read bandwidth probed on this box
            achieved = float(ctr[3]) / t_dom / 1e9
            roof = {"bound": "l2", "achieved": achieved, "peak": peaks["l2_read_GBps"], "unit": "GB/s", "frac": achieved / peaks["l2_read_GBps"],
                    "traffic": committed_traffic("qcqp::cd_lpc2_kernel"), "kernel": "qcqp::cd_lpc2_kernel (phase 2 of the separable dense path)",
                    "peak_source": "qcqp_probe_l2_bandwidth, this run",
                    "bytes_per_launch": int(ctr[3]), "rows_of_P0_per_launch": int(ctr[0]), "blocks_per_launch": int(ctr[1]),
                    "note": "P0 (8 MB) is L2-resident and a row is read only when its coordinate moves, so HBM does not bound this kernel (traffic = DRAM "
                            "bytes of the committed ncu capture).  In the tail of a launch the slowest restart's chain of dependent decisions bounds it.",
                    "model_only": {"what": "SURVEY 8d streaming model: every form read once per restart-sweep (%.0f B per phase-2 sweep), no credit for "
                                           "reuse -- NOT what the kernel does" % info.bytes_per_sweep_phase2,
                                   "GBps": sw2 * info.bytes_per_sweep_phase2 / t_dom / 1e9, "vs_hbm_peak": sw2 * info.bytes_per_sweep_phase2 / t_dom / 1e9 / hbm_peak,
                                   "hbm_peak_GBps": hbm_peak, "peak_source": peak_src}}
        else:
            kname = "qcqp::cd_blk_kernel" if not info.separable else "qcqp::cd_lpc_kernel"
            achieved = alg_bytes / t_cd / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": committed_traffic(kname), "kernel": kname, "peak_source": peak_src,
                    "note": "streaming model of SURVEY 8d (algorithmic bytes: phase-1 sweep %.0f B, phase-2 sweep %.0f B per restart) over the whole "
                            "qcqp_cd_improve launch sequence; the forms are L2-resident, so the model is an upper bound on what HBM sees"
                            % (info.bytes_per_sweep_phase1, info.bytes_per_sweep_phase2)}
        cpu = cpu_baseline_leg(cfg, key, forms, Xstar) if cpu_leg else None
        conf = static_config(cfg)
        line = {
            "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": env.world, "steps": steps, "warmup": max(warmup, 1 if light else 3),
            "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": conf,
            "detail": {"restarts_total": int(tot_restarts), "restarts_this_rank": R,
                       "sdr_solution": "X* of the SDP relaxation, host solve by qcqp_b200/relax.py (unit-diagonal mixing method), untimed" if sdr else None,
                       "l2": "flushed between timed iterations (384 MiB memset outside the event brackets)",
                       "sweeps_per_step_rank0": {"phase1": sw1, "phase2": sw2, "phase2_max_per_restart": int(st["w2"].max()) if R else 0,
                                                 "phase2_mean_per_restart": float(st["w2"].mean()) if R else 0.0,
                                                 "phase1_steps_fast_forwarded": int(st["skip"].sum()), "restarts_reaching_phase2": int(st["ran2"].sum())},
                       "restarts_per_s": tot_restarts / t_step,
                       "kernel_ms_rank0": {"sdr_sample_eval": float(ms[0]), "cd_improve": float(ms[1]), "best_and_collective": float(ms[2]),
                                           "cd_parts": None if parts is None else {"phase1_kernel": float(parts[0]), "gemm_G_eq_X_P0": float(parts[1]),
                                                                                   "phase2_kernel": float(parts[2]), "batched_eval": float(parts[3])}},
                       "collective": "one all-gather of (f0, maxviol, x) per rank + the library's best kernel over the ranks, enqueued on the launching "
                                     "stream inside the timed bracket" if env.world > 1 else "none (one GPU)",
                       "on_box_peaks": peaks, "device_vs_host_api_agree": agree, "wall_s_timed_loop": wall},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_sweeps / e2e_time, "unit": cfg["unit"], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_time,
                    "how": ("steady state of consecutive batches: the upload of the next step's inputs (qcqp_sdr_prefetch, pinned memory) is issued "
                            "before this step's qcqp_sdr_cd_pipeline call and runs beside its kernels; results read back in the call"
                            if (sdr and R) else "one host-buffer call per step, copies inside the call"),
                    "single_call": {"value": e2e_sweeps / e2e_single, "ms_per_step": 1e3 * e2e_single,
                                    "what": "the same call with nothing overlapped: upload, kernels, read-back in sequence"}},
            # per step: SDR (GEMM, GEMM row-dot, finish) + CD launch sequence + best (+ best over the ranks)
            "gpu_launches": ((3 if sdr else 0) + (5 if parts is not None else 1) + 1 + (1 if env.world > 1 else 0)) * steps,
            "clocks": clocks,
        }
    pack.close()
    return line


def committed_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/cd_traffic.json), or None."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "cd_traffic.json")))
        for e in (tj if isinstance(tj, list) else [tj]):
            if e.get("kernel", "").startswith(kernel):
                return e.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def cpu_baseline_leg(cfg, key, forms, Xstar):
    """cpu_baseline of the own arm (rank 0, N = 1): a bounded window of the unmodified Python reference where it applies (kind
    'reference'), and the C port beside it."""
    from oracle import oracle as orc
    from oracle import ref_python as rp
    pthreads = orc.lib().orc_max_threads()
    pr = PORT_SAMPLE[key]
    wk, dt = port_sample(cfg, forms, Xstar, pr, pthreads)
    w1, d1 = port_sample(cfg, forms, Xstar, min(pr, 4), 1)
    port = {"value": wk / dt, "unit": cfg["unit"], "cores": pthreads, "kind": "port", "single_core_value": w1 / d1,
            "sample": "%d restarts, complete runs, oracle/qcqp_oracle.c cached-f mode, %d threads, %.1f s" % (pr, pthreads, dt)}
    if cfg["kind"] == "sdr_cd" and rp.ref_root() is not None:
        cores = os.cpu_count() or 1
        K = 8 if key == "c2" else 2
        v, wall, st = reference_windows(cfg, cores, K)
        return {"value": v, "unit": cfg["unit"], "cores": cores, "kind": "reference",
                "sample": "unmodified cvxgrp/qcqp (baseline/_ref): %d worker processes x %d coordinate steps of coord_descent_phase2 each (%.1f s); "
                          "extrapolated from the window to restart-sweeps/s" % (cores, K, wall), "c_port": port}
    return port


def run_admm_config(env, key, steps, warmup, cpu_leg, light=False):
    """c4: improve_admm for this rank's share of the 16 rho values from one start, through qcqp_admm_improve_device."""
    t, L, lib = env.torch, env.L, env.lib
    from qcqp_b200 import engine
    cfg = CONFIGS[key]
    forms, _info, _ = build_problem(cfg)
    pack = engine.Pack(forms)
    n = pack.n
    rhos_all = np.sqrt(32) * 2.0 ** (np.arange(-8, 8) / 2.0)
    lo, hi = shard(len(rhos_all), env.rank, env.world)
    rhos = np.ascontiguousarray(rhos_all[lo:hi]); K = len(rhos)
    np.random.seed(4)
    X0 = 2 * np.random.randn(1, n)
    pack.compute_eig()
    Zinv = np.ascontiguousarray(np.stack([pack.zinv(r) for r in rhos])) if K else np.zeros((0, n, n))
    d_rho, d_Z, d_X0 = t.from_numpy(rhos).to(env.dev), t.from_numpy(Zinv).to(env.dev), t.from_numpy(X0).to(env.dev)
    d_X = t.empty((max(K, 1), n), dtype=t.float64, device=env.dev); d_f = t.empty(max(K, 1), dtype=t.float64, device=env.dev); d_v = t.empty_like(d_f)
    d_st = t.zeros(max(K, 1) * C.sizeof(lib.AdmmStats), dtype=t.uint8, device=env.dev)
    d_best = t.zeros(1, dtype=t.int32, device=env.dev)
    prm = lib.AdmmParams(1000, 1e4, 1e-2, 1)
    pick = CrossRankBest(env, n)

    def step(ev):
        if ev: ev[0].record()
        if K:
            lib.check(L.qcqp_admm_improve_device(pack.handle, C.byref(prm), d_rho.data_ptr(), d_Z.data_ptr(), K, d_X0.data_ptr(), 1, d_X.data_ptr(),
                                                  d_f.data_ptr(), d_v.data_ptr(), d_st.data_ptr(), env.stream))
        if ev: ev[1].record()
        if K:
            lib.check(L.qcqp_best_device(d_f.data_ptr(), d_v.data_ptr(), K, 1e-4, d_best.data_ptr(), None, None, env.stream))
            pick(d_f, d_v, d_X, d_best)
        if ev: ev[2].record()

    ms, wall, clocks = timed_steps(env, step, steps, warmup, 3, 1 if light else 3)
    st = np.frombuffer(d_st.cpu().numpy().tobytes(), dtype=ADMM_DT)[:K]
    iters = float((st["p1"] + st["p2"]).sum())
    t_step = env.max_over_ranks([ms[2] * 1e-3])[0]
    tot_iters = env.sum_over_ranks([iters])[0]
    # e2e: the host API (Zinv and the start uploaded, results read back)
    e2e_t = []
    for it in range(1 + (1 if light else min(steps, 3))):
        env.barrier()
        t0 = time.perf_counter()
        if K:
            Xh, fh, vh, sth = pack.admm_improve(X0, rhos)
        env.barrier()
        if it > 0:
            e2e_t.append(time.perf_counter() - t0)
    e2e_time = env.max_over_ranks([float(np.mean(e2e_t))])[0]
    line = None
    if env.rank == 0:
        info = pack.info
        # streaming-equivalent bytes of one ADMM iteration: Q_i and Q_i^T of every constraint + vectors + Zinv
        m = pack.m
        b_iter = m * (2 * n * n * 8 + 5 * n * 8) + n * n * 8
        hbm, peak_src = measured_peaks()
        hbm_peak = float(hbm.get("hbm_gbs", 6650.0))
        achieved = tot_iters / max(env.world, 1) * b_iter / t_step / 1e9
        cpu = None
        if cpu_leg:
            from oracle import oracle as orc
            pthreads = orc.lib().orc_max_threads()
            wk, dt = port_sample(cfg, forms, None, 16, pthreads)
            cpu = {"value": wk / dt, "unit": cfg["unit"], "cores": pthreads, "kind": "port",
                   "sample": "the whole 16-rho sweep, oracle/qcqp_oracle.c improve_admm, %d threads, %.1f s" % (pthreads, dt)}
        line = {"metric": cfg["metric"], "value": tot_iters / t_step, "unit": cfg["unit"], "n_gpus": env.world, "steps": steps, "warmup": max(warmup, 3),
                "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": static_config(cfg),
                "detail": {"rhos_this_rank": K, "iterations_rank0": iters, "onecons_calls_rank0": int(st["calls"].sum()) if K else 0,
                           "feasible_runs_rank0": int((d_v.cpu().numpy()[:K] < 1e-2).sum()) if K else 0, "wall_s_timed_loop": wall},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": committed_traffic("qcqp::admm_res_kernel"),
                             "kernel": "qcqp::admm_res_kernel", "peak_source": peak_src,
                             "note": "streaming-equivalent model (%.0f B per ADMM iteration); Q_i is resident in shared memory, so this fraction only says "
                                     "how far the lock-step iteration is from streaming speed" % b_iter},
                "cpu_baseline": cpu,
                "e2e": {"value": tot_iters / e2e_time, "unit": cfg["unit"], "h2d_bytes_per_step": int(Zinv.nbytes + X0.nbytes + rhos.nbytes),
                        "d2h_bytes_per_step": int(K * n * 8 + 2 * K * 8 + K * C.sizeof(lib.AdmmStats)), "ms_per_step": 1e3 * e2e_time},
                "gpu_launches": (1 + 1 + (1 if env.world > 1 else 0)) * steps, "clocks": clocks}
    pack.close()
    return line


def run_own(args):
    env = Env()
    runner = {"sdr_cd": run_cd_config, "cd": run_cd_config, "admm": run_admm_config}
    cfg = CONFIGS[args.config]
    line = runner[cfg["kind"]](env, args.config, args.steps, args.warmup, cpu_leg=(env.world == 1 and not args.no_cpu))
    extras = {}
    if args.config == "c2" and not args.no_extras:
        # the other BASELINE.json configurations, a few steps each, so that the driver's default run carries them too
        for key in ("c3", "c4", "c5"):
            try:
                ex = runner[CONFIGS[key]["kind"]](env, key, 1 if key == "c5" else 3, 1, cpu_leg=False, light=True)
                if env.rank == 0:
                    extras[key] = {k: ex[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "roofline", "e2e", "gpu_launches", "steps")}
                    extras[key]["detail"] = {k: v for k, v in ex["detail"].items() if k in ("restarts_total", "sweeps_per_step_rank0", "restarts_per_s", "iterations_rank0", "feasible_runs_rank0", "kernel_ms_rank0")}
            except Exception as e:                                   # an extra must never cost the headline line
                if env.rank == 0:
                    extras[key] = {"error": "%s: %s" % (type(e).__name__, e)}
    if env.rank == 0:
        if extras:
            line["other_configs"] = extras
        print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="do not append the short c3/c4/c5 measurements to the default (c2) line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
