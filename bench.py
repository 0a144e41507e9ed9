#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200:

    coordinate-descent restart-sweeps/s on Boolean least squares n=1000 / m_rows=1500 (dense P0, 1000 constraints
    x_i^2 = 1), 1024 SDR samples + COORD_DESCENT per GPU  ("configs[1]").

One STEP = one pass of the hot path over one batch: SDR randomized rounding of 1024 draws (x = mu + z F, eval)
-> improve_coord_descent on the 1024 draws (each restart its own MT19937 stream, reference defaults
num_iters=1000, viol_tol=1e-2, tol=1e-4) -> best-pick.  Unit of work = one restart-sweep = n coordinate steps of one
restart (one outer iteration of qcqp.py:110 / :160); only EXECUTED coordinate steps are counted.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); restarts are sharded weak-scaling (1024 per GPU, disjoint seeds), no
collective on the data path, one 3-scalar all-reduce per step to pick the best point.
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

N_VAR, M_ROWS, SAMPLES = 1000, 1500, 1024
METRIC = "coord_descent_restart_sweeps_per_sec"
UNIT = "restart-sweeps/s"
WORKLOAD = "boolean_least_squares n=1000 m_rows=1500 (dense P0, 1000 x_i^2==1 constraints), 1024 SDR samples + COORD_DESCENT per GPU"


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
        time.sleep(0.15)
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


def build_problem():
    """The C2 instance and its SDP relaxation solution X* (host solve, outside every timed region -- the reference
    caches it too, qcqp.py:390-395)."""
    from qcqp_b200 import problems as pb, relax
    from qcqp_b200.forms import QCQPForm
    forms, _ = pb.boolean_least_squares(N_VAR, M_ROWS, seed=1)
    Xstar, _bound = relax.solve_sdr(QCQPForm.from_tuples(forms))
    return forms, Xstar


STATS_DT = np.dtype([("s1", "<i8"), ("s2", "<i8"), ("u1", "<i8"), ("u2", "<i8"), ("w1", "<i4"), ("w2", "<i4"),
                     ("status", "<i4"), ("ran2", "<i4"), ("skip", "<i8")])


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host cores (the oracle port; the reference is pure Python and its
# own code cannot be compiled)
# ------------------------------------------------------------------------------------------------------------
def cpu_sample(forms, Xstar, restarts, threads, fast=True, num_iters=1000, seed0=1000):
    from oracle import oracle as orc
    P = orc.Problem(forms)
    mu, _Sigma, F = orc.sdr_factor(Xstar)
    rs = np.random.RandomState(2)
    Z = rs.standard_normal((restarts, N_VAR))
    t0 = time.perf_counter()
    X0, _f, _v = P.sdr_sample_eval(mu, F, Z, nthreads=threads)
    rngs = (orc.RngState * restarts)()
    for r in range(restarts):
        rngs[r] = orc.RngState.from_seed(seed0 + r)
    X, f0, mv, st = P.improve_cd_batch(X0, rngs, fast=fast, nthreads=threads, num_iters=num_iters)
    dt = time.perf_counter() - t0
    steps = sum(s.steps_p1 + s.steps_p2 for s in st)
    return steps / float(N_VAR), dt, float(np.min(f0))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    forms, Xstar = build_problem()
    cores = orc.lib().orc_max_threads()
    restarts = max(32 * cores, 64)
    for _ in range(args.warmup):
        cpu_sample(forms, Xstar, cores, cores)
    sweeps, secs = 0.0, 0.0
    for _ in range(args.steps):
        s, dt, _ = cpu_sample(forms, Xstar, restarts, cores)
        sweeps += s; secs += dt
    value = sweeps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d restarts per step (bounded sample of the 1024)" % restarts},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d SDR draws + full improve_coord_descent each, oracle/qcqp_oracle.c fast mode (cached f_j, "
                                   "incidence lists), %d threads" % (restarts, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    from qcqp_b200 import _lib, engine
    from qcqp_b200.dist import local_best, global_best

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    forms, Xstar = build_problem()
    pack = engine.Pack(forms)
    mu, _Sigma, F = engine.sdr_factor(Xstar)             # host SVD once, as np.random.multivariate_normal would per draw
    R = SAMPLES
    seed0 = 1000 + rank * R                              # disjoint MT19937 streams per rank (weak scaling)
    rs = np.random.RandomState(2 + rank)
    Z = rs.standard_normal((R, N_VAR))
    rng_host = engine.rng_states(seeds=[seed0 + r for r in range(R)])
    rng_bytes = engine.rng_states_as_tensor_bytes(rng_host)

    # ---- device-resident inputs for `value` ----
    d_mu = torch.from_numpy(mu).to(dev); d_F = torch.from_numpy(F).to(dev); d_Z = torch.from_numpy(Z).to(dev)
    d_rng0 = torch.from_numpy(rng_bytes).to(dev); d_rng = torch.empty_like(d_rng0)
    d_X0 = torch.empty((R, N_VAR), dtype=torch.float64, device=dev); d_X = torch.empty_like(d_X0)
    d_f = torch.empty(R, dtype=torch.float64, device=dev); d_v = torch.empty_like(d_f)
    d_fs = torch.empty_like(d_f); d_vs = torch.empty_like(d_f)
    d_stats = torch.zeros(R * C.sizeof(_lib.CdStats), dtype=torch.uint8, device=dev)
    d_best = torch.zeros(1, dtype=torch.int32, device=dev); d_bb = torch.zeros(1, dtype=torch.int64, device=dev)
    d_bf = torch.zeros(1, dtype=torch.float64, device=dev)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2
    prm = _lib.CdParams(1000, 1e-2, 1e-4, 1, 0, 0)
    stream = torch.cuda.current_stream().cuda_stream

    def device_step(ev=None):
        """SDR sample+eval -> CD improve -> best, all enqueued on torch's current stream."""
        d_rng.copy_(d_rng0)
        if ev: ev[0].record()
        _lib.check(L.qcqp_sdr_sample_eval_device(pack.handle, d_mu.data_ptr(), d_F.data_ptr(), d_Z.data_ptr(), 0, R, d_X0.data_ptr(),
                                                  d_fs.data_ptr(), d_vs.data_ptr(), stream))
        if ev: ev[1].record()
        _lib.check(L.qcqp_cd_improve_device(pack.handle, C.byref(prm), d_X0.data_ptr(), R, d_rng.data_ptr(), d_X.data_ptr(), d_f.data_ptr(),
                                             d_v.data_ptr(), d_stats.data_ptr(), stream))
        if ev: ev[2].record()
        _lib.check(L.qcqp_best_device(d_f.data_ptr(), d_v.data_ptr(), R, 1e-4, d_best.data_ptr(), d_bb.data_ptr(), d_bf.data_ptr(), stream))
        if ev: ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.zero_(); device_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, sdr_ms, cd_ms, cd_parts = [], [], [], []
    part_buf = (C.c_double * 4)(); part_cnt = C.c_int32(0)
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (outside the event brackets)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        device_step(ev)
        if world > 1:                                   # the only collective: best (bucket, f0, index) across ranks
            b = int(d_bb.item()); f = float(d_bf.item()); i = int(d_best.item())
            global_best(b, f, seed0 - 1000 + i, device=dev)
        torch.cuda.synchronize()
        step_ms.append(ev[0].elapsed_time(ev[3])); sdr_ms.append(ev[0].elapsed_time(ev[1])); cd_ms.append(ev[1].elapsed_time(ev[2]))
        _lib.check(L.qcqp_cd_get_timing(pack.handle, part_buf, C.byref(part_cnt)))
        if part_cnt.value == 4:
            cd_parts.append([part_buf[i] for i in range(4)])
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None

    st = np.frombuffer(d_stats.cpu().numpy().tobytes(), dtype=STATS_DT)
    assert (st["status"] == 0).all()
    info = pack.info
    sweeps_p1 = float(st["s1"].sum()) / N_VAR; sweeps_p2 = float(st["s2"].sum()) / N_VAR
    sweeps = sweeps_p1 + sweeps_p2
    alg_bytes = sweeps_p1 * info.bytes_per_sweep_phase1 + sweeps_p2 * info.bytes_per_sweep_phase2
    t_step = float(np.mean(step_ms)) * 1e-3
    t_cd = float(np.mean(cd_ms)) * 1e-3
    parts = np.mean(np.array(cd_parts), axis=0) if cd_parts else None       # [phase-1 kernel, G GEMM, phase-2 kernel, eval] ms
    # the dominant kernel: phase 2 (cd_lpc_kernel stage 2) when the launch sequence is split, else the whole CD launch
    t_dom = float(parts[2]) * 1e-3 if parts is not None else t_cd
    dom_bytes = sweeps_p2 * info.bytes_per_sweep_phase2 if parts is not None else alg_bytes
    if world > 1:
        tt = torch.tensor([t_step, t_cd, t_dom], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)            # device-timed, max over ranks
        t_step, t_cd, t_dom = float(tt[0].item()), float(tt[1].item()), float(tt[2].item())
        ts = torch.tensor([sweeps], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        total_sweeps = float(ts.item())
    else:
        total_sweeps = sweeps
    value = total_sweeps / t_step

    # ---- e2e: the public host-buffer API -- ONE C-ABI call per step (qcqp_sdr_cd_pipeline: draws -> coordinate descent -> best),
    #      host buffers pinned, every copy inside the timed region.  Per step the host supplies the S x n standard normals and one
    #      np.random.seed value per restart and reads back the improved points with their (f0, maxviol), statistics and the best
    #      index; mu / F are cached on the pack by the first (untimed) call, as the reference caches them on self (qcqp.py:394-395).
    Zp = torch.from_numpy(Z).pin_memory().numpy()
    seeds_e = np.array([seed0 + r for r in range(R)], dtype=np.uint32)
    out_e = (torch.empty((R, N_VAR), dtype=torch.float64).pin_memory().numpy(), torch.empty(R, dtype=torch.float64).pin_memory().numpy(),
             torch.empty(R, dtype=torch.float64).pin_memory().numpy())
    e2e_t, e2e_sweeps = [], 0.0
    h2d = Zp.nbytes + seeds_e.nbytes
    d2h = R * N_VAR * 8 + 2 * R * 8 + R * C.sizeof(_lib.CdStats) + 4
    for it in range(1 + min(args.steps, 3)):
        barrier()
        t0 = time.perf_counter()
        res = pack.sdr_cd_pipeline(seeds_e, mu=mu if it == 0 else None, F=F if it == 0 else None, Z=Zp, out=out_e)
        fh, vh, sth, bi = res["f0"], res["maxviol"], res["stats"], res["best"]
        if world > 1:
            b, f, i = local_best(fh, vh)
            global_best(b, f, seed0 - 1000 + i, device=dev)
        barrier()
        dt = time.perf_counter() - t0
        if it > 0:
            e2e_t.append(dt)
            e2e_sweeps = sum(s.steps_p1 + s.steps_p2 for s in sth) / float(N_VAR)
    e2e_time = float(np.mean(e2e_t))
    if world > 1:
        tt = torch.tensor([e2e_time], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_time = float(tt.item())
        ts = torch.tensor([e2e_sweeps], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        e2e_sweeps = float(ts.item())
    # the device path and the host path must agree (same inputs, same streams)
    agree = bool(np.allclose(fh, d_f.cpu().numpy(), rtol=1e-9, atol=0) and bi == int(d_best.item()))

    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = dom_bytes / t_dom / 1e9
        traffic = None
        kname = "qcqp::cd_lpc_kernel (stage 2: phase 2)" if parts is not None else ("qcqp::cd_lpc_kernel" if info.separable else "qcqp::cd_kernel")
        tpath = os.path.join(ROOT, "profiles", "cd_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch") if tj.get("kernel") == kname else None
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import oracle as orc
            cores = orc.lib().orc_max_threads()
            n_cpu = max(32 * cores, 64)
            s_fast, dt_fast, _ = cpu_sample(forms, Xstar, n_cpu, cores, fast=True)
            s_1, dt_1, _ = cpu_sample(forms, Xstar, 4, 1, fast=True)
            s_ff, dt_ff, _ = cpu_sample(forms, Xstar, cores, cores, fast=False, num_iters=1)
            cpu = {"value": s_fast / dt_fast, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d of the 1024 SDR draws, full improve_coord_descent each (oracle/qcqp_oracle.c, cached-f mode), %d threads, %.1f s"
                             % (n_cpu, cores, dt_fast),
                   "single_core_value": s_1 / dt_1,
                   "faithful_value": s_ff / dt_ff,
                   "faithful_sample": "%d draws, num_iters=1, every get_onevar_func recomputing t0 as utilities.py:99-105 does, %d threads, %.1f s"
                                      % (cores, cores, dt_ff)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "restarts_per_gpu": R, "n": N_VAR, "m": N_VAR, "num_iters": 1000, "viol_tol": 1e-2,
                       "tol": 1e-4, "rng": "MT19937 stream per restart (np.random.seed(1000 + r))",
                       "sdr_solution": "X* of the SDP relaxation, host solve by qcqp_b200/relax.py (unit-diagonal mixing method), untimed",
                       "l2": "flushed between timed iterations (384 MiB memset outside the event brackets)",
                       "sweeps_per_step": {"phase1": sweeps_p1, "phase2": sweeps_p2,
                                           "phase2_max_per_restart": int(st["w2"].max()), "phase2_mean_per_restart": float(st["w2"].mean()),
                                           "phase1_steps_fast_forwarded": int(st["skip"].sum())},
                       "kernel_ms": {"sdr_sample_eval": float(np.mean(sdr_ms)), "cd_improve": float(np.mean(cd_ms)),
                                     "cd_parts": None if parts is None else {"phase1_kernel": float(parts[0]), "gemm_G_eq_X_P0": float(parts[1]),
                                                                             "phase2_kernel": float(parts[2]), "batched_eval": float(parts[3])}},
                       "cd_sequence_model_GBps": alg_bytes / t_cd / 1e9,
                       "device_vs_host_api_agree": agree, "wall_s_timed_loop": wall},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": kname, "peak_source": peak_src,
                         "model": "algorithmic streaming bytes (SURVEY 8d): phase-2 sweep %.0f B, phase-1 sweep %.0f B per restart, i.e. "
                                  "every form read once per restart-sweep.  The kernel reads a row of P0 only when its coordinate moves "
                                  "(cached g = P0 x), so frac > 1 means fewer bytes than the model, not skipped sweeps; `traffic` is the "
                                  "measured DRAM bytes per launch (P0 is L2-resident)" % (info.bytes_per_sweep_phase2, info.bytes_per_sweep_phase1)},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_sweeps / e2e_time, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_time},
            # per step: SDR (GEMM, GEMM row-dot, finish) + CD (phase-1 kernel, GEMM, phase-2 kernel, GEMM row-dot, finish) + best
            "gpu_launches": (9 if parts is not None else 5) * args.steps,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
