#!/usr/bin/env python
"""bench.py -- headline benchmark of the mceik hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   (the reference's CPU path, rank 0 only)
    python bench.py --config 2 | --config 5                   (the other BASELINE configs that fit one GPU)

A *step* is one pass of the eikonal hot path over one batch of synthetic input.  Default = BASELINE config 3:
64 stations x (P, S) = 128 fields on the 256^3 checkerboard model, solved to convergence (boundary conditions +
all sweeps + convergence tests) and delivered as fp32 tables; for N > 1 the fields are shared out over the ranks
by the library (mceik_fsm_solve_sharded_dev: balanced assignment, tables written into the replicated buffer and put
into the other ranks' copies over NVLink as the fields converge) -- strong scaling.
`value` = node-updates of all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e`   = the same through the host-pointer C-ABI call (mceik_fsm_solve_batched_host): pinned host slowness in,
          fp64 fields out, copies inside the timed region.
`parity_checked` = fields / events of THIS run compared bit for bit with the CPU oracle at the configured size.
The second half of the metric (events located/s, BASELINE config 4 shard) is the `events` object of the same JSON
line with its own roofline / cpu_baseline / e2e.  CPU legs (the only code that executes oracle/) run in a
subprocess after the GPU timing: `python bench.py --cpu-leg ...`.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import cases  # noqa: E402

H = 1000.0
BYTES_PER_UPDATE = 24  # fp64 read u, write u, read slow
METRIC = "eikonal_node_updates_per_s"
UNIT = "Gnode-updates/s"
CPU_ITERS = 2  # iterations per field of the bounded CPU samples (cost per node-update does not depend on the iteration)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 5], help="BASELINE config (3 = the headline)")
    ap.add_argument("--grid", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--fields", type=int, default=0, help=argparse.SUPPRESS)  # total over all ranks
    ap.add_argument("--gs-events", type=int, default=12500, help="events per GPU of the grid-search step (config 4 shard)")
    ap.add_argument("--gs-stations", type=int, default=128, help=argparse.SUPPRESS)
    ap.add_argument("--skip-gs", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--skip-cpu", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-leg", default="", help=argparse.SUPPRESS)   # fsm | gs : run one CPU leg and print its JSON
    ap.add_argument("--cpu-out", default="", help=argparse.SUPPRESS)   # directory for the parity arrays of a CPU leg
    a = ap.parse_args()
    if a.grid == 0:
        a.grid = {2: 128, 3: 256, 5: 96}[a.config]
    if a.fields == 0:
        a.fields = {2: 1, 3: 128, 5: 64 * 32}[a.config]
    return a


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def c3_inputs(a):
    """(slowness [2, N], field_model, xs, ys, zs): field f < fields/2 is station f with the P model, else station
    f - fields/2 with the S model."""
    n, half = a.grid, max(a.fields // 2, 1)
    slow = np.stack([cases.checkerboard_slowness(n, n, n, cell=32, vs=False), cases.checkerboard_slowness(n, n, n, cell=32, vs=True)])
    xs, ys, zs = cases.interior_sources(half, n, n, n, H, seed=3)
    ids = np.arange(a.fields)
    st = ids % half
    return slow, (ids >= half).astype(np.int32), xs[st], ys[st], zs[st]


def workload_config(a, n_gpus):
    if a.config == 2:
        return {"workload": f"BASELINE config 2: fsm3d single station, 1-D layered velocity (8 layers, 6500..3000 m/s), {a.grid}^3 grid, "
                            f"P-wave, tol 1e-6, maxit 20, solved to convergence",
                "fields_total": 1, "grid": [a.grid] * 3, "l2": "L2 flushed between timed solves (write of a 256 MB buffer)"}
    if a.config == 5:
        return {"workload": f"BASELINE config 5 (bounded slice): MCMC forward loop, {a.fields // 32} perturbed velocity models x 32 stations "
                            f"= {a.fields} P fields on {a.grid}^3, solved to convergence + fp32 tables + catalogue misfit "
                            f"(1000 events x 32 picks) per proposal; the full config has 1024 models",
                "fields_total": a.fields, "grid": [a.grid] * 3,
                "l2": f"inputs larger than L2: {a.fields * a.grid ** 3 * 8 / 1e9:.1f} GB of fp64 fields"}
    per = a.fields // n_gpus
    return {"workload": f"BASELINE config 3: fsm3d batched, {a.fields // 2} stations x (P,S) = {a.fields} fields on "
                        f"{a.grid}^3 checkerboard velocity (+-10%, 32-node cells; vs = vp/sqrt3), tol 1e-6, maxit 20, "
                        f"solved to convergence + fp32 tables"
                        + (" + replication of the tables on every GPU (one-sided puts over NVLink as the fields converge)" if n_gpus > 1 else ""),
            "fields_total": a.fields, "fields_per_gpu": per, "grid": [a.grid] * 3,
            "sharding": (f"sources shared out x{n_gpus} by mceik_fsm_assign_fields: fields of a slowness model dealt over the ranks "
                         f"holding it, longest first by the iteration counts of the previous solve") if n_gpus > 1 else "one GPU",
            "l2": f"inputs larger than L2: {per * a.grid ** 3 * 8 / 1e9:.1f} GB of fp64 fields per GPU"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, units_per_launch):
    """DRAM bytes per launch of `kernel`: bytes per unit of work (node-update / event) from the committed
    ncu --set full capture (profiles/roofline.json) x the units one launch of THIS run processes; or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline.json")) as f:
            return json.load(f)[kernel]["dram_bytes_per_unit"] * units_per_launch
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the only code that executes oracle/.  Run as `python bench.py --cpu-leg fsm|gs` in a subprocess of
# our arm (so the GPU process maps no oracle library) or directly by --impl reference.
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _fsm_field_worker(args):
    """One process = one core = one field, the first `maxit` iterations (16 sweeps for 2): serial port of fsm3d.f90."""
    n, model, f, maxit, barrier_path, idx = args
    import oracle_lib as O
    O.set_threads(1)
    a = argparse.Namespace(grid=n, fields=128 if model == 3 else 1, config=model)
    if model == 3:
        slow, fm, xs, ys, zs = c3_inputs(a)
        sl = slow[fm[f]]
    else:  # configs 2 and 5: the layered base model (the cost of a node-update does not depend on the perturbation)
        sl, xs, ys, zs = c2_inputs(a)
        f = 0
    O.lib()
    open(f"{barrier_path}.ready.{idx}", "w").close()
    while not os.path.exists(barrier_path):  # all workers leave the gate together
        time.sleep(0.002)
    u, ierr, it, dt = O.eikonal_serial_timed(n, n, n, H, sl, 0.0, xs[f], ys[f], zs[f], tol=1e-6, maxit=maxit)
    return n ** 3 * 8 * it, dt


def cpu_fsm_field_per_core(a, nproc, maxit=CPU_ITERS):
    """`nproc` fields at once, one per core (the strongest CPU arrangement: no intra-solve parallelism is needed
    when there are as many fields as cores).  Returns (Gnode-updates/s, seconds)."""
    import multiprocessing as mp
    gate = os.path.join(tempfile.gettempdir(), f"mceik_gate_{os.getpid()}_{time.time_ns()}")
    half = max(a.fields // 2, 1)
    jobs = [(a.grid, a.config, (i % 2) * half + (i // 2) % half if a.config == 3 else 0, maxit, gate, i) for i in range(nproc)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(nproc) as pool:
        res = pool.map_async(_fsm_field_worker, jobs, chunksize=1)
        t_end = time.time() + 600
        while not all(os.path.exists(f"{gate}.ready.{i}") for i in range(nproc)) and time.time() < t_end:
            time.sleep(0.01)  # imports + input construction happen before the gate opens
        open(gate, "w").close()
        out = res.get()
    for pth in [gate] + [f"{gate}.ready.{i}" for i in range(nproc)]:
        if os.path.exists(pth):
            os.remove(pth)
    upd = sum(u for u, _ in out)
    dt = max(t for _, t in out)
    return upd / dt / 1e9, dt


def cpu_leg_fsm(a):
    """CPU baseline of the eikonal half, the three arrangements of BASELINE.md section 3 / SURVEY 8d on this box's cores:
    (1) one core, (2) the reference's own parallelism -- OpenMP over the nodes of a hyperplane, fsm3d.f90:437 -- on all
    threads, (3) one field per core on all cores.  (2) runs two fields (one P, one S) to convergence and saves them
    for the parity check of the GPU arm."""
    import oracle_lib as O
    n, cores = a.grid, host_threads()
    if a.config == 3:
        slow, fm, xs, ys, zs = c3_inputs(a)
        picks = [0, a.fields // 2] if a.fields >= 2 else [0]
    elif a.config == 5:
        slow, fm, xs, ys, zs = c5_inputs(a)
        picks = [0, min(a.fields, 256) - 1]  # both inside the slice of fields the GPU arm's e2e step returns
    else:
        s1, xs, ys, zs = c2_inputs(a)
        slow, fm, picks = s1[None], np.zeros(1, np.int32), [0]
    out = {"cores": cores, "arrangements": {}, "parity_fields": []}
    O.set_threads(cores)
    upd = secs = 0.0
    for f in picks:
        u, ierr, it, dt = O.eikonal_serial_timed(n, n, n, H, slow[fm[f]], 0.0, xs[f], ys[f], zs[f], tol=1e-6, maxit=20)
        upd += n ** 3 * 8 * it
        secs += dt
        if a.cpu_out:
            path = os.path.join(a.cpu_out, f"fsm_field_{f}.npy")
            np.save(path, u)
            out["parity_fields"].append({"field": int(f), "iterations": int(it), "path": path})
    out["arrangements"]["omp_intra_level"] = {
        "value": upd / secs / 1e9, "cores": cores, "seconds": secs,
        "sample": f"{len(picks)} field(s) of the workload to convergence, OpenMP over hyperplane nodes on {cores} threads"}
    v1, t1 = cpu_fsm_field_per_core(a, 1)
    out["arrangements"]["one_core"] = {"value": v1, "cores": 1, "seconds": t1,
                                       "sample": f"1 field, first {CPU_ITERS} iterations ({8 * CPU_ITERS} sweeps), 1 thread"}
    vN, tN = cpu_fsm_field_per_core(a, cores)
    out["arrangements"]["field_per_core"] = {"value": vN, "cores": cores, "seconds": tN,
                                             "sample": f"{cores} fields at once, one per core, first {CPU_ITERS} iterations each"}
    return out


def gs_inputs(a, rank):
    """Config-4 shard: station positions / velocities of the analytic tables and the event list of this rank."""
    n, ns = a.grid if a.config == 3 else 256, a.gs_stations
    rng = np.random.default_rng(4)
    sx, sy = rng.uniform(0, (n - 1) * H, ns), rng.uniform(0, (n - 1) * H, ns)
    sz = np.full(ns, (n - 1) * H)
    X, Y, Z = np.repeat(sx, 2), np.repeat(sy, 2), np.repeat(sz, 2)
    V = np.tile(np.array([5000.0, 5000.0 / np.sqrt(3.0)]), ns)
    return n, X, Y, Z, V


def cpu_leg_gs(a):
    """CPU baseline of the grid-search half: the reference's OWN locate.c object (oracle/_ref,
    locate_l2_gridSearch__double64 + locate_minLocDouble64, locate.c:923-1047, 811-830) on the full 256 tables of
    the workload (promoted from fp32 as the catalogue path does, locate.f90:414), one event per host thread.
    The events and their picks come from the GPU arm (a .npz in --cpu-out); the located nodes go back for the
    parity check."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    import oracle_lib as O
    with np.load(os.path.join(a.cpu_out, "gs_events.npz")) as dz:
        d = {k: dz[k] for k in dz.files}  # read now: the worker threads below must not share the lazy zip reader
    n, X, Y, Z, V = gs_inputs(a, 0)
    N, ntab = n ** 3, len(X)
    cores = host_threads()
    ne = int(d["tobs"].shape[0])
    ref = O.ref()
    kind = "reference" if ref is not None else "port"
    ld = N  # 256^3 * 8 B is a multiple of 64
    t_build = time.perf_counter()
    tables = O.aligned(ntab * ld, np.float64)

    def build(t):  # homog.c:594-621, rounded to fp32 like the resident tables, promoted back
        tt = O.homogeneous_traveltimes(n, n, n, 0.0, 0.0, 0.0, H, H, H, X[t], Y[t], Z[t], V[t])
        tables[t * ld:(t + 1) * ld] = tt.astype(np.float32)
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(build, range(ntab)))
    t_build = time.perf_counter() - t_build
    pd = lambda x: x.ctypes.data_as(O.c_dbl_p)
    L = O.lib()

    def locate(fn, mfn, e, nt):
        """One event against tables [0, nt): (node, t0, objective) with the flavour-1 arithmetic of locate.c."""
        fn.restype = C.c_int
        t0, obj = O.aligned(N, np.float64), O.aligned(N, np.float64)
        mask = np.ascontiguousarray(1 - d["use"][e][:nt], dtype=np.int32)  # locate.c: mask == 0 means the pick is used
        tobs, var = np.ascontiguousarray(d["tobs"][e][:nt]), np.ascontiguousarray(d["var"][e][:nt])
        rc = fn(C.c_int(ld), C.c_int(N), C.c_int(nt), C.c_int(1), C.c_double(0.0), mask.ctypes.data_as(O.c_int_p), pd(tobs), None,
                pd(var), pd(tables), pd(t0), pd(obj))
        assert rc == 0
        i = int(mfn(C.c_int(N), pd(obj)))
        return i, float(t0[i]), float(obj[i]), int((1 - mask).sum())

    def timed(fn, mfn, nt):
        t = time.perf_counter()
        with ThreadPoolExecutor(max_workers=min(cores, ne)) as ex:
            res = list(ex.map(lambda e: locate(fn, mfn, e, nt), range(ne)))
        return res, time.perf_counter() - t

    # the whole workload (all tables) through the port: the reference's own object indexes the tables with a 32-bit
    # int (ibeg = ldgrd*iobs, locate.c:1023, 1039), which overflows beyond 128 tables of 256^3 nodes
    res, dt = timed(L.oracle_l2_gridsearch_f64, L.oracle_minloc_f64, ntab)
    out = {"value": ne / dt, "unit": "events/s", "cores": min(cores, ne), "kind": "port", "seconds": dt,
           "iopt": [r[0] for r in res], "t0": [r[1] for r in res], "obj": [r[2] for r in res]}
    note = ""
    if ref is not None:  # the reference object on as many tables as it can address, as a cross-check of the port's speed
        nt = min(ntab, (2 ** 31 - 1) // ld)
        res_r, dt_r = timed(ref.locate_l2_gridSearch__double64, ref.locate_minLocDouble64, nt)
        same = nt < ntab or [r[:3] for r in res_r] == [r[:3] for r in res]
        res_p, dt_p = (res, dt) if nt == ntab else timed(L.oracle_l2_gridsearch_f64, L.oracle_minloc_f64, nt)
        agree = [r[:3] for r in res_r] == [r[:3] for r in res_p]
        note = (f"; the reference's own locate.c object (oracle/_ref) on "
                + (f"the first {nt} tables (its 32-bit table index overflows beyond {nt} tables of {n}^3)" if nt < ntab else "the same tables")
                + f": {ne / dt_r:.3f} events/s vs the port's {ne / dt_p:.3f}, results " + ("bit-equal" if agree and same else "DIFFERENT"))
        out["reference_object"] = {"tables": nt, "value": ne / dt_r, "port_same_tables": ne / dt_p, "bit_equal": bool(agree and same)}
        if nt == ntab:
            out["kind"], out["value"] = "reference", ne / dt_r
    out["sample"] = (f"{ne} events of the workload x all {ntab} tables on {n}^3 (fp64, promoted from the fp32 tables as locate.f90:414 does) "
                     f"by oracle/locate_oracle.c (locate.c:923-1047 restated with 64-bit indexing), one event per host thread "
                     f"({min(cores, ne)} threads); tables built on the host in {t_build:.1f} s (not timed)" + note)
    return out


def run_cpu_leg_subprocess(a, leg, outdir):
    env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}  # torchrun exports OMP_NUM_THREADS=1
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", leg, "--cpu-out", outdir, "--config", str(a.config),
           "--grid", str(a.grid), "--fields", str(a.fields), "--gs-stations", str(a.gs_stations)]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1800)
    if out.returncode != 0:
        raise RuntimeError(f"cpu leg {leg} failed: {out.stderr[-2000:]}")
    return json.loads(out.stdout.strip().splitlines()[-1])


def best_arrangement(legs):
    k = max(legs["arrangements"], key=lambda k: legs["arrangements"][k]["value"])
    return k, legs["arrangements"][k]


def cpu_baseline_from(legs):
    k, b = best_arrangement(legs)
    others = "; ".join(f"{n}: {v['value']:.3f} G/s on {v['cores']} core(s) ({v['sample']})" for n, v in legs["arrangements"].items())
    return {"value": b["value"], "unit": UNIT, "cores": b["cores"], "kind": "port",
            "sample": f"strongest of three arrangements = {k}.  " + others +
                      ".  oracle/fsm3d_oracle.c = C port of fsm3d.f90's serial path (gcc -O2 -fopenmp; no Fortran compiler exists "
                      "to build the reference itself, profiles/toolchain_probe_r2.txt)",
            "arrangements": {n: {kk: v[kk] for kk in ("value", "cores", "seconds")} for n, v in legs["arrangements"].items()}}


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path (oracle port; no Fortran compiler exists to
    build fsm3d.f90 itself) with the reference's own parallelism -- OpenMP over the nodes of a hyperplane,
    fsm3d.f90:437 -- on all host threads, the strongest of the three CPU arrangements on the GPU boxes
    (cpu_baseline.arrangements of our arm lists all three); each step = one field of the workload to convergence."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    n = a.grid
    if a.config == 3:
        slow, fm, xs, ys, zs = c3_inputs(a)
        sl = slow[fm[0]]
    else:
        sl, xs, ys, zs = c2_inputs(a)
    cores = O.set_threads(host_threads())  # explicit: torchrun exports OMP_NUM_THREADS=1 to its children
    vals, it = [], 0
    for i in range(a.warmup + a.steps):
        u, ierr, it, dt = O.eikonal_serial_timed(n, n, n, H, sl, 0.0, xs[0], ys[0], zs[0], tol=1e-6, maxit=20)
        if i >= a.warmup:
            vals.append((n ** 3 * 8 * it / dt / 1e9, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3
    desc = (f"1 of the {a.fields} fields per step ({n}^3, {it} iterations, {n ** 3 * 8 * it / 1e9:.2f} G node-updates) by "
            f"oracle/fsm3d_oracle.c (C port of fsm3d.f90's serial path, gcc -O2 -fopenmp, OpenMP over hyperplane nodes, {cores} threads)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if a.config == 3 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# config 2 / config 5 inputs
# ------------------------------------------------------------------------------------------------
def c2_inputs(a):
    n = a.grid
    xs, ys, zs = cases.interior_sources(1, n, n, n, H, seed=1)
    return cases.layered_slowness(n, n, n, layer=max(n // 8, 1)), xs, ys, zs


def c5_inputs(a):
    """96^3: base 1-D model x (1 + 0.05 * smooth gaussian field), one P model per proposal, 32 stations."""
    n, nmod, nst = a.grid, a.fields // 32, 32
    rng = np.random.default_rng(5)
    base = cases.layered_slowness(n, n, n, layer=max(n // 8, 1)).reshape(n, n, n)
    k = np.fft.fftfreq(n)
    kk = np.sqrt(k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2)
    filt = np.exp(-0.5 * (kk / 0.04) ** 2)
    slow = np.empty((nmod, n ** 3))
    for m in range(nmod):
        g = np.fft.ifftn(np.fft.fftn(rng.standard_normal((n, n, n))) * filt).real
        g /= np.abs(g).max()
        slow[m] = (base / (1.0 + 0.05 * g)).ravel()
    xs, ys, zs = cases.interior_sources(nst, n, n, n, H, seed=6)
    fm = np.repeat(np.arange(nmod, dtype=np.int32), nst)
    st = np.tile(np.arange(nst), nmod)
    return slow, fm, xs[st], ys[st], zs[st]


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import mceik_b200
    from mceik_b200 import _lib, sharding
    from mceik_b200.eikonal import EikonalSolver
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    if a.config != 3 and world != 1:
        raise SystemExit("--config 2 / 5 are single-GPU lines")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated non-default stream shared by torch (events) and the library
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = mceik_b200.Context(local, stream=stream.cuda_stream)
    if world > 1:  # the library's own NCCL communicator: rank 0's id goes round with torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(mceik_b200.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    n = a.grid
    N = n ** 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())
    max_over_ranks = lambda x: reduce(x, dist.ReduceOp.MAX) if world > 1 else x
    sum_over_ranks = lambda x: reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    # ---------------- inputs ----------------
    if a.config == 3:
        if a.fields % world:
            raise SystemExit("--fields must be a multiple of --gpus")
        slow_np, fmodel, xs, ys, zs = c3_inputs(a)
    elif a.config == 2:
        s1, xs, ys, zs = c2_inputs(a)
        slow_np, fmodel = s1[None], np.zeros(1, np.int32)
    else:
        slow_np, fmodel, xs, ys, zs = c5_inputs(a)
    nf_tot = fmodel.size
    slow_h = torch.from_numpy(np.ascontiguousarray(slow_np)).pin_memory()
    d_slow = slow_h.cuda()
    ts = np.zeros(nf_tot)
    sol = EikonalSolver(ctx, n, n, n, H, tol=1e-6, maxit=20)
    slots = (nf_tot + world - 1) // world
    # the (replicated) table buffer: for N > 1 the library's own, which the other ranks put their tables into over NVLink
    # as their fields converge (mceik_tables_alloc_replicated)
    d_tab = (ctx.tables_alloc_replicated(world * slots, N) if world > 1
             else torch.empty((world * slots, N), dtype=torch.float32, device="cuda"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if a.config == 2 else None  # 2 x L2
    misfit = C5Misfit(a, ctx, d_tab, fmodel) if a.config == 5 else None
    cost = None

    def fsm_step():
        nonlocal cost
        if flush is not None:
            flush.fill_(1)
        if world > 1:
            iters, _, _ = sol.solve_sharded(d_slow, fmodel, ts, xs, ys, zs, d_tab, cost=cost)
            cost = iters.copy()  # next step: longest first by these iteration counts
        else:
            sol.solve_device(d_slow, fmodel, ts, xs, ys, zs, d_tables=d_tab)
        if misfit is not None:
            misfit.step()
        return sol.node_updates

    for _ in range(a.warmup):
        fsm_step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = mceik_b200.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    updates, sweep_ms, sweep_launches = 0, 0.0, 0
    for _ in range(a.steps):
        updates += fsm_step()
        ms, nl = sol.sweep_stats
        sweep_ms += ms
        sweep_launches += nl
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = mceik_b200.kernel_launch_count() - launches0
    my_ms = e0.elapsed_time(e1)
    dt_ms = max_over_ranks(my_ms)
    tot_updates = sum_over_ranks(float(updates))
    value = tot_updates / (dt_ms * 1e-3) / 1e9
    li = np.asarray(sol.last_iters)
    iters = {"min": int(li.min()), "max": int(li.max()), "mean": float(li.mean())}
    peak, peak_src = measured_peak()
    sweep_updates = updates  # every node-update of the solve happens inside the sweep kernel
    ach = BYTES_PER_UPDATE * sweep_updates / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": "sweep_bricks16_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak if ach else None, "traffic": ncu_traffic("sweep_bricks16_kernel", sweep_updates / max(sweep_launches, 1)),
                "peak_source": peak_src, "launches": sweep_launches,
                "avg_launch_ms": sweep_ms / max(sweep_launches, 1),
                "algorithmic_bytes_per_launch": BYTES_PER_UPDATE * sweep_updates / max(sweep_launches, 1),
                "kernel_share_of_step": sweep_ms / my_ms if my_ms > 0 else None,
                "note": "rank 0's launches; the kernel is bound by instruction issue (fp64 arithmetic), not by HBM: "
                        "DRAM traffic is ~1.1x the algorithmic bytes (profiles/ncu_summary_r2.md)"}

    # ---------------- end to end through the host-pointer C ABI (this rank's share of the fields) ----------------
    if world > 1:
        rk, _, _ = sharding.assign_fields(fmodel, world, cost)
        mine = np.nonzero(rk == rank)[0]
    else:
        mine = np.arange(nf_tot)
    nf = len(mine)
    fm_l, xs_l, ys_l, zs_l, ts_l = fmodel[mine].copy(), xs[mine].copy(), ys[mine].copy(), zs[mine].copy(), ts[mine].copy()
    e2e_fields = nf if a.config != 5 else min(nf, 256)  # config 5: a slice (the full set is 14 GB of fp64 output per step)
    u_h = torch.empty((e2e_fields, N), dtype=torch.float64).pin_memory()
    slow_flat = slow_h.numpy().reshape(-1, N)
    sp = np.arange(e2e_fields + 1, dtype=np.int32)
    iters_h, ferr_h = np.zeros(e2e_fields, np.int32), np.zeros(e2e_fields, np.int32)
    lib = _lib.load()
    P = lambda x, t: x.ctypes.data_as(t)
    nmod_e2e = slow_flat.shape[0] if a.config != 5 else int(fm_l[:e2e_fields].max()) + 1

    def e2e_step():
        rc = lib.mceik_fsm_solve_batched_host(ctx.handle, C.byref(sol.grid), nmod_e2e, P(slow_flat, _lib.c_dbl_p), e2e_fields,
                                              P(fm_l, _lib.c_int_p), P(sp, _lib.c_int_p), P(ts_l, _lib.c_dbl_p),
                                              P(xs_l, _lib.c_dbl_p), P(ys_l, _lib.c_dbl_p), P(zs_l, _lib.c_dbl_p),
                                              C.cast(u_h.data_ptr(), _lib.c_dbl_p), None, 0, P(iters_h, _lib.c_int_p),
                                              P(ferr_h, _lib.c_int_p))
        assert rc == 0, _lib.last_error()
        return sol.node_updates

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_updates = 0
    for _ in range(a.steps):
        e2e_updates += e2e_step()
    barrier()
    e2e_dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": sum_over_ranks(float(e2e_updates)) / e2e_dt / 1e9, "unit": UNIT,
           "h2d_bytes_per_step": int(nmod_e2e * N * 8 + 4 * 8 * e2e_fields), "d2h_bytes_per_step": int(e2e_fields * N * 8),
           "api": "mceik_fsm_solve_batched_host (pinned host slowness in, fp64 fields out)"
                  + (f"; {e2e_fields} of the {nf} fields per step" if e2e_fields != nf else "")}

    # ---------------- CPU legs + parity at the configured size (N = 1 only: the contract; --impl reference covers N > 1) --------
    cpu = parity = None
    if rank == 0 and world == 1 and not a.skip_cpu:
        with tempfile.TemporaryDirectory(prefix="mceik_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
            legs = run_cpu_leg_subprocess(a, "fsm", tmp)
            cpu = cpu_baseline_from(legs)
            ok, detail = True, []
            for pf in legs["parity_fields"]:
                f = pf["field"]
                ref = np.load(pf["path"])
                same = bool(np.array_equal(u_h[f].numpy(), ref)) and int(iters_h[f]) == pf["iterations"]
                ok &= same
                detail.append(f"field {f}: {'0 ulp, ' + str(pf['iterations']) + ' iterations' if same else 'DIFFERS'}")
            parity = {"ok": ok, "checked": f"{len(legs['parity_fields'])}/{nf_tot} fields at {n}^3 against oracle/fsm3d_oracle.c, bit for bit "
                                           f"(fields and iteration counts): " + "; ".join(detail)}
            assert ok, parity
    del u_h

    # ---------------- grid search: events located/s (BASELINE config 4 shard) ----------------
    events = None
    if a.config == 3 and not a.skip_gs:
        del d_tab
        if world > 1:
            barrier()
            ctx.tables_free_replicated()
        torch.cuda.empty_cache()
        events = run_gs(a, ctx, rank, world, barrier, max_over_ranks, sum_over_ranks)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dt_ms / a.steps, "higher_is_better": True, "scaling": "strong" if a.config == 3 else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(a, world), iterations_per_field=iters),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity_checked": parity, "gpu_launches": int(launches),
                "clocks": clocks, "events": events}
        if misfit is not None:
            line["misfit"] = misfit.summary()
        emit(line)
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()


class C5Misfit:
    """Config 5: after every solve, the catalogue misfit of every proposal against its own 32 tables -- 1000 events
    located at fixed nodes, analytic origin time, weighted L2 residual (mceik_catalog_misfit_dev); one fp64 per model."""

    def __init__(self, a, ctx, d_tab, fmodel):
        import torch
        from mceik_b200.locate import catalog_misfit_device
        self.fn, self.ctx, self.d_tab = catalog_misfit_device, ctx, d_tab
        n, nst, ne = a.grid, 32, 1000
        self.nmod, self.nst, self.ne, self.N = a.fields // nst, nst, ne, n ** 3
        rng = np.random.default_rng(7)
        self.node = torch.from_numpy(rng.integers(0, n ** 3, ne).astype(np.int32)).cuda()
        self.tobs = torch.from_numpy(rng.uniform(1.0, 30.0, (ne, nst))).cuda()
        self.var = torch.from_numpy(rng.choice(np.array([0.1, 0.25, 0.5]), (ne, nst))).cuda()
        self.use = torch.from_numpy((rng.uniform(size=(ne, nst)) >= 0.1).astype(np.int32)).cuda()
        self.out = torch.empty(self.nmod, dtype=torch.float64, device="cuda")

    def step(self):
        self.fn(self.ctx, self.d_tab, self.N, self.nmod, self.nst, self.ne, self.node, self.tobs, self.var, self.use, self.out)

    def summary(self):
        o = self.out.cpu().numpy()
        return {"models": int(self.nmod), "events": self.ne, "picks_per_event": self.nst, "misfit_min": float(o.min()), "misfit_max": float(o.max())}


def run_gs(a, ctx, rank, world, barrier, max_over_ranks, sum_over_ranks):
    """Relocation shard of BASELINE config 4: `gs_events` events per GPU against 2*gs_stations fp32 tables on the 256^3
    grid (analytic homogeneous tables generated on the device), all picks, 10 % masked, variances in {0.1, 0.25, 0.5},
    job 2 (analytic origin time)."""
    import torch
    import ctypes as C
    from mceik_b200 import _lib
    from mceik_b200.locate import Locator
    import mceik_b200
    n, X, Y, Z, V = gs_inputs(a, rank)
    N, ntab, ns = n ** 3, len(X), a.gs_stations
    ne = a.gs_events
    d_tables = torch.empty((ntab, N), dtype=torch.float32, device="cuda")
    lib = _lib.load()
    p = lambda x: x.ctypes.data_as(_lib.c_dbl_p)
    rc = lib.mceik_homogeneous_tables_dev(ctx.handle, n, n, n, 0.0, 0.0, 0.0, H, H, H, ntab, p(X), p(Y), p(Z), p(V),
                                          C.c_void_p(d_tables.data_ptr()), N)
    assert rc == 0, _lib.last_error()
    rng = np.random.default_rng(100 + rank)
    true_node = rng.integers(0, N, ne)
    tori = rng.uniform(0, 10, ne)
    d_true = torch.from_numpy(true_node).cuda()
    tobs = torch.empty((ne, ntab), dtype=torch.float64, device="cuda")
    for lo in range(0, ne, 1024):  # tobs = table value at the true node + origin time
        hi = min(lo + 1024, ne)
        tobs[lo:hi] = d_tables[:, d_true[lo:hi]].T.double() + torch.from_numpy(tori[lo:hi]).cuda()[:, None]
    tobs = tobs.view(-1)
    use_h = rng.uniform(size=ne * ntab) >= 0.1
    tid_h = np.where(use_h, np.tile(np.arange(ntab), ne), -1).astype(np.int32)
    var_h = rng.choice(np.array([0.1, 0.25, 0.5]), ne * ntab)
    obs_ptr_h = (np.arange(ne + 1) * ntab).astype(np.int32)
    tobs_h = tobs.cpu().numpy()
    tid, var, obs_ptr = torch.from_numpy(tid_h).cuda(), torch.from_numpy(var_h).cuda(), torch.from_numpy(obs_ptr_h).cuda()
    iopt = torch.empty(ne, dtype=torch.int32, device="cuda")
    t0 = torch.empty(ne, dtype=torch.float64, device="cuda")
    obj = torch.empty(ne, dtype=torch.float64, device="cuda")
    loc = Locator(ctx)
    loc.set_tables_device(d_tables, N)
    stream = torch.cuda.current_stream()
    big = ne > 1024  # a 12 500-event step takes tens of seconds: one timed step after warm-up steps on a 256-event prefix
    steps, warm = (1, 3) if big else (max(1, min(a.steps, 3)), max(1, min(a.warmup, 3)))
    nw = min(ne, 256) if big else ne

    def step(k=ne):
        loc.locate_device(2, k, ntab, obs_ptr[:k + 1], tid, tobs, var, None, iopt, t0, obj)

    for _ in range(warm):
        step(nw)
    barrier()
    l0 = mceik_b200.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    launches = mceik_b200.kernel_launch_count() - l0
    dt = max_over_ranks(e0.elapsed_time(e1)) * 1e-3
    ev_s = sum_over_ranks(float(ne * steps)) / dt
    hit = int((iopt.long() == d_true).sum())
    nuse = int(use_h.sum())
    alg_bytes = nuse * N * 4 * steps  # each needed fp32 table value once per event
    peak, peak_src = measured_peak()
    ach = alg_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "locate_uniform_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": ncu_traffic("locate_uniform_kernel", ne), "peak_source": peak_src,
            "fp64_tflops": 8.0 * nuse * N * steps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
            "note": "tables are reused across the 8 events of a CTA, so the binding limit is the fp64 pipe "
                    "(8 non-fused flops per event x pick x node), not HBM; frac may exceed 1"}
    iopt_h, t0_h, obj_h = iopt.cpu().numpy(), t0.cpu().numpy(), obj.cpu().numpy()
    # end to end: host CSR picks in, host results out, through mceik_locate_batched_host
    t = time.perf_counter()
    for _ in range(steps):
        io_h, _, _ = loc.locate_host(2, obs_ptr_h, tid_h, tobs_h, var_h)
    barrier()
    e2e_dt = max_over_ranks(time.perf_counter() - t)
    assert np.array_equal(io_h, iopt_h)
    e2e = {"value": sum_over_ranks(float(ne * steps)) / e2e_dt, "unit": "events/s",
           "h2d_bytes_per_step": int(obs_ptr_h.nbytes + tid_h.nbytes + tobs_h.nbytes + var_h.nbytes),
           "d2h_bytes_per_step": int(ne * 20), "api": "mceik_locate_batched_host (tables resident in HBM)"}
    cpu = parity = None
    if rank == 0 and world == 1 and not a.skip_cpu:
        nev = min(ne, host_threads())
        del d_tables
        torch.cuda.empty_cache()
        with tempfile.TemporaryDirectory(prefix="mceik_cpu_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
            np.savez(os.path.join(tmp, "gs_events.npz"), tobs=tobs_h.reshape(ne, ntab)[:nev], var=var_h.reshape(ne, ntab)[:nev],
                     use=use_h.reshape(ne, ntab)[:nev].astype(np.int32))
            leg = run_cpu_leg_subprocess(a, "gs", tmp)
        same = (list(iopt_h[:nev]) == leg["iopt"] and list(t0_h[:nev]) == leg["t0"] and list(obj_h[:nev]) == leg["obj"])
        parity = {"ok": bool(same), "checked": f"{nev}/{ne} events at {n}^3 x {ntab} tables against "
                                               f"{'the reference locate.c object (oracle/_ref)' if leg['kind'] == 'reference' else 'oracle/locate_oracle.c'}: "
                                               f"located node, origin time and objective " + ("bit-equal" if same else "DIFFER")}
        assert same, (parity, list(iopt_h[:nev]), leg["iopt"])
        cpu = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    return {"metric": "events_located_per_s", "value": ev_s, "unit": "events/s", "steps": steps, "warmup": warm,
            "ms_per_step": dt / steps * 1e3,
            "config": {"workload": f"BASELINE config 4 per-GPU shard: {ne} events/GPU x {ntab} fp32 tables "
                                   f"({ns} stations x P,S, homogeneous analytic) on {n}^3, all picks, 10% masked, job 2"
                                   + (f"; warm-up on a {nw}-event prefix" if big else ""),
                       "sharding": f"events x{world}", "l2": f"inputs larger than L2: {ntab * N * 4 / 1e9:.1f} GB of tables"},
            "located_on_true_node": f"{hit}/{ne}", "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "parity_checked": parity,
            "gpu_launches": int(launches)}


def emit(line):
    """The one JSON line of the contract, written to the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    # Libraries (NCCL prints its version banner) write to fd 1: keep the real stdout for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.cpu_leg:
        emit(cpu_leg_fsm(args) if args.cpu_leg == "fsm" else cpu_leg_gs(args))
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
