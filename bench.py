#!/usr/bin/env python
"""bench.py -- Stokes steps/sec (classify + assemble + PCG + write-back) at 256^3 on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle restatement)

A *step* is one full solveGasSubclass body (exec/HDK_PolyStokes.C:344-584) on the synthetic scene S3
(BASELINE.json configs[2], the single-GPU 256^3 configuration the metric is quoted on).
Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how every number is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stokes_steps_per_sec_256cubed"
UNIT = "steps/s"
SCENE_N = 256


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample_size(steps_total):
    """S3 at reduced resolution, sized so steps_total oracle steps finish within a few minutes."""
    budget = 150.0 / max(1, steps_total)
    for n, est in ((128, 22.0), (96, 9.0), (64, 3.5)):
        if est <= budget:
            return n
    return 48


def ref_sample_size(steps_total):
    """Same for the compiled reference (oracle/_ref/libps_ref_full.so: one job + the three omp sections of its operator apply)."""
    budget = 150.0 / max(1, steps_total)
    for n, est in ((96, 50.0), (80, 28.0), (64, 13.0), (48, 6.0)):
        if est <= budget:
            return n
    return 32


def reference_available():
    try:
        from oracle import ref_full
        return ref_full.available()
    except Exception:
        return False


def run_ref_sample(n, full_counts=None):
    """One full step of the REFERENCE'S OWN solver (all of exec/HDK_PolyStokesSolver*.cpp + lib/, compiled unmodified into
    oracle/_ref/libps_ref_full.so on the HDK stand-in / Eigen facade) on S3 at n^3, timed by the reference's own setup / solve clocks
    (S.cpp setupClockStart..End, S.cpp:760-805) plus the write-back; extrapolated to 256^3 like run_cpu_sample."""
    from polystokes_b200 import scenes
    from oracle.ref_full import RefFull
    sc = scenes.scene_s3(n)
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)      # the reference chats on stdout
    try:
        t0 = time.perf_counter()
        R = RefFull(sc).setup()
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        rc = R.solve()
        t_solve = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1); os.close(saved); os.close(devnull)
    iters = max(1, R.count("iterations") + 1)
    nsys = max(1, R.count("nSystemSize"))
    cg_s = R.real("solveWallclockMs") * 1e-3
    R.close()
    vox_ratio = (SCENE_N / n) ** 3
    if full_counts:
        it_full, n_full = full_counts["iterations"] + 1, full_counts["nSystemSize"]
    else:
        it_full, n_full = iters * SCENE_N / n, nsys * vox_ratio
    t_full = (t_setup + (t_solve - cg_s)) * vox_ratio + (cg_s / iters) * (n_full / nsys) * it_full
    return dict(n=n, seconds=t_setup + t_solve, setup_s=t_setup, solve_s=t_solve, cg_s=cg_s, iterations=iters - 1, nSystemSize=nsys, full_seconds=t_full,
                cores=3, result=rc)


def run_cpu_sample(n, full_counts=None, threads=0):
    """One full oracle step on S3 at n^3; returns measured seconds and the extrapolation to 256^3.

    Extrapolation: setup scales with the voxel count; a CG iteration scales with the
    system size; the iteration count at 256^3 is the one the GPU arm measured (else scaled ~ linearly in n).
    """
    from polystokes_b200 import scenes
    from oracle.oracle import Oracle, lib
    sc = scenes.scene_s3(n)
    t0 = time.perf_counter()
    o = Oracle(sc, threads=threads).setup()
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.solve()
    o.writeback()
    t_solve = time.perf_counter() - t0
    iters = max(1, o.count("iterations") + 1)
    nsys = max(1, o.count("nSystemSize"))
    vox_ratio = (SCENE_N / n) ** 3
    if full_counts:
        it_full, n_full = full_counts["iterations"] + 1, full_counts["nSystemSize"]
    else:
        it_full, n_full = iters * SCENE_N / n, nsys * vox_ratio
    t_full = t_setup * vox_ratio + (t_solve / iters) * (n_full / nsys) * it_full
    out = dict(n=n, seconds=t_setup + t_solve, setup_s=t_setup, solve_s=t_solve, iterations=iters - 1, nSystemSize=nsys,
               full_seconds=t_full, cores=lib().orc_num_threads())
    return out


def emit(obj):
    """The ONE JSON line of the contract, on the real stdout (see main: fd 1 is pointed at stderr while the bench runs)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries chat on stdout (NCCL prints its version banner there when the first communicator is created): keep the
    # real stdout for the result line only and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=SCENE_N, help="grid resolution of S3 (the metric is quoted at 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-reps", type=int, default=30)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    config = {"workload": f"S3 honey-coil style scene {a.n}^3: floor + pool + jet, reduced tiles 16 pad 2, boundary layers 2/2, CG tol 1e-3 "
                          "(BASELINE.json configs[2])", "grid": [a.n] * 3, "tileSize": 16, "tilePadding": 2,
              "l2": "inputs (9 fp32 grids, 0.6 GB at 256^3) and matrices (>2 GB) exceed the 126 MB L2; no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        use_ref = reference_available()
        if use_ref:
            n = ref_sample_size(a.steps + a.warmup) if a.n == SCENE_N else a.n
            run = run_ref_sample
        else:
            n = cpu_sample_size(a.steps + a.warmup) if a.n == SCENE_N else a.n
            run = run_cpu_sample
        for _ in range(a.warmup):
            run(n)
        res = [run(n) for _ in range(a.steps)]
        full = sum(r["full_seconds"] for r in res) / len(res)
        samp = sum(r["seconds"] for r in res) / len(res)
        value = 1.0 / full
        if use_ref:
            kind = "reference"
            sample = (f"the reference's own solver (exec/HDK_PolyStokesSolver*.cpp + lib/, compiled unmodified into oracle/_ref/libps_ref_full.so on the HDK "
                      f"stand-in / Eigen facade; one job + the 3 omp sections of its operator apply) ran the full step on S3 at {n}^3 in {samp:.2f} s "
                      f"(setup {res[0]['setup_s']:.2f} s, {res[0]['iterations']} CG iterations in {res[0]['cg_s']:.2f} s, n={res[0]['nSystemSize']}); scaled to 256^3: "
                      f"setup + write-back x{(SCENE_N / n) ** 3:.1f} voxels, CG time x system-size ratio x iteration ratio")
        else:
            kind = "port"
            sample = (f"oracle (CPU restatement of the reference path, OpenMP) ran the full step on S3 at {n}^3 in {samp:.2f} s "
                      f"({res[0]['iterations']} CG iterations, n={res[0]['nSystemSize']}); scaled to 256^3: setup x{(SCENE_N / n) ** 3:.1f} voxels, "
                      "CG time x system-size ratio x iteration ratio")
        emit(dict({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config, "impl": "reference",
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": res[0]["cores"], "kind": kind, "sample": sample},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import numpy as np
    import torch
    from polystokes_b200 import PolyStokesSolver, scenes, PS_SUCCESS

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; polystokes_b200 has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sc = scenes.scene_s3(a.n)
    solver = PolyStokesSolver.from_scene(sc, device=local)
    part = None
    if world > 1:
        # ONE scene, z-slab decomposed over the ranks: NCCL halo exchange of p / w + two scalar
        # all-reduces per CG iteration inside the library; torch.distributed only carries the NCCL unique id
        part = solver.init_distributed()
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    d_in = [dev(sc.surface), dev(sc.collision), dev(sc.viscosity)]
    d_vel = [dev(v) for v in sc.vel]
    d_cvel = [dev(v) for v in sc.colvel]
    d_out = [v.clone() for v in d_vel]
    d_valid = [torch.zeros_like(v) for v in d_vel]

    def step_device():
        rc = solver.step(d_in[0], d_in[1], d_in[2], d_vel, d_cvel, d_out, d_valid)
        if rc != PS_SUCCESS:
            raise SystemExit(f"bench.py: solver returned {rc}")

    for _ in range(max(a.warmup, 1)):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    solver.timer_start()                   # CUDA event on the solver's own stream (torch events only see torch's stream)
    launches = 0
    dev_ms = 0.0
    for _ in range(a.steps):
        step_device()
        launches += int(solver.stats.gpu_launches)
        dev_ms += sum(solver.stats.stage_ms)
    elapsed = solver.timer_stop() * 1e-3   # records the stop event on that stream and waits for it
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([elapsed, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, wall = float(t[0].item()), float(t[1].item())
    value = a.steps / elapsed              # N > 1: the SAME 256^3 step, slab-decomposed over the ranks (strong scaling)
    stage_ms = solver.stage_ms()
    counts = {k: solver.count(k) for k in ["nCenter", "nActiveVs", "nSystemSize", "regionCount", "nRowsExt", "nTotalDOFs", "iterations"]}

    # ---- end to end through the public API with HOST (pinned) buffers: H2D + D2H inside the timed region
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    h_in = [pin(sc.surface), pin(sc.collision), pin(sc.viscosity)]
    h_vel = [pin(v) for v in sc.vel]
    h_cvel = [pin(v) for v in sc.colvel]
    h_out = [torch.empty_like(v).pin_memory() for v in h_vel]
    h_valid = [torch.empty_like(v).pin_memory() for v in h_vel]
    npv = lambda ts: [t.numpy() for t in ts]

    def step_host():
        rc = solver.step(h_in[0].numpy(), h_in[1].numpy(), h_in[2].numpy(), npv(h_vel), npv(h_cvel), npv(h_out), npv(h_valid))
        if rc != PS_SUCCESS:
            raise SystemExit(f"bench.py: solver returned {rc}")

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host()
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    h2d = sum(t.numel() * 4 for t in h_in + h_vel + h_cvel)
    d2h = sum(t.numel() * 4 for t in h_out + h_valid)
    e2e = {"value": a.steps / e2e_elapsed, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_elapsed / a.steps * 1e3,
           "stage_ms": {k: round(v, 3) for k, v in solver.stage_ms().items()},
           "overlap": "SDFs first; viscosity / velocity / collision velocity H2D under weights + classification; valid D2H under the CG loop; velocity D2H pipelined per axis"}

    # ---- roofline of the dominant kernel (pass 2 = SpMV over K_ext^T, the longest kernel of a CG iteration),
    #      timed live with CUDA events on the solver's stream
    peak, peak_src = measured_peak()
    solver.setup(d_in[0], d_in[1], d_in[2], d_vel, d_cvel)
    kern = {}
    for k in ("pass1", "pass2", "apply", "cg_iteration"):
        ms = solver.time_kernel(k, a.kernel_reps)
        by = solver.kernel_bytes(k) / world          # per-GPU share of the algorithmic bytes (rows are split ~evenly over the slabs)
        kern[k] = {"ms": ms, "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
            traffic = next((v for k, v in tj.items() if "pass2_kernel" in k), None)
    except Exception:
        pass
    roofline = {"kernel": "pass2_kernel (y = -K_ext^T w - 1/2 mu^-1 x, fused p.Ap)", "bound": "hbm", "achieved": kern["pass2"]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": kern["pass2"]["frac"], "traffic": traffic, "peak_source": peak_src, "kernels": kern}

    # ---- CPU baseline: the oracle on the box's host cores, bounded sample (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = run_cpu_sample(128 if a.n >= 128 else a.n, full_counts=counts if a.n == SCENE_N else None)
        port = {"value": 1.0 / r["full_seconds"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": (f"oracle (OpenMP restatement, all host cores) full step on S3 at {r['n']}^3: {r['seconds']:.2f} s (setup {r['setup_s']:.2f} s, {r['iterations']} CG its, "
                           f"n={r['nSystemSize']}); scaled to 256^3 (setup x{(SCENE_N / r['n']) ** 3:.0f} voxels, per-iteration time x system-size ratio, GPU-measured "
                           f"iteration count) = {r['full_seconds']:.1f} s/step")}
        cpu = port
        if reference_available():
            q = run_ref_sample(64 if a.n >= 64 else a.n, full_counts=counts if a.n == SCENE_N else None)
            cpu = {"value": 1.0 / q["full_seconds"], "unit": UNIT, "cores": q["cores"], "kind": "reference",
                   "sample": (f"the reference's own solver (oracle/_ref/libps_ref_full.so: exec/HDK_PolyStokesSolver*.cpp + lib/ compiled unmodified on the HDK stand-in / "
                              f"Eigen facade; one job + the 3 omp sections of its operator apply) full step on S3 at {q['n']}^3: {q['seconds']:.2f} s (setup {q['setup_s']:.2f} s, "
                              f"{q['iterations']} CG its in {q['cg_s']:.2f} s, n={q['nSystemSize']}); scaled to 256^3 (setup + write-back x{(SCENE_N / q['n']) ** 3:.0f} voxels, "
                              f"per-iteration time x system-size ratio, GPU-measured iteration count) = {q['full_seconds']:.1f} s/step"),
                   "port_all_cores": port}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": elapsed / a.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": dict(config, parallelism="single GPU" if world == 1 else
                                                   f"z-slab decomposition over {world} GPUs, cuts {part[4]} (one process per GPU; per CG iteration: halo exchange of p and w + "
                                                   "2 scalar all-reduces, " + ("fused into the kernels over NVLink peer memory" if solver.count("peerTransport") else "NCCL") +
                                                   "; classification replicated)",
                                                   counts=counts, timing="CUDA events on the solver's stream around the K timed steps (max over ranks), bracketed by "
                                                   "barrier + synchronize; wall_ms_per_step = host clock over the same region; "
                                                   "device_ms_per_step = sum of per-stage CUDA-event times"),
               "wall_ms_per_step": wall / a.steps * 1e3, "device_ms_per_step": dev_ms / a.steps, "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
               "cg_iterations": counts["iterations"], "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
               "roofline": roofline, "cpu_baseline": cpu}
        emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
