"""Quick performance probe on the GPU box: one full step of a scene + kernel roofline numbers."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from polystokes_b200 import PolyStokesSolver, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="S3")
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--override", default="{}")
a = ap.parse_args()
ov = json.loads(a.override)
t0 = time.time()
if a.scene == "S3": sc = scenes.scene_s3(a.n, **ov)
elif a.scene == "S2": sc = scenes.scene_s2(a.n, **ov)
elif a.scene == "S1": sc = scenes.scene_s1(**ov)
elif a.scene == "S4": sc = scenes.scene_s4(a.n, **ov)
elif a.scene == "S5": sc = scenes.scene_s5(a.n / 256.0, **ov)
else: sc = scenes.box_scene(a.n, **ov)
print(f"scene {sc.name} {sc.res} built in {time.time()-t0:.1f}s params {sc.params}", flush=True)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
surf, col, visc = d(sc.surface), d(sc.collision), d(sc.viscosity)
vel = [d(v) for v in sc.vel]; cvel = [d(v) for v in sc.colvel]
vout = [v.clone() for v in vel]; valid = [torch.zeros_like(v) for v in vel]
s = PolyStokesSolver.from_scene(sc)
for it in range(a.steps):
    torch.cuda.synchronize(); t = time.perf_counter()
    rc = s.step(surf, col, visc, vel, cvel, vout, valid)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"step {it}: rc={rc} wall={dt*1e3:.1f} ms iters={s.count('iterations')} err={s.real('solveError'):.3e} launches={s.stats.gpu_launches}", flush=True)
    print("   stages ms:", {k: round(v, 2) for k, v in s.stage_ms().items()}, flush=True)
print("counts:", {k: s.count(k) for k in ["nCenter","nActiveVs","nSystemSize","regionCount","nRowsExt","nTotalDOFs","fixLoops"]}, flush=True)
peak = 6448.7
try: peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
for k in ["pass1", "pass1_sweep", "pass2", "pass2_dots", "apply", "cg_update", "cg_iteration"]:
    ms = s.time_kernel(k, a.reps); by = s.kernel_bytes(k)
    csr = s.kernel_bytes("csr_" + k)
    extra = f"   [as CSR SpMV (12 B/nnz): {csr/1e9:6.3f} GB -> {csr/ms/1e6:8.1f} GB/s]" if csr else ""
    print(f"{k:13s} {ms:8.4f} ms  {by/1e9:7.3f} GB algorithmic  -> {by/ms/1e6:8.1f} GB/s  ({by/ms/1e6/peak*100:5.1f}% of {peak} GB/s measured peak){extra}", flush=True)
print("velocity out max:", [float(v.abs().max()) for v in vout], "valid:", [int(v.sum()) for v in valid])
