"""Synthetic scenes S1..S5 plus small test scenes.

The reference's example scenes are Houdini ``.hipnc`` files that cannot be opened offline
(scenes/); these analytic scenes provide the same *kinds* of input the
DOP node receives (exec/HDK_PolyStokes.C:235-246): a liquid ``surface`` SDF, a solid ``collision``
SDF, a ``viscosity`` field (all centre sampled, fp32), face-sampled ``vel`` and ``collisionvel``
(fp32) and a constant density.  Layout is dense, x fastest: ``a[k, j, i]`` (numpy C order with
shape ``(nz, ny, nx)``).

Sign conventions: ``surface < 0`` inside liquid; ``collision < 0`` inside solid.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


@dataclass
class Scene:
    name: str
    nx: int
    ny: int
    nz: int
    dx: float
    dt: float
    density: float
    surface: np.ndarray
    collision: np.ndarray
    viscosity: np.ndarray
    vel: list
    colvel: list
    params: dict = field(default_factory=dict)

    @property
    def res(self):
        return (self.nx, self.ny, self.nz)


def _centres(nx, ny, nz, dx):
    """Cell-centre coordinates as broadcastable fp64 arrays shaped (nz,1,1),(1,ny,1),(1,1,nx)."""
    x = ((np.arange(nx) + 0.5) * dx).reshape(1, 1, nx)
    y = ((np.arange(ny) + 0.5) * dx).reshape(1, ny, 1)
    z = ((np.arange(nz) + 0.5) * dx).reshape(nz, 1, 1)
    return x, y, z


def sd_box(x, y, z, lo, hi):
    """Signed distance to the axis-aligned box [lo, hi] (negative inside)."""
    cx, cy, cz = [(lo[a] + hi[a]) * 0.5 for a in range(3)]
    hx, hy, hz = [(hi[a] - lo[a]) * 0.5 for a in range(3)]
    qx, qy, qz = np.abs(x - cx) - hx, np.abs(y - cy) - hy, np.abs(z - cz) - hz
    outside = np.sqrt(np.maximum(qx, 0) ** 2 + np.maximum(qy, 0) ** 2 + np.maximum(qz, 0) ** 2)
    inside = np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0)
    return outside + inside


def sd_sphere(x, y, z, c, r):
    return np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r


def sd_cylinder_y(x, y, z, cx, cz, r, y0, y1):
    """Capped cylinder along y."""
    d_r = np.sqrt((x - cx) ** 2 + (z - cz) ** 2) - r
    d_y = np.abs(y - 0.5 * (y0 + y1)) - 0.5 * (y1 - y0)
    outside = np.sqrt(np.maximum(d_r, 0) ** 2 + np.maximum(d_y, 0) ** 2)
    inside = np.minimum(np.maximum(d_r, d_y), 0)
    return outside + inside


def face_shapes(nx, ny, nz):
    return [(nz, ny, nx + 1), (nz, ny + 1, nx), (nz + 1, ny, nx)]


def _velocities(nx, ny, nz, dt, seed, noise=0.05):
    """u* = (0, -9.81 dt, 0) + uniform(-noise, noise) on every face (PCG64(seed))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    vel = []
    for a, shp in enumerate(face_shapes(nx, ny, nz)):
        v = rng.uniform(-noise, noise, size=shp).astype(np.float32)
        if a == 1:
            v += np.float32(-9.81 * dt)
        vel.append(np.ascontiguousarray(v))
    colvel = [np.zeros(shp, dtype=np.float32) for shp in face_shapes(nx, ny, nz)]
    return vel, colvel


def _finish(name, n, dx, dt, density, surf, col, visc, seed, params, noise=0.05):
    nx, ny, nz = n
    # keep >= 1 non-liquid cell at every domain face so field border modes never matter
    x, y, z = _centres(nx, ny, nz, dx)
    surf = np.maximum(surf, sd_box(x, y, z, [1.25 * dx] * 3, [(nx - 1.25) * dx, (ny - 1.25) * dx, (nz - 1.25) * dx]))
    vel, colvel = _velocities(nx, ny, nz, dt, seed, noise)
    return Scene(name, nx, ny, nz, dx, dt, density,
                 np.ascontiguousarray(surf, dtype=np.float32), np.ascontiguousarray(np.broadcast_to(col, (nz, ny, nx)), dtype=np.float32),
                 np.ascontiguousarray(np.broadcast_to(visc, (nz, ny, nx)), dtype=np.float32), vel, colvel, params)


DEFAULT_PARAMS = dict(tolerance=1e-3, maxIterations=5000, liquidLayers=2, solidLayers=2,
                      doReduced=1, doTile=1, tileSize=16, tilePadding=2)


def box_scene(n=32, liquid_hi_frac=0.5625, shell=4, visc=100.0, density=1000.0, seed=1001, name=None, **overrides):
    """S1-style: liquid block in a solid box (free surface on top).  n may be int or (nx,ny,nz)."""
    if isinstance(n, int):
        n = (n, n, n)
    nx, ny, nz = n
    dx = 1.0 / max(n)
    dt = 1.0 / 60.0
    x, y, z = _centres(nx, ny, nz, dx)
    lo = [shell * dx] * 3
    hi = [(nx - shell) * dx, (ny - shell) * dx, (nz - shell) * dx]
    col = -sd_box(x, y, z, lo, hi)                      # solid outside the interior box
    top = int(round(liquid_hi_frac * ny))
    surf = sd_box(x, y, z, lo, [hi[0], top * dx, hi[2]])
    params = dict(DEFAULT_PARAMS)
    params.update(overrides)
    return _finish(name or f"box{nx}x{ny}x{nz}", n, dx, dt, density, surf, col, np.float32(visc), seed, params)


def scene_s1(**overrides):
    """S1: 64^3, liquid block [4,60)x[4,36)x[4,60) in a 4-cell solid shell, uniform solve, tol 1e-6."""
    p = dict(doReduced=0, tolerance=1e-6, maxIterations=5000)
    p.update(overrides)
    return box_scene(64, liquid_hi_frac=36 / 64, shell=4, visc=100.0, density=1000.0, seed=1001, name="S1", **p)


def scene_s2(n=128, **overrides):
    """S2: 128^3 cantilever beam [4,100)x[40,88)x[40,88) clamped to a solid slab x<4; tiles 8/1."""
    s = n / 128.0
    nx = ny = nz = n
    dx = 1.0 / n
    dt = 1.0 / 60.0
    x, y, z = _centres(nx, ny, nz, dx)
    col = x - 4 * s * dx                                  # solid slab x < 4 cells
    surf = sd_box(x, y, z, [0.0, 40 * s * dx, 40 * s * dx], [100 * s * dx, 88 * s * dx, 88 * s * dx])
    params = dict(DEFAULT_PARAMS)
    params.update(dict(tileSize=8, tilePadding=1, liquidLayers=2, solidLayers=2, tolerance=1e-3))
    params.update(overrides)
    return _finish("S2", (nx, ny, nz), dx, dt, 1.0, surf, col, np.float32(100.0), 1002, params)


def scene_s3(n=256, **overrides):
    """S3: 256^3 honey-coil style: floor y<8, pool [8,248)x[8,104)x[8,248) + vertical jet r=12; tiles 16/2."""
    s = n / 256.0
    nx = ny = nz = n
    dx = 1.0 / n
    dt = 1.0 / 60.0
    x, y, z = _centres(nx, ny, nz, dx)
    col = y - 8 * s * dx                                  # floor
    pool = sd_box(x, y, z, [8 * s * dx, 0.0, 8 * s * dx], [248 * s * dx, 104 * s * dx, 248 * s * dx])
    jet = sd_cylinder_y(x, y, z, 0.5, 0.5, 12 * s * dx, 100 * s * dx, 248 * s * dx)
    surf = np.minimum(pool, jet)
    params = dict(DEFAULT_PARAMS)
    params.update(dict(tileSize=16, tilePadding=2, liquidLayers=2, solidLayers=2, tolerance=1e-3))
    params.update(overrides)
    return _finish("S3", (nx, ny, nz), dx, dt, 1000.0, surf, col, np.float32(35.0), 1003, params)


def scene_s4(n=384, **overrides):
    """S4: 384^3 pool (y<230) in a 4-cell solid box with 6 solid spheres; tiles 32/3, layers 3/3."""
    s = n / 384.0
    nx = ny = nz = n
    dx = 1.0 / n
    dt = 1.0 / 60.0
    x, y, z = _centres(nx, ny, nz, dx)
    sh = max(1, int(round(4 * s)))
    col = -sd_box(x, y, z, [sh * dx] * 3, [(n - sh) * dx] * 3)
    rng = np.random.Generator(np.random.PCG64(1004))
    for _ in range(6):
        r = rng.uniform(20, 40) * s * dx
        c = [rng.uniform(0.2, 0.8), rng.uniform(0.15, 0.5), rng.uniform(0.2, 0.8)]
        col = np.minimum(col, sd_sphere(x, y, z, c, r))
    surf = np.broadcast_to(y - 230 * s * dx, (nz, ny, nx))
    params = dict(DEFAULT_PARAMS)
    params.update(dict(tileSize=32, tilePadding=3, liquidLayers=3, solidLayers=3, tolerance=1e-4, maxIterations=10000))
    params.update(overrides)
    return _finish("S4", (nx, ny, nz), dx, dt, 1000.0, surf, col, np.float32(2000.0), 1004, params)


def scene_s5(scale=1.0, **overrides):
    """S5: 512x256x256 ellipsoidal blob on a floor, variable viscosity 400*exp(0.7 s(x)); layers 3/3."""
    nx, ny, nz = int(512 * scale), int(256 * scale), int(256 * scale)
    dx = 1.0 / (256 * scale)
    dt = 1.0 / 60.0
    x, y, z = _centres(nx, ny, nz, dx)
    col = y - 8 * scale * dx
    a, b, c = 200 * scale * dx, 90 * scale * dx, 90 * scale * dx
    cx, cy, cz = nx * dx * 0.5, (8 * scale * dx + b * 0.9), nz * dx * 0.5
    k0 = np.sqrt(((x - cx) / a) ** 2 + ((y - cy) / b) ** 2 + ((z - cz) / c) ** 2)
    surf = (k0 - 1.0) * min(a, b, c)                      # approximate ellipsoid distance
    rng = np.random.Generator(np.random.PCG64(1005))
    sfield = np.zeros((nz, ny, nx))
    for _ in range(8):
        kx, ky, kz = rng.uniform(0.5, 3.0, size=3) * 2 * np.pi
        ph = rng.uniform(0, 2 * np.pi)
        sfield = sfield + np.sin(kx * x / (nx * dx) + ky * y + kz * z + ph) / 8.0
    visc = (400.0 * np.exp(0.7 * sfield)).astype(np.float32)
    params = dict(DEFAULT_PARAMS)
    params.update(dict(tileSize=16, tilePadding=3, liquidLayers=3, solidLayers=3, tolerance=1e-3, maxIterations=10000))
    params.update(overrides)
    return _finish("S5", (nx, ny, nz), dx, dt, 1000.0, surf, col, visc, 1005, params)


def blob_scene(n=40, seed=7, tile=8, pad=1, **overrides):
    """Irregular test scene: liquid = union of overlapping spheres resting in a solid bowl with a pillar.

    Small but nasty: partial weights everywhere, ragged reduced regions, region-boundary fixing,
    small-region removal and variable viscosity are all exercised.
    """
    if isinstance(n, int):
        n = (n, n, n)
    nx, ny, nz = n
    dx = 1.0 / max(n)
    dt = 1.0 / 60.0
    x, y, z = _centres(nx, ny, nz, dx)
    rng = np.random.Generator(np.random.PCG64(seed))
    L = [nx * dx, ny * dx, nz * dx]
    col = -sd_box(x, y, z, [2.3 * dx] * 3, [L[0] - 2.3 * dx, L[1] - 2.3 * dx, L[2] - 2.3 * dx])
    col = np.minimum(col, sd_cylinder_y(x, y, z, 0.31 * L[0], 0.64 * L[2], 0.07 * L[0], 0.0, 0.45 * L[1]))
    surf = None
    for _ in range(5):
        c = [rng.uniform(0.3, 0.7) * L[0], rng.uniform(0.25, 0.5) * L[1], rng.uniform(0.3, 0.7) * L[2]]
        r = rng.uniform(0.18, 0.3) * min(L)
        d = sd_sphere(x, y, z, c, r)
        surf = d if surf is None else np.minimum(surf, d)
    visc = (50.0 * np.exp(0.5 * np.sin(7.0 * x + 3.0 * y) * np.cos(5.0 * z))).astype(np.float32)
    params = dict(DEFAULT_PARAMS)
    params.update(dict(tileSize=tile, tilePadding=pad, tolerance=1e-4))
    params.update(overrides)
    sc = _finish(f"blob{nx}x{ny}x{nz}_s{seed}", n, dx, dt, 800.0, surf, col, visc, seed + 100, params)
    # moving solid: give the collision velocity a rigid translation so the solid-boundary RHS terms run
    for a, v in enumerate((0.02, -0.01, 0.015)):
        sc.colvel[a][...] = np.float32(v)
    return sc


SCENES = {"S1": scene_s1, "S2": scene_s2, "S3": scene_s3, "S4": scene_s4, "S5": scene_s5}
