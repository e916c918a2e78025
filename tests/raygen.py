"""Deterministic ray batches for the hit-record gate (gate 1): camera rays, interior rays,
surface-start secondary rays and adversarial cases (axis-parallel, corner/edge, grazing,
signed zeros, zero and NaN directions)."""
import numpy as np

SCENE_INFO = {
    # name: (extent for random origins, eye, look-at)
    "cornell_box": (300.0, (0, 0, 800), (0, 0, 0)),
    "sphere_cornell_box": (300.0, (0, 0, 800), (0, 0, 0)),
    "random_scene": (12.0, (13, 2, 3), (0, 0, 0)),
    "random_scene_list": (12.0, (13, 2, 3), (0, 0, 0)),
    "two_perlin_spheres": (8.0, (13, 2, 3), (0, 0, 0)),
    "light_spheres": (8.0, (13, 2, 3), (0, 0, 0)),
    "earth": (6.0, (13, 2, 3), (0, 0, 0)),
    "textured_lit": (10.0, (13, 2, 3), (0, 0, 0)),
}


def camera_rays(n, eye, lookat, fov_deg, rng):
    eye = np.asarray(eye, np.float64)
    w = eye - np.asarray(lookat, np.float64)
    w /= np.linalg.norm(w)
    u = np.cross([0, 1, 0], w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    h = np.tan(np.radians(fov_deg) / 2)
    s = rng.uniform(-1, 1, n) * h
    t = rng.uniform(-1, 1, n) * h
    d = -w[None, :] + s[:, None] * u[None, :] + t[:, None] * v[None, :]
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = eye
    rays[:, 3:6] = d * rng.uniform(0.5, 20.0, (n, 1))  # directions are not normalised in the reference
    return rays


def interior_rays(n, extent, rng):
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = rng.uniform(-extent, extent, (n, 3))
    rays[:, 3:6] = rng.normal(size=(n, 3))
    rays[:, 6] = rng.uniform(0, 1, n)
    return rays


def adversarial_rays(extent, eye):
    e = float(extent)
    L = []
    axes = np.eye(3)
    for o in [(0, 0, 0), (0.5 * e, -0.25 * e, 0.1 * e), tuple(eye)]:
        for a in axes:
            for sgn in (1.0, -1.0):
                L.append((*o, *(sgn * a), 0.0))          # axis-parallel: two direction components are 0
                d = sgn * a + 0.0
                d[d == 0] = -0.0                         # negative zeros
                L.append((*o, *d, 0.0))
    # towards the corners / edge mid-points of the [-e,e]^3 box from the eye and from the centre
    for o in [tuple(eye), (0.0, 0.0, 0.0)]:
        for cx in (-e, 0.0, e):
            for cy in (-e, 0.0, e):
                for cz in (-e, e):
                    d = np.array([cx, cy, cz], np.float64) - np.array(o, np.float64)
                    if np.any(d != 0):
                        L.append((*o, *d, 0.0))
    # origins exactly on the box planes, directions along / away from the plane
    for k in (-e, e):
        L.append((k, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0))
        L.append((k, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0))
        L.append((0.0, k, 0.0, 0.0, -1.0, 0.0, 0.0))
        L.append((0.0, k, 0.0, 1.0, 1e-7, 0.0, 0.0))
        L.append((0.0, 0.0, k, 0.0, 0.0, 1.0, 0.0))
    # degenerate directions
    L.append((0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
    L.append((1.0, 2.0, 3.0, np.nan, np.nan, np.nan, 0.0))
    L.append((1.0, 2.0, 3.0, np.nan, 1.0, 0.0, 0.0))
    L.append((1.0, 2.0, 3.0, np.inf, 1.0, 0.0, 0.0))
    L.append((0.0, 0.0, 0.0, 1e-30, 1e-30, -1e-30, 0.0))
    L.append((0.0, 0.0, 0.0, 1e30, -1e30, 1e30, 0.0))
    return np.array(L, np.float32)


def secondary_rays(hits, rng):
    """surface-start rays: origin = a reference hit point, direction = normal + random vector"""
    ok = hits["hit"] == 1
    p = hits["p"][ok]
    n = hits["n"][ok]
    m = p.shape[0]
    r = rng.normal(size=(m, 3))
    r /= np.linalg.norm(r, axis=1, keepdims=True)
    rays = np.zeros((m, 7), np.float32)
    rays[:, 0:3] = p
    rays[:, 3:6] = n + 0.999 * r
    return rays


def primary_batch(scene, n_cam, n_int, seed):
    extent, eye, lookat = SCENE_INFO[scene]
    rng = np.random.default_rng(seed)
    fov = 90.0 if "cornell" in scene else 25.0
    return np.concatenate([camera_rays(n_cam, eye, lookat, fov, rng), interior_rays(n_int, extent, rng),
                           adversarial_rays(extent, eye)], axis=0)
