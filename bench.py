#!/usr/bin/env python
"""bench.py -- benchmark of the B200 path-tracing core on the BASELINE.json configurations.

Default workload = BASELINE.json configs[3], the configuration the metric is quoted on: Cornell box with
metal box + glass sphere, 1200x1200, 2048 spp; variant A = the reference's shipped config.ini
(fov 90, recursion depth 15, aperture 0.1, shutter 0-0), camera main.cpp:87-91, light-sampling
list main.cpp:99-106. One "step" = one full render of that frame. `--config {1,2,3a,3b,4,5}` selects
the other BASELINE configurations (same line shape).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]   our arm (CUDA core via the C-ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...        the reference's CPU sample loop
  torchrun --nproc-per-node N bench.py --gpus N ...                  one rank per GPU (static tile split)

Prints ONE JSON line (rank 0). `value` = paths/s over all GPUs with the scene resident in HBM,
timed with CUDA events on the library's stream, max over ranks. `e2e` = the same metric through
the reference-facing call (tpt_scene_create + tpt_render with HOST buffers: scene H2D, image
D2H, multi-rank gather of the tiles each rank owns) by wall clock. `parity_mode` = the same job in
TPT_MODE_PARITY (the arithmetic that is checked bit-for-bit against the reference) at the full
sample count with its own value / e2e / roofline. Under torchrun, `e2e_inprocess` = the same frame
rendered by rank 0 alone driving all N GPUs through tpt_render_multi (static split + work stealing
+ NVLink gather, no NCCL).
"""
import argparse
import hashlib
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 16
VARIANTS = {"A": dict(fov=90.0, depth=15), "B": dict(fov=61.93, depth=50)}
SM_FP32_LANES = 128
BOOK_FOCUS = 13.490737915039062  # float32 sqrt(13^2 + 2^2 + 3^2): the commented camera of main.cpp:82-84

# BASELINE.json configs. f_path = algorithmic FLOPs per path of the REFERENCE algorithm
# (operation weights x measured call multiplicities; fixed per config, independent of how
# much work the GPU skips): configs 4/5 per variant below; 1 and 2 as stated there (486 sphere tests
# per ray through the flat list; 54 box + 12 sphere tests per ray through the bvh_node tree);
# 3a / 3b derived with the same weights: two_perlin_spheres 2.65 rays per path,
# 2 sphere tests per ray, 1.65 lambertian bounces of 230 FLOP + a 5-octave turbulence texture of
# ~750 FLOP; earth 2.13 rays per path, one sphere test with uv, 1.13 bounces, nearest-texel lookup.
CONFIGS = {
    "1": dict(scene="random_scene_list", camera="book", nx=400, ny=400, spp=16, depth=15, fov=20.0, f_path=35e3,
              label="configs[0]: random-spheres scene as a flat hitable_list (no BVH), 400x400 @ 16 spp"),
    "2": dict(scene="random_scene", camera="book", nx=1600, ny=1600, spp=300, depth=15, fov=20.0, f_path=5.5e3,
              label="configs[1]: random-spheres scene through bvh_node, 1600x1600 @ 300 spp (README's 80.1 s case)",
              published_s=80.1),
    "3a": dict(scene="two_perlin_spheres", camera="book", nx=1600, ny=1600, spp=256, depth=15, fov=20.0, f_path=1.85e3,
               label="configs[2]a: two_perlin_spheres (Perlin turbulence texture), 1600x1600 @ 256 spp"),
    "3b": dict(scene="earth", camera="book", nx=1600, ny=1600, spp=256, depth=15, fov=20.0, f_path=0.42e3,
               label="configs[2]b: earth sphere, image_texture(earthmap.jpg 1024x512), 1600x1600 @ 256 spp"),
    "4": dict(scene="cornell_box", camera="cornell", nx=1200, ny=1200, spp=2048,
              label="configs[3]: Cornell box metal+glass 1200x1200 @ 2048 spp (README's 941 s headline)", published_s=941.0),
    "5": dict(scene="cornell_box", camera="cornell", nx=4096, ny=4096, spp=4096,
              label="configs[4]: Cornell box 4096x4096 @ 4096 spp, tile-partitioned"),
}
F_PATH_CORNELL = {"A": 1.2e3, "B": 4.8e3}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="4", choices=sorted(CONFIGS), help="BASELINE.json configuration (default 4 = the headline)")
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--variant", default="A", choices=["A", "B"], help="Cornell configs: A = shipped config.ini, B = frame-filling fov, depth 50")
    ap.add_argument("--kernel", default="wavefront", choices=["wavefront", "mega"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (default: the config's)")
    ap.add_argument("--size", type=int, default=0, help="override the frame edge in pixels (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU leg (cpu_baseline, stock binary, whole-program run)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_mode block")
    ap.add_argument("--no-inprocess", action="store_true", help="skip e2e_inprocess (tpt_render_multi from rank 0) under torchrun")
    ap.add_argument("--bundle-cull", action="store_true",
                    help="headline with the library's default pixel-bundle bounds test on (pixels that cannot see the "
                         "scene are finished untraced, exact); by default the bench traces every path and reports the "
                         "culled variant beside it")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the bounded baseline sample")
    return ap.parse_args()


def resolve_config(args):
    c = dict(CONFIGS[args.config])
    c["id"] = args.config
    if c["camera"] == "cornell":
        v = VARIANTS[args.variant]
        c.update(fov=v["fov"], depth=v["depth"], f_path=F_PATH_CORNELL[args.variant], variant=args.variant)
    if args.size:
        c["nx"] = c["ny"] = args.size
    if args.spp:
        c["spp"] = args.spp
    return c


def workload_text(c, mode=None, kernel=None):
    s = f"{c['label']}; {c['nx']}x{c['ny']} @ {c['spp']} spp, fov {c['fov']}, depth {c['depth']}, aperture 0.1, shutter 0-0"
    if c.get("variant"):
        s += f", variant {c['variant']}"
    if c["camera"] == "book":
        s += ", camera (13,2,3)->(0,0,0) (commented main.cpp:82-84), black background as at the reference's HEAD"
    if mode:
        s += f", {mode} mode, {kernel} kernel"
    return s


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons (B200_PROFILING.md). Started before the warm-up (the
    tool needs ~0.5 s to come up, longer than an 8-GPU timed region); rows are time-stamped and
    only those inside the timed window are used, falling back to warm-up + timed when the window
    holds fewer than three samples (said in the result)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(rows):
            sm, mx, pw, reasons = [], [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1]))
                    mx.append(float(r[2]))
                    pw.append(float(r[3]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, pw, reasons

        inside = [x for x in self.rows if self.t_begin is not None and self.t_begin <= x[0] <= (self.t_end or 1e30)]
        window = "timed"
        if len(inside) < 3:
            inside, window = self.rows, "warmup+timed"
        sm, mx, pw, reasons = parse(inside)
        load = [s for s, p in zip(sm, pw) if p >= statistics.median(pw)] if pw else sm  # samples under load
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ scene helpers
def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def scene_image(c):
    """earthmap.jpg as the reference's stb_image decodes it (fixture made by tests/golden/make_golden.py)"""
    return golden("earth_jpg_decoded")["rgb"] if c["scene"] == "earth" else None


def host_scene(T, c):
    import ctypes as C
    import numpy as np
    perlin = None
    if c["scene"] == "two_perlin_spheres":
        # the reference shuffles its Perlin tables with a wall-clock seed (src/perlin_noise.cc:15-17): any
        # instance is as good as another for throughput; the committed fixture makes the image reproducible
        g = golden("textures")
        perlin = T.PerlinTables()
        rv = np.ascontiguousarray(g["ranvec"], np.float32)
        C.memmove(perlin.ranvec, rv.ctypes.data, rv.nbytes)
        for name, dst in (("perm_x", perlin.perm_x), ("perm_y", perlin.perm_y), ("perm_z", perlin.perm_z)):
            a = np.ascontiguousarray(g[name], np.int32)
            C.memmove(dst, a.ctypes.data, a.nbytes)
    return T.HostScene(c["scene"], image=scene_image(c), perlin=perlin)


def product_camera(T, c):
    aspect = float(c["nx"]) / float(c["ny"])
    if c["camera"] == "cornell":
        return T.make_camera((0, 0, 800), (0, 0, 0), (0, 1, 0), c["fov"], aspect, 0.1, 10.0, 0.0, 0.0)
    return T.make_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), c["fov"], aspect, 0.1, BOOK_FOCUS, 0.0, 0.0)


def ref_camera_args(c):
    if c["camera"] == "cornell":
        return dict(lookfrom=(0, 0, 800), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=c["fov"], aperture=0.1, focus_dist=10.0)
    return dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=c["fov"], aperture=0.1, focus_dist=BOOK_FOCUS)


# ---------------------------------------------------------------------------- CPU reference
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


def ref_bench_run(c, spp, threads, repeats=1):
    """One run of oracle/_ref/ref_bench (the reference's sources + the sample-loop harness as a non-PIC executable,
    oracle/ref_bench_main.cc): [(seconds, paths, threads)] per repeat, scene built once per process."""
    cmd = [REF_BENCH, c["scene"], str(c["nx"]), str(c["ny"]), str(spp), str(c["depth"]), str(c["fov"]), c["camera"], str(threads),
           str(repeats)]
    if c["scene"] == "earth":
        cmd.append(os.path.join(ROOT, "tests", "golden", "earthmap.jpg"))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        raise RuntimeError(f"ref_bench failed: {r.stderr[-300:]}")
    out = []
    for line in r.stdout.splitlines():
        m = re.match(r"seconds ([0-9.eE+-]+) paths (\d+) threads (\d+)", line)
        if m:
            out.append((float(m.group(1)), int(m.group(2)), int(m.group(3))))
    return out


def cpu_reference_sample(c, target_seconds, threads=0):
    """The reference's own sample loop (unmodified color()/hit/scatter compiled from /root/reference, loop restated in
    oracle/ref_harness.cc from main.cpp:115-134) on all host threads, full frame at a bounded spp. Timed through
    oracle/_ref/ref_bench, built non-PIC like the reference's own Release executable."""
    if not os.path.exists(REF_BENCH):
        return None
    if threads == 0:
        # "all the host threads it can use": probe 1x / 2x / 4x the visible core count, keep the fastest
        base = os.cpu_count() or 1
        best = None
        for t in (base, 2 * base, 4 * base):
            sec, _, _ = ref_bench_run(c, 1, t)[0]
            if best is None or sec < best[1]:
                best = (t, sec)
        threads = best[0]
    per_spp = ref_bench_run(c, 1, threads)[0][0]
    spp = int(max(1, min(64, c["spp"], round(target_seconds / max(per_spp, 1e-3)))))
    sec, paths, nthreads = ref_bench_run(c, spp, threads)[0]
    return {"value": paths / sec / 1e6, "unit": "Mpaths/s", "cores": nthreads, "kind": "reference",
            "sample": f"config {c['id']} ({c['scene']}) full {c['nx']}x{c['ny']} frame at {spp} spp ({paths} paths, {sec:.1f} s): "
                      f"reference color()/hit/scatter (oracle/_ref/ref_bench: the reference's sources built like its Release "
                      f"executable, mt19937 drand_r) on {nthreads} threads of {os.cpu_count()} cores",
            "seconds": sec, "spp": spp, "_threads": threads}


def stock_reference_fit(c, budget_s=40.0):
    """The STOCK program (oracle/_ref/Path_tracer = the reference's main.cpp + src/*.cc, its own
    one-std::async-per-row loop, per-pixel mutex and P3 writer) run in a scratch directory with a
    generated config.ini at two sample counts; its own `time:` line (main.cpp:216-221) gives the
    labelled fit T = a + b * spp. Cornell configs only: scene and camera are hard-coded
    in the reference's main()."""
    exe = os.path.join(ROOT, "oracle", "_ref", "Path_tracer")
    if c["scene"] != "cornell_box" or not os.path.exists(exe):
        return None
    nx, ny = c["nx"], c["ny"]
    if nx * ny > 1600 * 1600:
        return None

    def run(spp):
        d = tempfile.mkdtemp(prefix="tpt_stock_")
        try:
            with open(os.path.join(d, "config.ini"), "w") as f:
                f.write(f"[DEFAULT]\nwidth={nx}\nheight={ny}\nsample={spp}\nrecur_depth={c['depth']}\nfov={c['fov']}\n"
                        "bonus_pic=4\nallow_bonus_pic=0\n[BLUR]\naperture=0.1\n[CAM_MOTION]\nstart_time=0.0\nend_time=0.0\n")
            t0 = time.perf_counter()
            r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=600)
            wall = time.perf_counter() - t0
            m = re.search(r"^time:\s*([0-9.eE+-]+)\s*s", r.stdout, re.M)  # not the bogus "estimate time:" line (main.cpp:160-172)
            return (float(m.group(1)) if m else None), wall
        finally:
            shutil.rmtree(d, ignore_errors=True)

    # Both points well above the program's fixed cost (about 2 s here: 1200 thread starts, 1.44 M string-keyed map
    # inserts under a mutex, P3 output): at 1 spp that cost overlaps the rendering and T(spp) is not yet linear
    # (r02 on the GPU box: 1 / 32 / 128 spp = 1.97 / 2.74 / 6.89 s).
    s1 = min(32, c["spp"])
    t1, w1 = run(s1)
    if t1 is None:
        return None
    s2 = int(max(2 * s1, min(4 * s1, c["spp"], s1 * (budget_s - w1) / max(w1, 0.05))))
    t2, w2 = run(s2)
    if t2 is None:
        return None
    b = (t2 - t1) / (s2 - s1)
    a = t1 - b * s1
    full = a + b * c["spp"]
    return {"kind": "stock binary", "a_seconds": a, "b_seconds_per_spp": b, "points": [[s1, t1], [s2, t2]],
            "extrapolated_seconds_full_job": full, "extrapolated_mpaths_per_s": nx * ny * c["spp"] / full / 1e6,
            "cores": os.cpu_count(),
            "note": "oracle/_ref/Path_tracer (unmodified main.cpp: one std::async per row, per-pixel mutex, P3 text output), its own "
                    f"`time:` line at {s1} and {s2} spp; T = a + b*spp is a labelled EXTRAPOLATION to {c['spp']} spp"}


def program_e2e(c, mode):
    """Whole-program time of the drop-in driver (tiny-path-tracer_b200/lib/Path_tracer_b200 = the
    reference's main() flow: config.ini -> scene classes -> flatten -> libtpt.so -> PPM + JPEG), the
    `time:` line it prints like the reference (main.cpp:216-221: dispatch -> last PPM written) and the
    process wall clock, for the text (P3, the reference's format) and the binary (P6) writer."""
    exe = os.path.join(ROOT, "tiny-path-tracer_b200", "lib", "Path_tracer_b200")
    if not os.path.exists(exe):
        return None
    out = {}
    for ppm in ("p3", "p6"):
        d = tempfile.mkdtemp(prefix="tpt_prog_")
        try:
            with open(os.path.join(d, "config.ini"), "w") as f:
                f.write(f"[DEFAULT]\nwidth={c['nx']}\nheight={c['ny']}\nsample={c['spp']}\nrecur_depth={c['depth']}\nfov={c['fov']}\n"
                        "bonus_pic=4\nallow_bonus_pic=0\n[BLUR]\naperture=0.1\n[CAM_MOTION]\nstart_time=0.0\nend_time=0.0\n"
                        f"[SCENE]\nname={c['scene']}\n[GPU]\nmode={mode}\nkernel=wavefront\n[OUTPUT]\nppm={ppm}\njpeg=native\n")
            if c["scene"] == "earth":
                shutil.copy(os.path.join(ROOT, "tests", "golden", "earthmap.jpg"), os.path.join(d, "earthmap.jpg"))
            best = None
            for _ in range(2):  # first run pays the CUDA context + module load of a fresh process as well; keep the better
                t0 = time.perf_counter()
                r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=900)
                wall = time.perf_counter() - t0
                if r.returncode != 0:
                    return {"error": (r.stderr or r.stdout)[-300:]}
                m = re.search(r"^time:\s*([0-9.eE+-]+)\s*s", r.stdout, re.M)  # not the bogus "estimate time:" line (main.cpp:160-172)
                g = re.search(r"gpu render:\s*([0-9.eE+-]+)\s*s", r.stdout)
                up = re.search(r"flatten \+ upload \+ render \+ download:\s*([0-9.eE+-]+)\s*s", r.stdout)
                of = re.search(r"output files:\s*([0-9.eE+-]+)\s*s", r.stdout)
                row = {"time_line_s": float(m.group(1)) if m else None, "process_wall_s": wall,
                       "gpu_render_s": float(g.group(1)) if g else None,
                       "render_call_s": float(up.group(1)) if up else None, "output_files_s": float(of.group(1)) if of else None,
                       "ppm_bytes": os.path.getsize(os.path.join(d, "img.ppm")),
                       "jpg_bytes": os.path.getsize(os.path.join(d, "img.jpg")) if os.path.exists(os.path.join(d, "img.jpg")) else 0}
                if best is None or (row["time_line_s"] or 1e9) < (best["time_line_s"] or 1e9):
                    best = row
            out[ppm] = best
        finally:
            shutil.rmtree(d, ignore_errors=True)
    out["note"] = ("Path_tracer_b200 in a scratch directory, generated config.ini, best of 2 runs; time_line_s = the program's own "
                   "`time:` line (flatten + scene upload + render + PPM write, as main.cpp:108-221 brackets it; render_call_s and "
                   "output_files_s are its two parts); process_wall_s adds process start, CUDA context creation (on a second host "
                   "thread while the scene is built) and the JPEG contact sheet")
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    c = resolve_config(args)
    total = args.steps + args.warmup
    per_step = max(4.0, min(30.0, 150.0 / max(total, 1)))
    first = cpu_reference_sample(c, per_step)
    if first is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_bench is not built"})
        return
    spp = first["spp"]
    nx, ny = c["nx"], c["ny"]
    paths = nx * ny * spp
    runs = ref_bench_run(c, spp, first["_threads"], repeats=total)  # scene built once, `total` renders
    secs = [sec for sec, _, _ in runs[args.warmup:]]
    t = sum(secs)
    value = paths * len(secs) / t / 1e6
    cpu = {"value": value, "unit": "Mpaths/s", "cores": first["cores"], "kind": "reference", "sample": first["sample"]}
    if not args.no_cpu_baseline and total >= 3:  # the stock executable beside the pooled loop (skipped in the tiny contract test)
        try:
            stock = stock_reference_fit(c)
            if stock:
                cpu["stock"] = stock
        except Exception as e:
            cpu["stock"] = {"error": str(e)}
    line = {
        "impl": "reference", "metric": f"{c['scene']} {nx}x{ny} path-tracing throughput", "value": value, "unit": "Mpaths/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": (value / (nx * ny * c["spp"] / c["published_s"] / 1e6)) if c.get("published_s") and not args.size else None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(c) + f"; bounded sample: {spp} spp per step of the {c['spp']}-spp job "
                               "(the reference's rate does not depend on spp)", "spp_per_step": spp, "config": c["id"]},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------- multi-rank tile gather
_OWNED_CACHE = {}


def owned_pixel_index(torch, nx, ny, rank, world, device):
    """flat indices of the pixels whose 16x16 tile t satisfies t % world == rank (the static split
    tpt_render_params.part_index / part_count describes), row-major; computed once per frame shape"""
    key = (nx, ny, rank, world, str(device))
    if key not in _OWNED_CACHE:
        _OWNED_CACHE[key] = _owned_pixel_index(torch, nx, ny, rank, world, device)
    return _OWNED_CACHE[key]


def _owned_pixel_index(torch, nx, ny, rank, world, device):
    tx = (nx + TILE - 1) // TILE
    jj = torch.arange(ny, device=device).view(-1, 1)
    ii = torch.arange(nx, device=device).view(1, -1)
    tile = (jj // TILE) * tx + (ii // TILE)
    return (tile % world == rank).flatten().nonzero().squeeze(1)


def gather_owned_tiles(torch, dist, frame, nx, ny, rank, world):
    """Tile-owned gather: `frame` is [ny*nx, C] on every rank with only the rank's own tiles filled
    in. Every rank packs the pixels it owns (1/world of the frame) and sends just those; rank 0
    scatters them into its frame. Moves (world-1)/world of ONE frame into rank 0 instead of the
    world full frames a reduce-SUM of zero-padded buffers moves."""
    dev = frame.device
    counts = [int(owned_pixel_index(torch, nx, ny, r, world, "cpu").numel()) for r in range(world)]
    cmax = max(counts)
    idx = owned_pixel_index(torch, nx, ny, rank, world, dev)
    packed = torch.zeros((cmax, frame.shape[1]), dtype=frame.dtype, device=dev)
    packed[: counts[rank]] = frame.index_select(0, idx)
    recv = [torch.empty_like(packed) for _ in range(world)] if rank == 0 else None
    dist.gather(packed, recv, dst=0)
    if rank == 0:
        for r in range(1, world):
            frame.index_copy_(0, owned_pixel_index(torch, nx, ny, r, world, dev), recv[r][: counts[r]])
    return frame


# --------------------------------------------------------------------------------- our arm
def source_fingerprint():
    """sha1 over the CUDA sources: profiles/ncu_summary.json records the fingerprint of the build its
    counters were captured from; constants of another build are not applied to live numbers."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "tiny-path-tracer_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, fn), "rb").read())
    return h.hexdigest()[:16]


def run_ours(args, rank, local_rank, world):
    import ctypes as C
    import numpy as np
    import torch
    import tpt_b200 as T

    dist = None
    cpu_group = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")  # the version banner goes to stdout; see emit() for the belt to these braces
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side barrier: no kernel spins on an idle GPU
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    c = resolve_config(args)
    NX, NY, NS = c["nx"], c["ny"], c["spp"]
    npix = NX * NY
    big = npix * NS > 2e10  # configs[4]-sized jobs: fewer repetitions of the side measurements
    hs = host_scene(T, c)
    scene = T.Scene(hs, device=local_rank)
    cam = product_camera(T, c)
    kernel = T.KERNEL_WAVEFRONT if args.kernel == "wavefront" else T.KERNEL_MEGA
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def params_for(mode, cull):
        return T.make_params(NX, NY, NS, c["depth"], mode=mode, seed=0x5EED, part_index=rank, part_count=world,
                             device=local_rank, kernel=kernel, bundle_cull=cull)

    def device_timed(params, steps, warmup):
        """W untimed renders, then K renders timed with CUDA events on the library's stream"""
        for _ in range(max(warmup, 0)):
            scene.render_device(cam, params)
        barrier()
        t0 = time.perf_counter()
        acc = dict(dev_ms=0.0, render_ms=0.0, paths=0, rays=0, launches=0, culled=0)
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            st = scene.render_device(cam, params)
            acc["dev_ms"] += st["render_ms"] + st["resolve_ms"]
            acc["render_ms"] += st["render_ms"]
            acc["paths"] += st["paths"]
            acc["rays"] += st["rays"]
            acc["culled"] += st["culled_paths"]
            acc["launches"] += st["kernel_launches"]
        barrier()
        acc["wall_ms"] = 1e3 * (time.perf_counter() - t0)
        out = {"ranges": st["reserved"][0], "steps": steps}
        for k in ("dev_ms", "render_ms", "wall_ms"):  # time: the slowest rank
            out[k] = max_over_ranks(acc[k])
        for k in ("paths", "rays", "launches", "culled"):  # units: the whole job
            out[k] = sum_over_ranks(float(acc[k]))
        out["value"] = out["paths"] / (out["dev_ms"] * 1e-3) / 1e6
        return out

    # ---- end to end through the reference-facing call: host scene in, host image out -------
    sum_host = torch.zeros((1, NY, NX, 3), dtype=torch.float32).pin_memory()
    rgb_host = torch.zeros((NY, NX, 3), dtype=torch.uint8).pin_memory()
    img = T.Image()
    img.sum_rgb = C.cast(sum_host.data_ptr(), C.POINTER(C.c_float))
    img.rgb8 = C.cast(rgb_host.data_ptr(), C.POINTER(C.c_uint8))

    class _DevBuf:  # zero-copy view of a library-owned device buffer for torch (multi-rank gather)
        def __init__(self, ptr, nbytes, typestr):
            itemsize = int(typestr[-1])
            self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False),
                                             "version": 2}

    def end_to_end(params, steps, warm):
        e2e_s, h2d, d2h, loss = [], 0, 0, 0.0
        for i in range(warm + steps):
            barrier()
            t0 = time.perf_counter()
            s = T.Scene(hs, device=local_rank)  # flattened scene -> HBM
            t_created = time.perf_counter()
            if dist is None:
                T._check(T.lib().tpt_render(s._s, C.byref(cam), C.byref(params), C.byref(img)))
            else:
                # every rank renders its tiles; each sends the pixels of the tiles it OWNS to rank 0 (NCCL gather
                # over NVLink: 1/world of the frame per rank), rank 0 scatters them into its products, ONE download
                T._check(T.lib().tpt_render_device(s._s, C.byref(cam), C.byref(params)))
                sp, sb, rp, rb = s.device_buffers()
                d_sum = torch.as_tensor(_DevBuf(sp, sb, "<f4"), device=dev).view(npix, 3)
                d_rgb = torch.as_tensor(_DevBuf(rp, rb, "|u1"), device=dev).view(npix, 3)
                gather_owned_tiles(torch, dist, d_sum, NX, NY, rank, world)
                gather_owned_tiles(torch, dist, d_rgb, NX, NY, rank, world)
                torch.cuda.synchronize()
                if rank == 0:
                    T._check(T.lib().tpt_render_fetch(s._s, C.byref(img)))
            st2 = s.stats()
            loss = float(sum_host[0, NY // 2, NX // 2, 1])  # the step's result is read on the host
            t_rendered = time.perf_counter()
            s.close()
            barrier()
            if i >= warm:
                e2e_s.append(time.perf_counter() - t0)
            if os.environ.get("TPT_BENCH_DEBUG") and rank == 0:
                sys.stderr.write(f"e2e iter {i}: {time.perf_counter() - t0:.4f} s (create {t_created - t0:.4f}, render+fetch {t_rendered - t_created:.4f}, destroy {time.perf_counter() - t_rendered:.4f})  render_ms {st2['render_ms']:.1f} wall_ms {st2['wall_ms']:.1f} d2h_ms {st2['d2h_ms']:.1f}\n")
            h2d = int(st2["h2d_bytes"])  # flattened scene blob + camera/params launch arguments
            d2h = int(st2["d2h_bytes"]) if rank == 0 else 0
        step = max_over_ranks(statistics.mean(e2e_s))
        return {"seconds_per_step": step, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "checksum": loss,
                "image_sha1": hashlib.sha1(rgb_host.numpy().tobytes()).hexdigest()[:16] if rank == 0 else None,
                "_rgb8": rgb_host.numpy().copy() if rank == 0 else None}

    mode = T.MODE_FAST if args.mode == "fast" else T.MODE_PARITY
    params = params_for(mode, args.bundle_cull)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 0)):
        scene.render_device(cam, params)
    sampler.mark_begin()
    main = device_timed(params, args.steps, 0)
    sampler.mark_end()
    clocks = sampler.stop()
    value = main["value"]
    e2e = end_to_end(params, args.steps, 1 if big else 2)
    e2e_value = (main["paths"] / args.steps) / e2e["seconds_per_step"] / 1e6
    st_main = scene.stats()
    sm_count = st_main["sm_count"]
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    peak_tflops = sm_count * SM_FP32_LANES * 2 * sm_max * 1e6 / 1e12

    def fp32_roofline(m):
        """(modelled TFLOP/s, paths per launch, ms per launch) of the render kernel of measurement m"""
        paths_per_launch = m["paths"] / m["steps"] / world
        ms_per_launch = m["render_ms"] / m["steps"]
        ach = c["f_path"] * paths_per_launch / (ms_per_launch * 1e-3) / 1e12
        return ach, paths_per_launch, ms_per_launch

    # ---- the same job in PARITY mode (the arithmetic that is checked bit-for-bit / to 1e-4 against the
    # reference) at the FULL sample count, with its own e2e and roofline --------------------------
    parity = None
    if args.mode == "fast" and not args.no_parity:
        pp = params_for(T.MODE_PARITY, args.bundle_cull)
        k = 1 if big else args.steps
        pm = device_timed(pp, k, 0 if big else 1)
        pe = end_to_end(pp, 1 if big else min(2, args.steps), 0 if big else 1)
        ach, _, _ = fp32_roofline(pm)
        parity = {"value": pm["value"], "unit": "Mpaths/s", "spp": NS, "steps": k, "ms_per_step": pm["dev_ms"] / k,
                  "mrays_per_s": pm["rays"] / (pm["dev_ms"] * 1e-3) / 1e6,
                  "e2e": {"value": (pm["paths"] / k) / pe["seconds_per_step"] / 1e6, "unit": "Mpaths/s",
                          "seconds_per_step": pe["seconds_per_step"], "h2d_bytes_per_step": pe["h2d_bytes_per_step"],
                          "d2h_bytes_per_step": pe["d2h_bytes_per_step"], "checksum": pe["checksum"]},
                  "roofline": {"bound": "fp32", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                               "flop_per_path": c["f_path"], "kernel": "render_wave_kernel<PAR=1>" if args.kernel == "wavefront" else "render_mega_kernel<PAR=1>"},
                  "sample_ranges_per_pixel": pm["ranges"],
                  "note": "TPT_MODE_PARITY: fp64 where the reference promotes, IEEE div/sqrt, no FMA contraction, the reference's "
                          "bvh_node / hitable_list tree replayed with its tie rules; gates 1-2 hold bit-for-bit / 0 pixels beyond 1e-4"}

    # ---- the same job with the library's default pixel-bundle bounds test (exact: tests/
    # test_gpu_properties.py::test_pixel_bundle_test_is_exact), reported beside the headline --------
    culled = None
    if not args.bundle_cull:
        cm = device_timed(params_for(mode, True), 1 if big else args.steps, 0 if big else 1)
        k = 1 if big else args.steps
        culled = {"value": cm["value"], "unit": "Mpaths/s", "ms_per_step": cm["dev_ms"] / k,
                  "culled_path_fraction": cm["culled"] / cm["paths"], "rays_per_path": cm["rays"] / cm["paths"],
                  "note": "library default (tpt_render_params.reserved[2] = 0): pixels none of whose rays can reach the "
                          "scene's bounds are finished untraced; bit-identical image"}

    # ---- e2e_inprocess: rank 0 alone drives all N GPUs through tpt_render_multi (north_star's design:
    # static split + work-stealing counter + gather, no NCCL); the other ranks idle at a HOST barrier ----
    inproc = None
    if world > 1 and not args.no_inprocess:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                from concurrent.futures import ThreadPoolExecutor
                pool = ThreadPoolExecutor(max_workers=world)

                def make_scenes():  # one host thread per GPU (ctypes releases the GIL inside tpt_scene_create)
                    return list(pool.map(lambda g: T.Scene(hs, device=g), range(world)))

                scenes = make_scenes()
                arr = (C.c_void_p * world)(*[sc._s for sc in scenes])
                p1 = T.make_params(NX, NY, NS, c["depth"], mode=mode, seed=0x5EED, kernel=kernel, bundle_cull=args.bundle_cull)
                secs, last = [], None
                reps = 1 if big else args.steps
                warm_ip = 1 if big else 2  # first call: peer access, pool growth on every GPU of this process
                for i in range(warm_ip + reps):
                    t0 = time.perf_counter()
                    if i >= warm_ip:  # timed iterations pay scene creation on every GPU too, like `e2e`
                        list(pool.map(lambda sc: sc.close(), scenes))
                        scenes = make_scenes()
                        arr = (C.c_void_p * world)(*[sc._s for sc in scenes])
                    T._check(T.lib().tpt_render_multi(arr, world, C.byref(cam), C.byref(p1), C.byref(img)))
                    last = scenes[0].stats()
                    _ = float(sum_host[0, NY // 2, NX // 2, 1])
                    if i >= warm_ip:
                        secs.append(time.perf_counter() - t0)
                sha = hashlib.sha1(rgb_host.numpy().tobytes()).hexdigest()[:16]
                # same Philox stream per (pixel, sample) on both paths; the two split the frame differently (N parts
                # vs 8 N batches), so a pixel's samples are summed in differently sized sub-ranges: the fp32 sums
                # agree to rounding, the 8-bit pictures to +-1 on a few pixels
                diff = np.abs(rgb_host.numpy().astype(np.int16) - e2e["_rgb8"].astype(np.int16))
                step = statistics.mean(secs)
                inproc = {"value": NX * NY * NS / step / 1e6, "unit": "Mpaths/s", "seconds_per_step": step, "n_gpus": world,
                          "batches_per_gpu": [int(x) for x in last.get("multi_batches", [])][:world],
                          "busy_ms_per_gpu": [round(float(x), 3) for x in last.get("multi_busy_ms", [])][:world],
                          "gather_ms": last["resolve_ms"], "render_ms_slowest_gpu": last["render_ms"],
                          "image_sha1": sha, "rgb8_max_abs_diff_vs_torchrun_e2e": int(diff.max()),
                          "rgb8_values_differing": int((diff > 0).sum()), "rgb8_values": int(diff.size),
                          "note": "tpt_render_multi from ONE process (rank 0; the other ranks wait at a gloo barrier): scene created on "
                                  "every GPU, 8 x N batches of interleaved tiles, 7/8 static (one launch per GPU) + work stealing, GPU 0 reads the peers' "
                                  "partial frames over NVLink, one download; no NCCL on this path"}
                for sc in scenes:
                    sc.close()
            except Exception as e:  # reported, never fatal for the contract line
                inproc = {"error": str(e)}
        dist.barrier(group=cpu_group)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel -----------------------------------------------------
    achieved, paths_per_launch, ms_per_launch = fp32_roofline(main)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic, prof_k, prof = None, {}, {}
    kname = "render_wave_kernel" if args.kernel == "wavefront" else "render_mega_kernel"
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        prof_k = prof.get(kname, {})
    except Exception:
        pass
    fingerprint = source_fingerprint()
    capture_matches = bool(prof_k) and prof_k.get("source_fingerprint") == fingerprint and prof_k.get("config", "4A") == f"{c['id']}{c.get('variant', '')}" \
        and prof_k.get("mode", "fast") == args.mode
    if capture_matches:
        traffic = prof_k.get("dram_bytes_per_launch")
    # FP32 FMA throughput measured on this device right now (independent FMA chains at full occupancy,
    # csrc/tpt_render_fast.cu fp32_peak_kernel): the denominator next to the data-sheet product
    fma_peak = None
    try:
        fma_peak = T.fp32_peak(local_rank)["tflops"]
    except Exception as e:
        sys.stderr.write(f"fp32 peak probe failed: {e}\n")
    # issue slots: what actually bounds this kernel (compares, selects, integer RNG and address work issue like
    # FMAs but count no FLOP). warp instructions per path from the committed ncu capture x the paths/s measured
    # here -- applied only when the capture was taken from THIS build of the kernels, same config and mode
    issue = {"stale_capture": True, "capture_fingerprint": prof_k.get("source_fingerprint"), "build_fingerprint": fingerprint,
             "note": "profiles/ncu_summary.json was captured from another build / config: its per-path constants are not applied"}
    if capture_matches and prof_k.get("warp_inst_per_path") and not args.bundle_cull:
        sm_mhz = clocks.get("sm_mhz") or sm_max
        ginst = prof_k["warp_inst_per_path"] * paths_per_launch / (ms_per_launch * 1e-3) / 1e9
        issue = {"warp_inst_per_path": prof_k["warp_inst_per_path"], "achieved_ginst_per_s": ginst,
                 "peak_ginst_per_s": sm_count * 4 * sm_mhz * 1e6 / 1e9, "frac": ginst / (sm_count * 4 * sm_mhz * 1e6 / 1e9),
                 "active_threads_per_inst": prof_k.get("avg_active_threads_per_inst"),
                 "ncu_issue_active_pct": prof_k.get("issue_active_pct"),
                 "capture": {"kernel": prof_k.get("kernel"), "source_fingerprint": prof_k.get("source_fingerprint"), "file": prof_k.get("file")},
                 "source": "profiles/ncu_summary.json (ncu smsp__inst_executed.sum of the same kernel build) x live paths/s; "
                           "peak = SMs x 4 schedulers x sm_mhz under load"}
    acc_bytes = npix / world * 12 * max(1, main["ranges"])  # R accumulator planes x 12 B per owned pixel
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
        "traffic": traffic, "kernel": kname,
        "achieved_is": "MODELLED: the reference algorithm's FLOPs per path (SURVEY 8d, flop_per_path) x the paths/s measured "
                       "live -- not FLOPs the kernel executed (it skips work the reference does: zombie paths, culled boxes); "
                       "`issue` is the executed-instruction view",
        "issue": issue,
        "peak_source": f"{sm_count} SMs x {SM_FP32_LANES} FP32 lanes x 2 FLOP x {sm_max:.0f} MHz (clocks.max.sm); no tensor/HBM "
                       "bound applies (SURVEY 8d): MEASURED_PEAKS.json has no FP32-issue figure, so this is the nominal one; "
                       "peak_measured_fma is the FMA rate the library's own probe kernel reaches on this device",
        "flop_per_path": c["f_path"],
        "peak_measured_fma": fma_peak, "frac_of_measured_fma": (achieved / fma_peak) if fma_peak else None,
        "frac_at_measured_clock": (achieved / (sm_count * SM_FP32_LANES * 2 * clocks["sm_mhz"] * 1e6 / 1e12)) if clocks.get("sm_mhz") else None,
        "hbm": {"algorithmic_bytes_per_launch": acc_bytes, "achieved_gbs": acc_bytes / (ms_per_launch * 1e-3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs", 6650.0), "peak_source": "measured" if peaks else "fallback"},
    }
    cpu = None
    prog = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_sample(c, args.cpu_seconds)
            if cpu:
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
                stock = stock_reference_fit(c)
                if stock:
                    cpu["stock"] = stock
        except Exception as e:  # the checker must never take the bench down
            cpu = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        try:
            if not big:
                del flush
                torch.cuda.empty_cache()
                prog = program_e2e(c, args.mode)
        except Exception as e:
            prog = {"error": str(e)}
    # README.md:22,34 publish wall seconds for configs 4 and 2 at their own sizes (other hardware, settings unstated)
    at_config_size = (NX, NY, NS) == (CONFIGS[c["id"]]["nx"], CONFIGS[c["id"]]["ny"], CONFIGS[c["id"]]["spp"])
    published_mpaths = (NX * NY * NS / c["published_s"] / 1e6) if (c.get("published_s") and at_config_size) else None
    line = {
        "metric": f"{c['scene']} {NX}x{NY} path-tracing throughput", "value": value, "unit": "Mpaths/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["dev_ms"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": (value / published_mpaths) if published_mpaths else None,
        "dtype": "f32" if args.mode == "fast" else "f32+f64", "data": "synthetic",
        "config": {"workload": workload_text(c, args.mode, args.kernel), "config": c["id"],
                   "paths_per_step": main["paths"] / args.steps, "partition": f"static interleaved 16x16 tiles over {world} rank(s)",
                   "l2": "256 MiB memset between steps (flush); scene working set is shared-memory resident",
                   "published_baseline": (f"README.md: {c['published_s']} s on a Xeon E5-2630 v4 = {published_mpaths:.2f} Mpaths/s "
                                          "(fov / depth / build flags unstated)") if published_mpaths else None},
        "wall_seconds_per_step": main["wall_ms"] / args.steps / 1e3, "mrays_per_s": main["rays"] / (main["dev_ms"] * 1e-3) / 1e6,
        "rays_per_path": main["rays"] / main["paths"], "clocks": clocks, "gpu_launches": int(main["launches"]),
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],  # noqa: E501
                "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "seconds_per_step": e2e["seconds_per_step"],
                "checksum": e2e["checksum"], "image_sha1": e2e["image_sha1"],
                "gather": None if world == 1 else "NCCL gather of the tiles each rank owns (1/N of the frame per rank), one D2H on rank 0"},
        "roofline": roofline,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if prog:
        line["program_e2e"] = prog
    if parity:
        line["parity_mode"] = parity
    if culled:
        line["bundle_cull"] = culled
    if inproc:
        line["e2e_inprocess"] = inproc
    line["config"]["paths_traced"] = ("every (pixel, sample) generated and counted; with the bundle test off a primary ray that misses the "
                                      "scene's bounds ends inside the generate step (one box test, no world->hit)" if not args.bundle_cull
                                      else "pixel-bundle bounds test on (library default)")
    emit(line)


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Native libraries write there too (NCCL prints its version
    banner to fd 1 at several NCCL_DEBUG levels), so fd 1 is pointed at stderr for the whole run and the
    result line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-launch one rank per GPU on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_JSON_FD))  # the ranks' fd 1 is the real stdout again; rank 0 claims it itself
    run_ours(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
