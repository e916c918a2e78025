#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing core.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): Cornell box with
metal box + glass sphere, 1200x1200, 2048 spp; variant A = the reference's shipped config.ini
(fov 90, recursion depth 15, aperture 0.1, shutter 0-0), camera main.cpp:87-91, light-sampling
list main.cpp:99-106. One "step" = one full render of that frame.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA core via the C-ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's CPU sample loop
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU (static tile split)

Prints ONE JSON line (rank 0). `value` = paths/s over all GPUs with the scene resident in HBM,
timed with CUDA events on the library's stream, max over ranks. `e2e` = the same metric through
the reference-facing call (tpt_scene_create + tpt_render with HOST buffers: scene H2D, image
D2H, multi-rank host gather) by wall clock.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = NY = 1200
NS = 2048
VARIANTS = {"A": dict(fov=90.0, depth=15), "B": dict(fov=61.93, depth=50)}
# algorithmic FLOPs per path of the REFERENCE algorithm (operation weights x
# measured call multiplicities); fixed per config, independent of how much work the GPU skips
F_PATH = {"A": 1.2e3, "B": 4.8e3}
PUBLISHED_MPATHS = 2.949e9 / 941.0 / 1e6  # README.md:22, 941 s on a Xeon E5-2630 v4 (fov/depth unstated)
SM_FP32_LANES = 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "parity"])
    ap.add_argument("--variant", default="A", choices=["A", "B"])
    ap.add_argument("--kernel", default="wavefront", choices=["wavefront", "mega"])
    ap.add_argument("--spp", type=int, default=NS, help="override samples per pixel (default: the headline 2048)")
    ap.add_argument("--size", type=int, default=NX, help="frame edge in pixels (default 1200; BASELINE configs[4] is 4096 with --spp 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bundle-cull", action="store_true",
                    help="headline with the library's default pixel-bundle bounds test on (pixels that cannot see the "
                         "scene are finished untraced, exact); by default the bench traces every path and reports the "
                         "culled variant beside it")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the bounded baseline sample")
    return ap.parse_args()


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


# ---------------------------------------------------------------------------- CPU reference
def cpu_reference_sample(variant, target_seconds, threads=0):
    """The reference's own sample loop (unmodified color()/hit/scatter compiled from
    /root/reference into oracle/_ref/libtptref.so, loop restated in oracle/ref_harness.cc from
    main.cpp:115-134) on all host threads, full 1200x1200 frame at a bounded spp."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ref as O
    if not O.available(False):
        return None
    v = VARIANTS[variant]
    rs = O.RefScene("cornell_box", det=False)
    cam = dict(O.CORNELL_CAM, vfov=v["fov"])
    _, _, st = rs.render(cam, NX, NY, 1, v["depth"], deterministic=False, threads=threads, count_rays=False)
    per_spp = st["seconds"]
    spp = int(max(1, min(64, round(target_seconds / max(per_spp, 1e-3)))))
    _, _, st = rs.render(cam, NX, NY, spp, v["depth"], deterministic=False, threads=threads, count_rays=False)
    return {"value": st["paths"] / st["seconds"] / 1e6, "unit": "Mpaths/s", "cores": st["threads"], "kind": "reference",
            "sample": f"Cornell variant {variant} full {NX}x{NY} frame at {spp} spp ({st['paths']} paths, {st['seconds']:.1f} s): "
                      f"reference color()/hit/scatter (oracle/_ref/libtptref.so, mt19937 drand_r) on {st['threads']} threads",
            "seconds": st["seconds"], "spp": spp}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    v = VARIANTS[args.variant]
    total = args.steps + args.warmup
    per_step = max(4.0, min(30.0, 150.0 / max(total, 1)))
    first = cpu_reference_sample(args.variant, per_step)
    if first is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libtptref.so is not built"})
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ref as O
    rs = O.RefScene("cornell_box", det=False)
    cam = dict(O.CORNELL_CAM, vfov=v["fov"])
    spp = first["spp"]
    secs = []
    paths = NX * NY * spp
    for i in range(total):
        _, _, st = rs.render(cam, NX, NY, spp, v["depth"], deterministic=False, threads=0, count_rays=False)
        if i >= args.warmup:
            secs.append(st["seconds"])
    t = sum(secs)
    value = paths * len(secs) / t / 1e6
    line = {
        "impl": "reference", "metric": f"Cornell {NX}x{NY} path-tracing throughput", "value": value, "unit": "Mpaths/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": value / PUBLISHED_MPATHS, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"Cornell box metal+glass {NX}x{NY}, variant {args.variant} (fov {v['fov']}, depth {v['depth']}), "
                               f"bounded sample {spp} spp per step of the 2048-spp job", "spp_per_step": spp},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": first["cores"], "kind": "reference",
                         "sample": first["sample"]},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import tpt_b200 as T

    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ.pop("NCCL_DEBUG")  # the version banner goes to stdout; see emit() for the belt to these braces
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
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

    v = VARIANTS[args.variant]
    mode = T.MODE_FAST if args.mode == "fast" else T.MODE_PARITY
    hs = T.HostScene("cornell_box")
    scene = T.Scene(hs, device=local_rank)
    cam = T.cornell_camera(NX, NY, fov=v["fov"])
    kernel = T.KERNEL_WAVEFRONT if args.kernel == "wavefront" else T.KERNEL_MEGA
    params = T.make_params(NX, NY, args.spp, v["depth"], mode=mode, seed=0x5EED, part_index=rank, part_count=world,
                           device=local_rank, kernel=kernel, bundle_cull=args.bundle_cull)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 0)):
        scene.render_device(cam, params)

    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    dev_ms, render_ms, paths, rays, launches = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        st = scene.render_device(cam, params)
        dev_ms += st["render_ms"] + st["resolve_ms"]
        render_ms += st["render_ms"]
        paths += st["paths"]
        rays += st["rays"]
        launches += st["kernel_launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.mark_end()
    clocks = sampler.stop()

    dev_ms_max = max_over_ranks(dev_ms)
    render_ms_max = max_over_ranks(render_ms)
    wall_ms_max = max_over_ranks(wall_ms)
    total_paths = sum_over_ranks(float(paths))
    total_rays = sum_over_ranks(float(rays))
    value = total_paths / (dev_ms_max * 1e-3) / 1e6

    # ---- end to end through the reference-facing call: host scene in, host image out -------
    npix = NX * NY
    sum_host = torch.zeros((1, NY, NX, 3), dtype=torch.float32).pin_memory()
    rgb_host = torch.zeros((NY, NX, 3), dtype=torch.uint8).pin_memory()
    import ctypes as C
    img = T.Image()
    img.sum_rgb = C.cast(sum_host.data_ptr(), C.POINTER(C.c_float))
    img.rgb8 = C.cast(rgb_host.data_ptr(), C.POINTER(C.c_uint8))
    e2e_s = []
    h2d = d2h = 0
    class _DevBuf:  # zero-copy view of a library-owned device buffer for torch (NCCL gather)
        def __init__(self, ptr, nbytes, typestr):
            itemsize = int(typestr[-1])
            self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False),
                                             "version": 2}

    E2E_WARM = 2  # untimed end-to-end iterations (pool growth, first-use paths)
    for i in range(E2E_WARM + args.steps):
        barrier()
        t0 = time.perf_counter()
        s = T.Scene(hs, device=local_rank)  # flattened scene -> HBM
        t_created = time.perf_counter()
        if dist is None:
            T._check(T.lib().tpt_render(s._s, C.byref(cam), C.byref(params), C.byref(img)))
        else:
            # every rank renders its tiles; the disjoint per-GPU products are combined over NVLink
            # (tiles a rank does not own are zero, so SUM is an exact gather), then ONE download on rank 0
            T._check(T.lib().tpt_render_device(s._s, C.byref(cam), C.byref(params)))
            sp, sb, rp, rb = s.device_buffers()
            d_sum = torch.as_tensor(_DevBuf(sp, sb, "<f4"), device=dev)
            d_rgb = torch.as_tensor(_DevBuf(rp, rb, "|u1"), device=dev)
            dist.reduce(d_sum, dst=0, op=dist.ReduceOp.SUM)
            dist.reduce(d_rgb, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                T._check(T.lib().tpt_render_fetch(s._s, C.byref(img)))
        st2 = s.stats()
        loss = float(sum_host[0, NY // 2, NX // 2, 1])  # the step's result is read on the host
        t_rendered = time.perf_counter()
        s.close()
        barrier()
        if i >= E2E_WARM:
            e2e_s.append(time.perf_counter() - t0)
        if os.environ.get("TPT_BENCH_DEBUG") and rank == 0:
            sys.stderr.write(f"e2e iter {i}: {time.perf_counter() - t0:.4f} s (create {t_created - t0:.4f}, render+fetch {t_rendered - t_created:.4f}, destroy {time.perf_counter() - t_rendered:.4f})  render_ms {st2['render_ms']:.1f} wall_ms {st2['wall_ms']:.1f} d2h_ms {st2['d2h_ms']:.1f}\n")
        h2d = int(st2["h2d_bytes"])  # flattened scene blob + camera/params launch arguments
        d2h = int(st2["d2h_bytes"]) if rank == 0 else 0
    e2e_step = max_over_ranks(statistics.mean(e2e_s))
    e2e_value = (total_paths / args.steps) / e2e_step / 1e6

    # ---- the same job in PARITY mode (the arithmetic that is checked bit-for-bit / to 1e-4 against the
    # reference), 256 of the 2048 spp: reported next to the headline, not instead of it -------------
    parity = None
    if args.mode == "fast":
        pp = T.make_params(NX, NY, min(256, args.spp), v["depth"], mode=T.MODE_PARITY, seed=0x5EED, part_index=rank,
                           part_count=world, device=local_rank, kernel=kernel, bundle_cull=args.bundle_cull)
        scene.render_device(cam, pp)
        barrier()
        stp = scene.render_device(cam, pp)
        p_ms = max_over_ranks(stp["render_ms"] + stp["resolve_ms"])
        p_paths = sum_over_ranks(float(stp["paths"]))
        parity = {"value": p_paths / (p_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "spp": min(256, args.spp),
                  "note": "TPT_MODE_PARITY: fp64 where the reference promotes, no FMA contraction, reference BVH walk"}

    # ---- the same job with the library's default pixel-bundle bounds test (exact: tests/
    # test_gpu_properties.py::test_pixel_bundle_test_is_exact), reported beside the headline --------
    culled = None
    if not args.bundle_cull:
        pc = T.make_params(NX, NY, args.spp, v["depth"], mode=mode, seed=0x5EED, part_index=rank, part_count=world,
                           device=local_rank, kernel=kernel, bundle_cull=True)
        scene.render_device(cam, pc)
        barrier()
        c_ms, c_paths, c_culled, c_rays = 0.0, 0, 0, 0
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            stc = scene.render_device(cam, pc)
            c_ms += stc["render_ms"] + stc["resolve_ms"]
            c_paths += stc["paths"]
            c_culled += stc["culled_paths"]
            c_rays += stc["rays"]
        barrier()
        c_ms = max_over_ranks(c_ms)
        c_paths, c_culled, c_rays = sum_over_ranks(float(c_paths)), sum_over_ranks(float(c_culled)), sum_over_ranks(float(c_rays))
        culled = {"value": c_paths / (c_ms * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": c_ms / args.steps,
                  "culled_path_fraction": c_culled / c_paths, "rays_per_path": c_rays / c_paths,
                  "note": "library default (tpt_render_params.reserved[2] = 0): pixels none of whose rays can reach the "
                          "scene's bounds are finished untraced; bit-identical image"}

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (render_mega_kernel) --------------------------------
    st = scene.stats()
    sm_count = st["sm_count"]
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    peak_tflops = sm_count * SM_FP32_LANES * 2 * sm_max * 1e6 / 1e12
    paths_per_launch = total_paths / args.steps / world
    ms_per_launch = render_ms_max / args.steps
    achieved = F_PATH[args.variant] * paths_per_launch / (ms_per_launch * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic, prof_k = None, {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        prof_k = prof.get("render_wave_kernel" if args.kernel == "wavefront" else "render_mega_kernel", {})
        traffic = prof_k.get("dram_bytes_per_launch")
    except Exception:
        pass
    # FP32 FMA throughput measured on this device right now (independent FMA chains at full occupancy,
    # csrc/tpt_render_fast.cu fp32_peak_kernel): the denominator next to the data-sheet product
    fma_peak = None
    try:
        fma_peak = T.fp32_peak(local_rank)["tflops"]
    except Exception as e:
        sys.stderr.write(f"fp32 peak probe failed: {e}\n")
    # issue slots: what actually bounds this kernel (compares, selects, integer RNG and address work
    # issue like FMAs but count no FLOP). warp instructions per path from the committed ncu capture of
    # the same kernel and variant x the paths/s measured here, against 4 warp instructions per clock and SM
    issue = None
    if prof_k.get("warp_inst_per_path") and args.variant == "A" and args.mode == "fast" and not args.bundle_cull:
        sm_mhz = clocks.get("sm_mhz") or sm_max
        ginst = prof_k["warp_inst_per_path"] * paths_per_launch / (ms_per_launch * 1e-3) / 1e9
        issue = {"warp_inst_per_path": prof_k["warp_inst_per_path"], "achieved_ginst_per_s": ginst,
                 "peak_ginst_per_s": sm_count * 4 * sm_mhz * 1e6 / 1e9, "frac": ginst / (sm_count * 4 * sm_mhz * 1e6 / 1e9),
                 "active_threads_per_inst": prof_k.get("avg_active_threads_per_inst"),
                 "source": "profiles/ncu_summary.json (ncu smsp__inst_executed.sum of the same kernel) x live paths/s; "
                           "peak = SMs x 4 schedulers x sm_mhz under load"}
    acc_bytes = npix / world * 12 * max(1, st["reserved"][0])  # R accumulator planes x 12 B per owned pixel
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
        "traffic": traffic, "kernel": "render_wave_kernel" if args.kernel == "wavefront" else "render_mega_kernel",
        "peak_source": f"{sm_count} SMs x {SM_FP32_LANES} FP32 lanes x 2 FLOP x {sm_max:.0f} MHz (clocks.max.sm); no tensor/HBM "
                       "bound applies (SURVEY 8d): MEASURED_PEAKS.json has no FP32-issue figure, so this is the nominal one; "
                       "peak_measured_fma is the FMA rate the library's own probe kernel reaches on this device",
        "flop_per_path": F_PATH[args.variant],
        "peak_measured_fma": fma_peak, "frac_of_measured_fma": (achieved / fma_peak) if fma_peak else None,
        "issue": issue,
        "frac_at_measured_clock": (achieved / (sm_count * SM_FP32_LANES * 2 * clocks["sm_mhz"] * 1e6 / 1e12)) if clocks.get("sm_mhz") else None,
        "hbm": {"algorithmic_bytes_per_launch": acc_bytes, "achieved_gbs": acc_bytes / (ms_per_launch * 1e-3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs", 6650.0), "peak_source": "measured" if peaks else "fallback"},
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_sample(args.variant, args.cpu_seconds)
            if cpu:
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker must never take the bench down
            cpu = {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    line = {
        "metric": f"Cornell {NX}x{NY} path-tracing throughput", "value": value, "unit": "Mpaths/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": value / PUBLISHED_MPATHS, "dtype": "f32" if args.mode == "fast" else "f32+f64",
        "data": "synthetic",
        "config": {"workload": f"Cornell box metal+glass {NX}x{NY} @ {args.spp} spp, variant {args.variant} "
                               f"(fov {v['fov']}, depth {v['depth']}, aperture 0.1), {args.mode} mode, {args.kernel} kernel",
                   "paths_per_step": total_paths / args.steps, "partition": f"static interleaved 16x16 tiles over {world} rank(s)",
                   "l2": "256 MiB memset between steps (flush); scene working set is shared-memory resident",
                   "published_baseline": "README.md:22: 941 s on Xeon E5-2630 v4 = 3.13 Mpaths/s (fov/depth unstated)"},
        "wall_seconds_per_step": wall_ms_max / args.steps / 1e3, "mrays_per_s": total_rays / (dev_ms_max * 1e-3) / 1e6,
        "rays_per_path": total_rays / total_paths, "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "seconds_per_step": e2e_step, "checksum": loss},
        "roofline": roofline,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if parity:
        line["parity_mode"] = parity
    if culled:
        line["bundle_cull"] = culled
    line["config"]["paths_traced"] = ("every (pixel, sample) traced (bundle test off)" if not args.bundle_cull
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
    global NX, NY
    args = parse()
    claim_stdout()
    NX = NY = args.size
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
