"""Secondary measurements printed under "extras" in bench.py's JSON line: the other BASELINE.json configs on one GPU
(inputs resident in HBM, CUDA events, after warm-up). Each entry carries its own algorithmic-byte roofline figure
(SURVEY.md §8d). These are informational; the headline metric is bench.py's terrain view batch."""
import ctypes as C
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _time(torch, fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / iters


def cpu_reference(out):
    """The unmodified reference (oracle/_ref SSE2 build, its own worker threads) on the host cores for configs 3 and 5, written next to the
    device numbers (reported baseline only; oracle/_ref is test infrastructure). The reference renders tiny triangles through
    renderer_begin / giveTask / end and runs its filters single-threaded; filter_mapRgbaU8 is timed with the wrapper's C++ lambda."""
    import refbind
    import sandbox_scene
    from dfpsr_b200 import abi, scenes
    if not refbind.available("sse"):
        return
    ref = refbind.Ref("sse")
    threads = int(ref.lib.ref_thread_count())
    if "tiny_triangles_4k" in out:
        nx, nz = 1000, 999
        sc = scenes.tiny_triangle_scene(nx, nz)
        model = ref.model(sc["points"], sc["polygons"])
        col, dep = ref.rgba(shape=(2160, 3840)), ref.f32(shape=(2160, 3840))
        cam = scenes.top_down_camera(nx, nz, 3840, 2160)
        times = []
        for _ in range(3):
            ref.lib.ref_image_fill_rgba(col, 0, 0, 0, 0)
            ref.lib.ref_image_fill_f32(dep, 0.0)
            t0 = time.perf_counter()
            ref.render(model, cam, col, dep, mode=1)
            times.append(1000.0 * (time.perf_counter() - t0))
        ref.free_all()
        e = out["tiny_triangles_4k"]
        e["cpu_reference_ms"] = min(times[1:])
        e["cpu_reference_threads"] = threads
        e["speedup_vs_cpu_reference"] = e["cpu_reference_ms"] / e["ms"]
    if "filter_chain_8192" in out:
        size = 8192
        src = ref.lib.ref_image_create_rgba(size, size, abi.PACK_RGBA)
        ref.lib.ref_filter_map(src, abi.MAP_XOR_PATTERN, None, -1, 0, 0)
        mapped = ref.lib.ref_image_create_rgba(size, size, abi.PACK_RGBA)
        params = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
        t0 = time.perf_counter()
        ref.lib.ref_filter_map(mapped, abi.MAP_AFFINE, refbind.ptr(params), src, 0, 0)
        t1 = time.perf_counter()
        ref.lib.ref_filter_resize(mapped, abi.SAMPLER_LINEAR, size // 2, size // 2)
        t2 = time.perf_counter()
        ref.free_all()
        e = out["filter_chain_8192"]
        e["cpu_reference_map_ms"], e["cpu_reference_resize_down_ms"] = 1000.0 * (t1 - t0), 1000.0 * (t2 - t1)
        e["cpu_reference_threads"] = 1
        e["speedup_vs_cpu_reference"] = (e["cpu_reference_map_ms"] + e["cpu_reference_resize_down_ms"]) / e["chain_ms"]


def run(cuda, lib, cpu=True):
    import torch
    import sandbox_scene
    from dfpsr_b200 import abi, scenes
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    out = {}
    IM = lib.image
    s = lib.stream_ptr()
    ident = abi.Transform3D.identity()

    # ---- config 1: one 1080p terrain frame (latency of a single frame, clear fused)
    sc = scenes.terrain_scene()
    tex = lib.DeviceTexture(sc["texture"], 5)
    model = lib.DeviceModel(sc["points"], sc["polygons"], abi.FILTER_SOLID, tex)
    color = torch.empty((1080, 1920), dtype=torch.int32, device="cuda")
    depth = torch.empty((1080, 1920), dtype=torch.float32, device="cuda")
    cams = (abi.Camera * 1)(lib.camera(scenes.orbit_camera(7, 1920, 1080)))
    ci, di = (abi.Image * 1)(IM(color)), (abi.Image * 1)(IM(depth))
    ms = _time(torch, lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(model.desc), C.byref(ident), ci, di, cams, 1, 1, s)), iters=50)
    out["terrain_1080p_single_frame"] = {"ms": ms, "fps": 1000.0 / ms, "mpix_per_s": 1920 * 1080 / ms / 1e3}

    # ---- config 3: 2 M tiny vertex-coloured triangles at 3840x2160
    nx, nz = 1000, 999
    tiny = scenes.tiny_triangle_scene(nx, nz)
    tmodel = lib.DeviceModel(tiny["points"], tiny["polygons"])
    color4k = torch.empty((2160, 3840), dtype=torch.int32, device="cuda")
    depth4k = torch.empty((2160, 3840), dtype=torch.float32, device="cuda")
    cams4k = (abi.Camera * 1)(lib.camera(scenes.top_down_camera(nx, nz, 3840, 2160)))
    c4, d4 = (abi.Image * 1)(IM(color4k)), (abi.Image * 1)(IM(depth4k))
    ms = _time(torch, lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(tmodel.desc), C.byref(ident), c4, d4, cams4k, 1, 1, s)), iters=10)
    submitted = 2 * nx * nz
    algorithmic = 12 * len(tiny["points"]) + 24 * submitted + 16 * submitted + 8 * 3840 * 2160  # SURVEY §8d config 3 ≈ 158 MB
    out["tiny_triangles_4k"] = {"ms": ms, "fps": 1000.0 / ms, "mtri_per_s": submitted / ms / 1e3, "submitted_triangles": submitted,
                                "algorithmic_gb_s": algorithmic / ms / 1e6, "frac_of_hbm_peak": algorithmic / ms / 1e6 / peak}

    # ---- config 2: Sandbox 800x600, 1 directed + 16 shadow-casting point lights
    sb = sandbox_scene.build(800, 600, lights=16, seed=5)
    gpu = sandbox_scene.CudaSandbox(cuda, sb)
    gpu.composite()
    ms_light_unbatched = _time(torch, lambda: gpu.light(), iters=3)
    ms_light_batched = _time(torch, lambda: gpu.light_batched(), iters=10)
    ms_light = _time(torch, lambda: gpu.light_fused(), iters=20)
    ms_comp = _time(torch, lambda: gpu.composite(), iters=5)
    px = 800 * 600
    algorithmic = 20 * px + 16 * 2 * 256 * 1536 * 4  # SURVEY §8d config 2 ≈ 60 MB
    out["sandbox_800x600_16_lights"] = {"ms_light_passes": ms_light, "ms_light_passes_one_cube_map_per_light_call": ms_light_unbatched, "ms_light_passes_batched_shadows_separate_light_kernels": ms_light_batched, "ms_compositing_40_sprites": ms_comp, "fps_light_passes": 1000.0 / ms_light,
                                        "algorithmic_gb_s": algorithmic / ms_light / 1e6, "frac_of_hbm_peak": algorithmic / ms_light / 1e6 / peak}
    blend_ms = _time(torch, lambda: lib.check(cuda.dfpsr_light_blend(C.byref(IM(gpu.C)), C.byref(IM(gpu.D)), C.byref(IM(gpu.L)), s)), iters=50)
    out["sandbox_800x600_16_lights"]["blend_us"] = 1000.0 * blend_ms

    # ---- config 2 through the sprite world API (spriteWorld_draw): host planner + batched device execution, colour image on the device
    try:
        import sprite_world_scene as sws
        assets = sws.build_assets()
        script = sws.sandbox_script(800, 600, lights=16, frames=12)
        pw = sws.ProductWorld(cuda, lib.check, assets, shadow_res=256)
        target = torch.zeros((600, 800), dtype=torch.int32, device="cuda")
        frame_ms, launches = [], []
        for action in script:
            if action[0] != "draw":
                pw.apply(action)
                continue
            torch.cuda.synchronize()
            cuda.dfpsr_reset_launch_count()
            t0 = time.perf_counter()
            lib.check(cuda.dfpsr_sprite_world_draw(pw.world, C.byref(IM(target)), s))
            torch.cuda.synchronize()
            frame_ms.append(1000.0 * (time.perf_counter() - t0))
            launches.append(int(cuda.dfpsr_launch_count()))
        # The warmed-up world again without waiting for each frame (camera at rest, so no background block is generated inside the timed
        # loop): spriteWorld_draw only queues work plus one wait for the set-up totals, so the host side of frame k + 1 (scene updates, planning,
        # shadow batch) can overlap the device side of frame k; one synchronisation at the end. Wall clock of the whole loop, scene updates
        # through ctypes included.
        frame_actions = [a for a in script[next(i for i, a in enumerate(script) if a[0] == "clear_temporary"):] if a[0] != "move_camera"]
        torch.cuda.synchronize()
        t0, draws = time.perf_counter(), 0
        for action in frame_actions:
            if action[0] != "draw":
                pw.apply(action)
                continue
            lib.check(cuda.dfpsr_sprite_world_draw(pw.world, C.byref(IM(target)), s))
            draws += 1
        torch.cuda.synchronize()
        pipelined_ms = 1000.0 * (time.perf_counter() - t0) / draws
        pw.close()
        steady = sorted(frame_ms[2:])
        entry = {"ms_per_frame_pipelined": pipelined_ms, "fps_pipelined": 1000.0 / pipelined_ms, "ms_per_frame_median": steady[len(steady) // 2], "ms_first_frame": frame_ms[0], "fps": 1000.0 / steady[len(steady) // 2], "kernel_launches_per_frame": launches[-1],
                 "scene": "625 floor tiles + ~220 objects + 6 dense models, 1 directed + 16 shadow-casting point lights (256^2 x 6 cube maps), 2 temporary sprites, camera pans every other frame; wall clock per spriteWorld_draw incl. host planning"}
        try:  # the unmodified reference on the host cores, same session (oracle/_ref is test infrastructure: reported baseline only)
            import refbind
            import tempfile
            if refbind.available("sse"):
                ref = refbind.Ref("sse")
                timing = []
                sws.run_reference(ref, assets, script, tempfile.mkdtemp(prefix="dfpsr_bench_"), shadow_res=256, keep_frames=False, timing=timing)
                steady_ref = sorted(timing[2:])
                entry["cpu_reference_ms_per_frame_median"] = 1000.0 * steady_ref[len(steady_ref) // 2]
                entry["cpu_reference_threads"] = int(ref.lib.ref_thread_count())
                entry["speedup_vs_cpu_reference"] = entry["cpu_reference_ms_per_frame_median"] / entry["ms_per_frame_median"]
                entry["pipelined_speedup_vs_cpu_reference"] = entry["cpu_reference_ms_per_frame_median"] / entry["ms_per_frame_pipelined"]
        except Exception as exc:
            entry["cpu_reference_error"] = repr(exc)
        out["sandbox_800x600_sprite_world"] = entry
    except Exception as exc:
        out["sandbox_800x600_sprite_world"] = {"error": repr(exc)}

    # ---- the same per-pixel light passes at a size where launch latency does not hide their bandwidth (8192 x 8192)
    big = 8192
    bigN = torch.randint(0, 2 ** 31 - 1, (big, big), dtype=torch.int32, device="cuda")
    bigD = torch.randint(0, 2 ** 31 - 1, (big, big), dtype=torch.int32, device="cuda")
    bigL, bigC = torch.zeros_like(bigN), torch.zeros_like(bigN)
    d = sb["directed"]
    ms_dir = _time(torch, lambda: lib.check(cuda.dfpsr_light_directed(C.byref(gpu.view), C.byref(IM(bigL)), C.byref(IM(bigN)), d["direction"].ctypes.data, d["intensity"], d["color"].ctypes.data, 0, s)), iters=20)
    ms_dir_add = _time(torch, lambda: lib.check(cuda.dfpsr_light_directed(C.byref(gpu.view), C.byref(IM(bigL)), C.byref(IM(bigN)), d["direction"].ctypes.data, d["intensity"], d["color"].ctypes.data, 1, s)), iters=20)
    ms_blend = _time(torch, lambda: lib.check(cuda.dfpsr_light_blend(C.byref(IM(bigC)), C.byref(IM(bigD)), C.byref(IM(bigL)), s)), iters=20)
    px4k = big * big
    out["light_passes_8192x8192"] = {
        "set_directed_ms": ms_dir, "set_directed_gb_s": 8 * px4k / ms_dir / 1e6, "set_directed_frac_of_hbm_peak": 8 * px4k / ms_dir / 1e6 / peak,
        "add_directed_ms": ms_dir_add, "add_directed_gb_s": 12 * px4k / ms_dir_add / 1e6, "add_directed_frac_of_hbm_peak": 12 * px4k / ms_dir_add / 1e6 / peak,
        "blend_ms": ms_blend, "blend_gb_s": 12 * px4k / ms_blend / 1e6, "blend_frac_of_hbm_peak": 12 * px4k / ms_blend / 1e6 / peak,
        "note": "256 MB per buffer, larger than the 126 MB L2; bytes = algorithmic 8 / 12 / 12 B per pixel"}
    del bigN, bigD, bigL, bigC

    # ---- image_fill / draw_copy / draw_higher at a size where their bandwidth shows (rows a18-a20): 8192 x 8192 buffers of 268 MB
    big = 8192
    fa = torch.randint(0, 2 ** 31 - 1, (big, big), dtype=torch.int32, device="cuda")
    fb = torch.empty_like(fa)
    ha = torch.rand((big, big), dtype=torch.float32, device="cuda")
    hb = torch.rand((big, big), dtype=torch.float32, device="cuda")
    ms_fill = _time(torch, lambda: lib.check(cuda.dfpsr_image_fill_rgba(C.byref(IM(fb)), 10, 20, 30, 40, s)), iters=20)
    ms_copy = _time(torch, lambda: lib.check(cuda.dfpsr_draw_copy_rgba(C.byref(IM(fb)), C.byref(IM(fa)), 0, 0, s)), iters=20)
    ms_higher = _time(torch, lambda: lib.check(cuda.dfpsr_draw_higher(C.byref(IM(hb)), C.byref(IM(ha)), C.byref(IM(fb)), C.byref(IM(fa)), None, None, 0, 0, 0.0, s)), iters=10)
    # the first application onto an empty target: every pixel passes the height test, so both heights and the source colour are read
    # (12 B per pixel) and height + colour are written (8 B per pixel); the target is reset outside of the timed calls
    first_ms = 0.0
    for _ in range(5):
        lib.check(cuda.dfpsr_image_fill_f32(C.byref(IM(hb)), -1.0e30, s))
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        lib.check(cuda.dfpsr_draw_higher(C.byref(IM(hb)), C.byref(IM(ha)), C.byref(IM(fb)), C.byref(IM(fa)), None, None, 0, 0, 0.0, s))
        stop.record()
        torch.cuda.synchronize()
        first_ms += start.elapsed_time(stop) / 5
    pxb = big * big
    out["fill_copy_higher_8192x8192"] = {
        "fill_ms": ms_fill, "fill_gb_s": 4 * pxb / ms_fill / 1e6, "fill_frac_of_hbm_peak": 4 * pxb / ms_fill / 1e6 / peak,
        "copy_ms": ms_copy, "copy_gb_s": 8 * pxb / ms_copy / 1e6, "copy_frac_of_hbm_peak": 8 * pxb / ms_copy / 1e6 / peak,
        "higher_ms": ms_higher, "higher_note": "height + one colour image; steady state of repeated calls: the target already holds the higher value everywhere, so both height fields are read (8 B per pixel) and nothing is written",
        "higher_gb_s": 8 * pxb / ms_higher / 1e6, "higher_frac_of_hbm_peak": 8 * pxb / ms_higher / 1e6 / peak,
        "higher_first_application_ms": first_ms, "higher_first_application_gb_s": 20 * pxb / first_ms / 1e6, "higher_first_application_frac_of_hbm_peak": 20 * pxb / first_ms / 1e6 / peak,
        "higher_first_application_note": "empty target (heights -1e30): every pixel is replaced; 12 B read + 8 B written per pixel"}
    del fa, fb, ha, hb

    # ---- config 5: 8192x8192 filter chain (map + bilinear resize), pure streaming
    size = 8192
    src = torch.empty((size, size), dtype=torch.int32, device="cuda")
    mapped = torch.empty_like(src)
    half = torch.empty((size // 2, size // 2), dtype=torch.int32, device="cuda")
    up = torch.empty((size, size), dtype=torch.int32, device="cuda")
    scratch = torch.empty(size * (size // 2), dtype=torch.int32, device="cuda")
    lib.check(cuda.dfpsr_filter_map(C.byref(IM(src)), abi.MAP_XOR_PATTERN, None, 0, None, 0, 0, s))
    prm = np.array(sandbox_scene.CHAIN_AFFINE, np.int32)
    ms_map = _time(torch, lambda: lib.check(cuda.dfpsr_filter_map(C.byref(IM(mapped)), abi.MAP_AFFINE, prm.ctypes.data, 8, C.byref(IM(src)), 0, 0, s)))
    ms_down = _time(torch, lambda: lib.check(cuda.dfpsr_filter_resize(C.byref(IM(half)), C.byref(IM(mapped)), abi.SAMPLER_LINEAR, 0, None, s)))
    ms_up = _time(torch, lambda: lib.check(cuda.dfpsr_filter_resize(C.byref(IM(up)), C.byref(IM(half)), abi.SAMPLER_LINEAR, 0, scratch.data_ptr(), s)))
    odd = torch.empty((3000, 5000), dtype=torch.int32, device="cuda")  # the reference's generic (16-bit weight) path: neither dimension is kept or halved
    odd_scratch_bytes = int(cuda.dfpsr_filter_resize_scratch_bytes(size, size, 5000, 3000))
    odd_scratch = torch.empty(max(odd_scratch_bytes // 4, 1), dtype=torch.int32, device="cuda")
    ms_odd = _time(torch, lambda: lib.check(cuda.dfpsr_filter_resize(C.byref(IM(odd)), C.byref(IM(mapped)), abi.SAMPLER_LINEAR, 0, odd_scratch.data_ptr() if odd_scratch_bytes else None, s)))
    odd_bytes = 4 * (size * size + 5000 * 3000)
    map_bytes, down_bytes = 8 * size * size, 4 * (size * size + (size // 2) ** 2)
    out["filter_chain_8192"] = {
        "map_ms": ms_map, "map_gb_s": map_bytes / ms_map / 1e6, "map_frac_of_hbm_peak": map_bytes / ms_map / 1e6 / peak,
        "resize_down_ms": ms_down, "resize_down_gb_s": down_bytes / ms_down / 1e6, "resize_down_frac_of_hbm_peak": down_bytes / ms_down / 1e6 / peak,
        "resize_up_ms": ms_up, "resize_up_gb_s": down_bytes / ms_up / 1e6,
        "resize_5000x3000_ms": ms_odd, "resize_5000x3000_gb_s": odd_bytes / ms_odd / 1e6, "resize_5000x3000_frac_of_hbm_peak": odd_bytes / ms_odd / 1e6 / peak,
        "chain_ms": ms_map + ms_down, "chain_gb_s": (map_bytes + down_bytes) / (ms_map + ms_down) / 1e6,
        "chain_frac_of_hbm_peak": (map_bytes + down_bytes) / (ms_map + ms_down) / 1e6 / peak,
    }
    del src, mapped, half, up, scratch, odd, odd_scratch
    if cpu:
        try:
            cpu_reference(out)
        except Exception as exc:
            out["cpu_reference_error"] = repr(exc)
    return out
