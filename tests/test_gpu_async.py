"""Asynchronous frames (dfpsr_renderer_set_async / dfpsr_set_default_async): renderer_end queues the whole frame without waiting for the
counts of its set-up pass; the pools are sized from earlier frames and a frame that does not fit is drawn again when the library next
looks at the renderer. Pixels must be identical to the synchronous path (and to the oracle) in every case."""
import ctypes as C
import time

import numpy as np
import pytest
import torch

from dfpsr_b200 import abi, lib, scenes
from gpuutil import CudaScene, assert_same_u32, bits, dev, host_f32, host_u32

pytestmark = pytest.mark.gpu


def make_renderer(cuda, asynchronous=True):
    r = C.c_void_p()
    lib.check(cuda.dfpsr_renderer_create(C.byref(r)))
    lib.check(cuda.dfpsr_renderer_set_async(r, 1 if asynchronous else 0))
    return r


def draw(cuda, r, scene, cam_params, tc, td, clear=True):
    cam = lib.camera(cam_params)
    ident = abi.Transform3D.identity()
    if clear:
        lib.check(cuda.dfpsr_renderer_begin_cleared(r, C.byref(lib.image(tc)), C.byref(lib.image(td)), 0, 0.0))
    else:
        lib.check(cuda.dfpsr_renderer_begin(r, C.byref(lib.image(tc)), C.byref(lib.image(td))))
    lib.check(cuda.dfpsr_renderer_give_task(r, C.byref(scene.model.desc), C.byref(ident), C.byref(cam), lib.stream_ptr()))
    lib.check(cuda.dfpsr_renderer_end(r, lib.stream_ptr()))


def test_async_frames_match_oracle(cuda, oracle):
    sc = scenes.terrain_scene()
    scene = CudaScene(sc["points"], sc["polygons"], diffuse_level0=sc["texture"], diffuse_levels=5)
    w, h = 640, 360
    r = make_renderer(cuda)
    zero_c, zero_d = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    for frame in (0, 9, 21, 33, 47):  # the first frame waits for its counts, the others do not
        tc, td = dev(zero_c), dev(zero_d)
        cam = scenes.orbit_camera(frame, w, h)
        draw(cuda, r, scene, cam, tc, td)
        lib.check(cuda.dfpsr_renderer_flush(r))  # the tensors are read by torch, not through this library
        exp_c, exp_d, _ = scene.render_oracle(oracle, cam, zero_c, zero_d)
        assert_same_u32(bits(host_f32(td)), bits(exp_d), f"depth of frame {frame}")
        assert_same_u32(host_u32(tc), exp_c, f"colour of frame {frame}")
    lib.check(cuda.dfpsr_renderer_destroy(r))


def far_and_near(w, h):
    far = abi.camera_params(True, scenes.look_at_transform((0, 0, -400), (0, 0, 0)), w, h)
    near = abi.camera_params(True, scenes.look_at_transform((0, 0, -5), (0, 0, 0)), w, h)
    return far, near


def test_frame_that_outgrows_the_pools_is_drawn_again(cuda, oracle):
    """A renderer whose history is a cloud of sub-pixel triangles far away meets the same triangles filling the screen: rows, tile entries
    and checkpoints outgrow the pools, the frame is dropped on the device and drawn again when the library next looks at the renderer."""
    big = scenes.random_soup(4000, 4, extent=3.0, tri_size=1.5, textured=False)
    scene = CudaScene(big["points"], big["polygons"])
    w, h = 320, 200
    far, near = far_and_near(w, h)
    zero_c, zero_d = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    r = make_renderer(cuda)
    for _ in range(2):  # the first frame waits for its counts and leaves the history, the second is asynchronous
        tc, td = dev(zero_c), dev(zero_d)
        draw(cuda, r, scene, far, tc, td)
    background = np.full((h, w), 0x11223344, np.uint32)
    tc, td = dev(background), dev(zero_d)
    draw(cuda, r, scene, near, tc, td, clear=False)
    count = C.c_int64()
    lib.check(cuda.dfpsr_renderer_last_command_count(r, C.byref(count), lib.stream_ptr()))  # looks at the renderer: verifies and redraws
    exp_c, exp_d, commands = scene.render_oracle(oracle, near, background, zero_d)
    assert commands > 1000 and count.value == commands
    assert_same_u32(bits(host_f32(td)), bits(exp_d), "depth")
    assert_same_u32(host_u32(tc), exp_c, "colour")
    # the next frame of the same size fits the grown pools and is not dropped
    tc2, td2 = dev(background), dev(zero_d)
    draw(cuda, r, scene, near, tc2, td2, clear=False)
    lib.check(cuda.dfpsr_renderer_flush(r))
    assert_same_u32(host_u32(tc2), exp_c, "colour of the following frame")
    lib.check(cuda.dfpsr_renderer_destroy(r))


def test_consumers_inside_the_library_see_the_redrawn_frame(cuda, oracle):
    """dfpsr_download (and every launch of this library) verifies frames in flight first: no explicit flush is needed."""
    big = scenes.random_soup(3000, 6, extent=3.0, tri_size=1.5, textured=False)
    scene = CudaScene(big["points"], big["polygons"])
    w, h = 256, 128
    far, near = far_and_near(w, h)
    zero_c, zero_d = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
    r = make_renderer(cuda)
    for _ in range(2):
        tc, td = dev(zero_c), dev(zero_d)
        draw(cuda, r, scene, far, tc, td)
    tc, td = dev(zero_c), dev(zero_d)
    draw(cuda, r, scene, near, tc, td)
    out = np.empty((h, w), np.uint32)
    lib.check(cuda.dfpsr_download(out.ctypes.data, tc.data_ptr(), out.nbytes, lib.stream_ptr()))
    lib.check(cuda.dfpsr_stream_synchronize(lib.stream_ptr()))
    exp_c, _, _ = scene.render_oracle(oracle, near, zero_c, zero_d)
    assert_same_u32(out, exp_c, "colour through dfpsr_download")
    lib.check(cuda.dfpsr_renderer_destroy(r))


def test_renderer_end_returns_before_the_frame_is_drawn(cuda):
    """The host leaves renderer_end while the device is still working on the frame (a batch of 1080p views takes milliseconds)."""
    sc = scenes.terrain_scene()
    texture = lib.DeviceTexture(sc["texture"], 5)
    model = lib.DeviceModel(sc["points"], sc["polygons"], abi.FILTER_SOLID, texture)
    views, w, h = 64, 1920, 1080
    color = torch.empty((views, h, w), dtype=torch.int32, device="cuda")
    depth = torch.empty((views, h, w), dtype=torch.float32, device="cuda")
    cams = (abi.Camera * views)(*[lib.camera(scenes.orbit_camera(v, w, h, frames_per_lap=views)) for v in range(views)])
    ci = (abi.Image * views)(*[lib.image(color[v]) for v in range(views)])
    di = (abi.Image * views)(*[lib.image(depth[v]) for v in range(views)])
    ident = abi.Transform3D.identity()
    lib.check(cuda.dfpsr_set_default_async(1))
    try:
        call = lambda: lib.check(cuda.dfpsr_model_render_views(C.byref(model.desc), C.byref(ident), ci, di, cams, views, 1, lib.stream_ptr()))
        for _ in range(2):  # the first call waits for its counts and leaves the history, the second grows the pools to their asynchronous head room
            call()
            lib.check(cuda.dfpsr_flush())  # torch reads the targets directly: the frame must have been verified (and redrawn if it outgrew the pools)
        torch.cuda.synchronize()
        reference = color[views - 1].clone()
        color.zero_()
        torch.cuda.synchronize()
        done = torch.cuda.Event()
        t0 = time.perf_counter()
        call()
        host_us = 1e6 * (time.perf_counter() - t0)
        done.record()
        returned_early = not done.query()
        done.synchronize()
        device_us = 1e6 * (time.perf_counter() - t0)
        lib.check(cuda.dfpsr_flush())
        print(f"renderer_end returned after {host_us:.0f} us, the batch was drawn after {device_us:.0f} us")
        assert returned_early, "the host waited for the frame inside renderer_end"
        assert host_us < 0.5 * device_us
        assert torch.equal(color[views - 1], reference)
    finally:
        lib.check(cuda.dfpsr_set_default_async(0))
