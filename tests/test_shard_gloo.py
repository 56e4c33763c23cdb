"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (view sharding, strip split, the one gather step)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dfpsr_b200 import shard


def test_view_range_partitions_every_view_once():
    for total in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 4, 8):
            seen = [v for r in range(world) for v in shard.view_range(r, world, total)]
            assert seen == list(range(total))
            sizes = [len(shard.view_range(r, world, total)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_strip_rows_follow_the_reference_split():
    """ref: implementation/render/renderCore.cpp:459-470 with clipBound.top() == 0: y2 = height * (j + 1) / jobs, even unless last."""
    for height in (1, 2, 7, 600, 1080, 2160, 1081):
        for world in (1, 2, 4, 8, 12):
            bounds = shard.strip_rows(height, world)
            assert bounds[0][0] == 0 and bounds[-1][1] == height
            for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
                assert a1 == b0 and a0 <= a1
            for y0, y1 in bounds[:-1]:
                assert y1 % 2 == 0
            y1 = 0
            for j, (s0, s1) in enumerate(bounds):
                y2 = (height * (j + 1)) // world
                if j < world - 1:
                    y2 = (y2 // 2) * 2
                assert (s0, s1) == (y1, y2)
                y1 = y2
    for y0, y1 in shard.strip_rows(1080, 8, align=4)[:-1]:
        assert y1 % 4 == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import orcbind
        from dfpsr_b200 import abi, scenes
        oracle = orcbind.load()
        w, h = 320, 182
        sc = scenes.terrain_scene()
        buf, tex = orcbind.build_texture(sc["texture"], 5)
        model, keep = orcbind.model_of(sc["points"], sc["polygons"], diffuse=tex)
        ident = abi.Transform3D.identity()

        def render(frame):
            c, d = np.zeros((h, w), np.uint32), np.zeros((h, w), np.float32)
            oracle.orc_model_render(C.byref(model), C.byref(ident), C.byref(orcbind.image_of(c)), C.byref(orcbind.image_of(d)), C.byref(orcbind.camera(scenes.orbit_camera(frame, w, h))))
            return c, d

        # ---- strips: every rank keeps only its rows of the frame, one all_gather rebuilds it
        full_c, full_d = render(5)
        bounds = shard.strip_rows(h, world, align=4)
        y0, y1 = bounds[rank]
        color = torch.zeros((h, w), dtype=torch.int32)
        depth = torch.zeros((h, w), dtype=torch.float32)
        color[y0:y1] = torch.from_numpy(full_c.view(np.int32))[y0:y1]
        depth[y0:y1] = torch.from_numpy(full_d)[y0:y1]
        shard.gather_strips(color, bounds)
        shard.gather_strips(depth, bounds)
        assert np.array_equal(color.numpy().view(np.uint32), full_c)
        assert np.array_equal(depth.numpy().view(np.uint32), full_d.view(np.uint32))

        # ---- independent views: disjoint view sets, no data-path collective; a checksum of checksums is the only exchange
        total = 5
        mine = shard.view_range(rank, world, total)
        sums = torch.zeros(total, dtype=torch.int64)
        for v in mine:
            c, _ = render(v)
            sums[v] = int(c.astype(np.uint64).sum() % (2 ** 62))
        dist.all_reduce(sums)
        expected = [int(render(v)[0].astype(np.uint64).sum() % (2 ** 62)) for v in range(total)] if rank == 0 else None
        if rank == 0:
            assert sums.tolist() == expected
        open(os.path.join(result_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_strip_gather_and_view_sharding_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


class HostPeerTransport:
    """Stand-in for shard.CudaPeerTransport on a machine without GPUs: 'device memory' is a file under a shared directory mapped by
    every process, a 'pointer' is (region << 32 | offset), signal / wait are synchronous stores and polls. It exercises the handle
    exchange and the frame hand-shake of shard.PeerStripFrame with real processes; the CUDA transport is covered on two GPUs by
    tests/test_gpu_multi.py."""

    def __init__(self, directory, rank):
        self.directory, self.rank, self.regions, self.count = directory, rank, {}, 0

    def _map(self, path, nbytes=None):
        self.count += 1
        mode = "r+" if os.path.exists(path) else "w+"
        self.regions[self.count] = np.memmap(path, dtype=np.uint8, mode=mode, shape=(nbytes,) if nbytes else None)
        return self.count << 32

    def alloc(self, nbytes):
        path = os.path.join(self.directory, f"r{self.rank}_{self.count}")
        ptr = self._map(path, nbytes)
        return ptr, path.encode().ljust(64, b"\0")[:64] if len(path) <= 64 else path.encode()

    def open(self, handle):
        return self._map(handle.rstrip(b"\0").decode())

    def close(self, ptr):
        self.regions.pop(ptr >> 32).flush()

    free = close

    def u32(self, ptr):
        region, offset = self.regions[ptr >> 32], ptr & 0xFFFFFFFF
        return region[offset:offset + 4].view(np.uint32)

    def signal(self, flag_ptrs, value):
        for p in flag_ptrs:
            self.u32(p)[0] = value
            self.regions[p >> 32].flush()

    def wait(self, flags_ptr, count, value, status_ptr):
        import time
        deadline = time.time() + 20.0
        while any(int(self.u32(flags_ptr + 4 * i)[0]) < value for i in range(count)):
            if time.time() > deadline:
                self.u32(status_ptr)[0] = 1
                return
            time.sleep(0.0005)

    def read_u32(self, ptr):
        return int(self.u32(ptr)[0])

    def rows(self, ptr, stride, y0, y1):
        region = self.regions[ptr >> 32]
        return region[y0 * stride:y1 * stride]


def _peer_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 22, 8
        t = HostPeerTransport(result_dir, rank)
        frame = shard.PeerStripFrame(t, h, w, rank, world, presenter=0, align=4)
        assert frame.bounds == shard.strip_rows(h, world, align=4) and frame.rows == frame.bounds[rank]
        y0, y1 = frame.rows
        for k in range(1, 6):
            frame.begin_frame(k)                                   # a rank may not overwrite frame k - 1 before the presenter has consumed it
            t.rows(frame.color_ptr, frame.stride, y0, y1)[:] = (10 * k + rank) & 255
            frame.end_frame(k)                                     # presenter returns from here only when every strip of frame k has arrived
            if frame.is_presenter:
                for r, (a, b) in enumerate(frame.bounds):
                    got = np.unique(t.rows(frame.color_ptr, frame.stride, a, b))
                    assert got.tolist() == [(10 * k + r) & 255], (k, r, got)
            frame.release_frame(k)
        assert not frame.timed_out()
        frame.close()
        open(os.path.join(result_dir, f"peer_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_strip_frame_handshake_world_size_2(tmp_path):
    """Handle exchange + done/consumed hand-shake of shard.PeerStripFrame with two processes (host-memory stand-in for NVLink peer memory)."""
    world = 2
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"peer_ok{r}") for r in range(world))
