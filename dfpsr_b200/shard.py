"""Multi-GPU partitioning of the rendering hot path: one process per GPU, torch.distributed for plumbing.

Two modes, both without any collective on the data path of the kernels themselves:
  * independent views (BASELINE config 4): rank r renders views view_range(r, world, total) — no communication at all;
  * screen strips of one frame: rank r rasterises rows strip_rows(height, world)[r] of a replicated scene (the same split
    as the reference's worker threads, ref: implementation/render/renderCore.cpp:459-470, which is pixel-neutral because
    interpolation restarts from the target origin on every row pair, ref: shader/fillerTemplates.h:329-337), followed by
    ONE exchange step: an all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests) of the strips.
"""
import torch
import torch.distributed as dist


def view_range(rank, world, total):
    """Contiguous block of views for `rank`; sizes differ by at most one."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def strip_rows(height, world, align=2):
    """[(y0, y1)] per rank. Boundaries are multiples of `align` (2 in the reference so that 2x2 quads never straddle
    two workers; the CUDA path uses its tile height) except the last, which ends at `height`."""
    bounds, y1 = [], 0
    for j in range(world):
        y2 = (height * (j + 1)) // world
        if j < world - 1:
            y2 = (y2 // align) * align
        y2 = max(y2, y1)
        bounds.append((y1, y2))
        y1 = y2
    return bounds


def gather_strips(frame, bounds, group=None):
    """frame: (height, width) tensor whose rows bounds[rank] are valid on this rank. After the call every rank holds the
    whole frame. One all_gather of equally sized (padded) strips."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(bounds) == world
    rows = max(y1 - y0 for y0, y1 in bounds)
    if rows == 0:
        return frame
    y0, y1 = bounds[rank]
    mine = torch.zeros((rows,) + tuple(frame.shape[1:]), dtype=frame.dtype, device=frame.device)
    mine[: y1 - y0] = frame[y0:y1]
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    for (a, b), part in zip(bounds, parts):
        frame[a:b] = part[: b - a]
    return frame


# ---------------------------------------------------------------------------------------------- strips over NVLink peer memory

class CudaPeerTransport:
    """dfpsr_peer_* of the CUDA library (include/dfpsr_b200.h): IPC-exported device memory, flag stores and flag waits as kernels on
    the current stream. tests/test_shard_gloo.py drives PeerStripFrame with a host-memory stand-in of the same six calls."""

    def __init__(self, cuda, timeout_ms=2000):
        import ctypes as C
        from . import lib
        self.C, self.lib, self.cuda, self.timeout_ms = C, lib, cuda, timeout_ms
        self.status_ptr = None

    def alloc(self, nbytes):
        C = self.C
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        self.lib.check(self.cuda.dfpsr_peer_alloc(C.byref(ptr), nbytes, handle))
        return ptr.value, bytes(handle)

    def free(self, ptr):
        self.lib.check(self.cuda.dfpsr_peer_free(self.C.c_void_p(ptr)))

    def open(self, handle):
        C = self.C
        ptr = C.c_void_p()
        self.lib.check(self.cuda.dfpsr_peer_open(C.byref(ptr), (C.c_uint8 * 64).from_buffer_copy(handle)))
        return ptr.value

    def close(self, ptr):
        self.lib.check(self.cuda.dfpsr_peer_close(self.C.c_void_p(ptr)))

    def signal(self, flag_ptrs, value):
        C = self.C
        array = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self.lib.check(self.cuda.dfpsr_peer_signal(array, len(flag_ptrs), value, self.lib.stream_ptr()))

    def wait(self, flags_ptr, count, value, status_ptr):
        self.lib.check(self.cuda.dfpsr_peer_wait(self.C.c_void_p(flags_ptr), count, value, self.timeout_ms, self.C.c_void_p(status_ptr), self.lib.stream_ptr()))

    def reset_status(self, status_ptr):
        self.lib.check(self.cuda.dfpsr_peer_reset_status(self.C.c_void_p(status_ptr), self.lib.stream_ptr()))

    def read_u32(self, ptr):
        C = self.C
        out = C.c_uint32()
        self.lib.check(self.cuda.dfpsr_download(C.byref(out), C.c_void_p(ptr), 4, self.lib.stream_ptr()))
        self.lib.check(self.cuda.dfpsr_stream_synchronize(self.lib.stream_ptr()))
        return out.value


class PeerStripFrame:
    """One frame whose row strips are rendered by different ranks straight into the presenting rank's memory (screen strips;
    the reference's workers share one target the same way, ref: implementation/render/renderCore.cpp:449-480).

    The presenter exports frame + `done` flags (one u32 per rank); every rank exports a `consumed` flag of its own. Per frame k = 1, 2, ...:
        rank r:     begin_frame(k)   waits (on its stream) until the presenter has consumed frame k - 1
                    ... renders rows bounds[r] into `color_ptr` (dfpsr_renderer_begin_cleared + set_clip_rows + give_task + end) ...
                    end_frame(k)     signals done[r] = k in the presenter's memory; the presenter then waits for all done[*] >= k
        presenter:  ... consumes the frame on its stream ...
                    release_frame(k) signals consumed = k on every other rank
    Nothing in the loop touches the host or launches a collective; the data path is the tile kernel's own stores."""

    FLAG_BYTES = 256

    def __init__(self, transport, height, width, rank, world, presenter=0, align=4, bytes_per_pixel=4, group=None):
        assert 0 <= presenter < world and world <= 16
        self.t, self.rank, self.world, self.presenter, self.group = transport, rank, world, presenter, group
        self.height, self.width, self.stride = height, width, width * bytes_per_pixel
        self.bounds = strip_rows(height, world, align)
        self.is_presenter = rank == presenter
        self.mapped, self.owned = [], []
        mine = {}
        # own block: [0] consumed flag, [64] status of the begin_frame wait, [128] status of the end_frame wait (one block per wait site)
        self.local_ptr, mine["consumed"] = transport.alloc(self.FLAG_BYTES)
        self.owned.append(self.local_ptr)
        if self.is_presenter:
            self.color_ptr, mine["frame"] = transport.alloc(max(height * self.stride, 4))
            self.done_ptr, mine["done"] = transport.alloc(self.FLAG_BYTES)
            self.owned += [self.color_ptr, self.done_ptr]
        everyone = [None] * world
        if world > 1:
            dist.all_gather_object(everyone, mine, group=group)
        else:
            everyone[0] = mine
        self.consumed_ptrs = []
        if self.is_presenter:
            for r in range(world):
                if r != rank:
                    p = transport.open(everyone[r]["consumed"])
                    self.mapped.append(p)
                    self.consumed_ptrs.append(p)
        else:
            self.color_ptr = transport.open(everyone[presenter]["frame"])
            self.done_ptr = transport.open(everyone[presenter]["done"])
            self.mapped += [self.color_ptr, self.done_ptr]
        self.status_ptr = self.local_ptr + 64
        self.end_status_ptr = self.local_ptr + 128

    @property
    def rows(self):
        return self.bounds[self.rank]

    def begin_frame(self, k):
        if not self.is_presenter and k > 1:
            self.t.wait(self.local_ptr, 1, k - 1, self.status_ptr)

    def end_frame(self, k):
        self.t.signal([self.done_ptr + 4 * self.rank], k)
        if self.is_presenter:
            self.t.wait(self.done_ptr, self.world, k, self.end_status_ptr)

    def release_frame(self, k):
        if self.is_presenter and self.consumed_ptrs:
            self.t.signal(self.consumed_ptrs, k)

    def timed_out(self):
        """Number of this rank's waits that gave up since the last reset (a peer never signalled): the frames since then may be torn —
        the caller should stop presenting them and resynchronise (barrier + reset_status). Synchronises the stream."""
        return self.t.read_u32(self.status_ptr) + self.t.read_u32(self.end_status_ptr)

    def reset_status(self):
        if hasattr(self.t, "reset_status"):
            self.t.reset_status(self.status_ptr)
            self.t.reset_status(self.end_status_ptr)

    def close(self):
        if self.world > 1:
            dist.barrier(group=self.group)
        for p in self.mapped:
            self.t.close(p)
        if self.world > 1:
            dist.barrier(group=self.group)
        for p in self.owned:
            self.t.free(p)
        self.mapped, self.owned = [], []
