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
