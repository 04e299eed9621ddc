"""Work partitioning across the GPUs of one box (one process per GPU) — SURVEY §8e.

Three partitions, all without a data-path collective except the one exchange the path really has:
  * instances (skinning / morph / shadow streams): contiguous ranges, palettes travel with their instance;
  * offscreen views (config 4): view i -> rank i mod G, static geometry replicated once;
  * one large frame (config 5): contiguous strips of screen-tile rows; every rank rasterises only its
    strip (ogrecu_rt_set_strip) and the disjoint strips are GATHERED to one rank at readback — the only
    collective, <= 24.9 MB for a 4K RGB8 frame, no blending.
torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests); nothing is computed.
"""
import numpy as np

TILE = 16   # raster tile edge (ogre_b200/csrc/raster.cu)


def instance_range(num_instances, rank, world):
    """Contiguous [lo, hi) of instances owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(num_instances), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def view_indices(num_views, rank, world):
    """Views rendered by `rank`: i mod G (interleaved so neighbouring cameras spread over GPUs)."""
    return list(range(int(rank), int(num_views), int(world)))


def strip_rows(height, rank, world, tile=TILE):
    """Pixel rows [y0, y1) of the screen strip owned by `rank`: whole tile rows, contiguous, balanced.
    y0 is a multiple of `tile`; y1 is a multiple of `tile` or `height` (ogrecu_rt_set_strip's contract).
    Ranks beyond the number of tile rows get the empty strip (0, 0)."""
    tile_rows = (int(height) + tile - 1) // tile
    lo, hi = instance_range(tile_rows, rank, world)
    if lo == hi:
        return 0, 0
    return lo * tile, min(hi * tile, height)


def gather_strips(strip, y0, y1, height, dst=0, group=None):
    """Composite at readback: every rank passes its rows [y0, y1) of the frame (`strip`, a torch tensor of
    shape [y1-y0, width, channels], on the GPU for NCCL or on the CPU for gloo); rank `dst` gets the full
    [height, width, channels] frame, the others None.  Strips are padded to a common row count because
    gather needs equal shapes; the padding is never copied into the frame."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = max(strip_rows(height, r, world)[1] - strip_rows(height, r, world)[0] for r in range(world))
    w, c = strip.shape[1], strip.shape[2]
    send = torch.zeros((rows, w, c), dtype=strip.dtype, device=strip.device)
    send[: y1 - y0] = strip
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    frame = torch.empty((height, w, c), dtype=strip.dtype, device=strip.device)
    for r in range(world):
        a, b = strip_rows(height, r, world)
        frame[a:b] = recv[r][: b - a]
    return frame


class DevicePlane:
    """Zero-copy torch view of a device plane owned by libogrecu (e.g. a render target's colour plane)."""

    def __init__(self, ptr, shape, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=device)
