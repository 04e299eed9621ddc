"""Work partitioning across the GPUs of one box (one process per GPU) — SURVEY §8e.

Three partitions, all without a data-path collective except the one exchange the path really has:
  * instances (skinning / morph / shadow streams): contiguous ranges, palettes travel with their instance;
  * offscreen views (config 4): view i -> rank i mod G, static geometry replicated once;
  * one large frame (config 5): contiguous strips of screen-tile rows (ogrecu_rt_set_strip) or, load-balanced,
    interleaved bands of tile rows (ogrecu_rt_set_bands: band b -> rank b mod G); every rank rasterises only
    its rows and the disjoint rows are GATHERED to one rank at readback — the only collective, <= 24.9 MB for a
    4K RGB8 frame, no blending.
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


def band_row_indices(height, rank, world, band_rows=TILE):
    """Pixel rows owned by `rank` under the interleaved split (ogrecu_rt_set_bands): bands of `band_rows` rows,
    band b -> rank b mod G.  Dense and empty screen regions spread over all ranks."""
    rows = np.arange(int(height))
    return rows[(rows // int(band_rows)) % int(world) == int(rank)]


def gather_bands(plane, height, band_rows=TILE, dst=0, group=None):
    """Composite of the interleaved split at readback: every rank passes its whole plane (torch tensor [height, width,
    channels], only its own bands valid); its rows are packed, gathered to `dst` (padded to a common count) and scattered
    back into place there.  Returns the full frame on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    idx = [band_row_indices(height, r, world, band_rows) for r in range(world)]
    rows = max(len(i) for i in idx)
    mine = torch.as_tensor(idx[rank], device=plane.device, dtype=torch.long)
    send = torch.zeros((rows,) + tuple(plane.shape[1:]), dtype=plane.dtype, device=plane.device)
    send[: len(mine)] = plane.index_select(0, mine)
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    frame = torch.empty_like(plane)
    for r in range(world):
        ir = torch.as_tensor(idx[r], device=plane.device, dtype=torch.long)
        frame.index_copy_(0, ir, recv[r][: len(ir)])
    return frame


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
