"""Work partitioning across the GPUs of one box (one process per GPU) — SURVEY §8e.

Three partitions, all without a data-path collective except the one exchange the path really has:
  * instances (skinning / morph / shadow streams): contiguous ranges, palettes travel with their instance;
  * offscreen views (config 4): view i -> rank i mod G, static geometry replicated once;
  * one large frame (config 5), sort-last: rank r draws the r-th slice of the triangles into a full-size target and the
    targets are merged per pixel by depth (composite_sort_last: one all-to-all of strips + ogrecu_composite_depth_dev) —
    nothing is replicated, so this is the form that scales; valid for opaque depth-tested content;
  * one large frame (config 5), sort-first: contiguous strips of screen-tile rows (ogrecu_rt_set_strip) or, load-balanced,
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


def triangle_range(num_triangles, rank, world):
    """Sort-last split: contiguous slice [lo, hi) of the triangle list (submission order) drawn by `rank`."""
    return instance_range(num_triangles, rank, world)


def composite_sort_last(ctx, colour, depth, height, clear_rgba, clear_depth, dst=0, group=None, want_depth=False, merge=None):
    """Sort-last composite (ogrecu_composite_depth_dev): every rank passes its full-size planes (torch CUDA tensors
    colour [H, W, 3] uint8, depth [H, W] float32) holding ITS slice of the triangles.  One all-to-all per plane sends
    rows [q R, (q+1) R) of every rank to rank q (R = ceil(H / G) pixel rows), which merges the G versions of its rows in
    rank order with one kernel (smallest depth, later rank on ties); the merged rows are then gathered to `dst`.
    Returns (colour, depth or None) of the whole frame on `dst`, None elsewhere.  The path's one real exchange step:
    (G-1)/G of 7 bytes per pixel per rank over NVLink.  When G divides H the planes are sent in place (no packing).
    `merge(recv_c, recv_z)` replaces the device merge in the CPU (gloo) tests of the exchange layout: it must fold
    recv_*[1:] into recv_*[0] in place."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from .capi import check, lib
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    w = colour.shape[1]
    rows = (height + world - 1) // world
    if rows * world == height:
        send_c, send_z = colour.view(world, rows, w, 3), depth.view(world, rows, w)
    else:                                                  # pad the last owner's rows; padding is never read back
        send_c = torch.empty((world * rows, w, 3), dtype=torch.uint8, device=colour.device)
        send_z = torch.empty((world * rows, w), dtype=torch.float32, device=colour.device)
        send_c[:height] = colour; send_z[:height] = depth
        send_c, send_z = send_c.view(world, rows, w, 3), send_z.view(world, rows, w)
    recv_c, recv_z = torch.empty_like(send_c), torch.empty_like(send_z)
    dist.all_to_all_single(recv_c, send_c, group=group)
    dist.all_to_all_single(recv_z, send_z, group=group)
    n = rows * w
    if merge is not None:
        merge(recv_c, recv_z)
        return _gather_rows(recv_c[0], recv_z[0], rows, world, rank, height, w, dst, group, want_depth)
    clear = (C.c_float * 4)(*[float(v) for v in clear_rgba])
    same_stream = int(lib().ogrecu_get_stream(ctx.handle) or 0) == int(torch.cuda.current_stream().cuda_stream)
    if not same_stream:
        torch.cuda.current_stream().synchronize()       # the merge runs on the context's stream
    if world > 1:                                       # recv[0] (rank 0's version) accumulates ranks 1 .. G-1, in order, in one launch
        check(lib().ogrecu_composite_depth_dev(ctx.handle, recv_c[0].data_ptr(), recv_z[0].data_ptr(), recv_c[1].data_ptr(),
                                               recv_z[1].data_ptr(), n, world - 1, n, clear, float(clear_depth)))
    if not same_stream:
        ctx.sync()

    return _gather_rows(recv_c[0], recv_z[0], rows, world, rank, height, w, dst, group, want_depth)


def _gather_rows(part_c, part_z, rows, world, rank, height, w, dst, group, want_depth):
    import torch
    import torch.distributed as dist

    def gather_rows(part, shape, dtype):
        full = torch.empty((world * rows,) + shape, dtype=dtype, device=part.device) if rank == dst else None
        dist.gather(part.contiguous(), [full[r * rows:(r + 1) * rows] for r in range(world)] if rank == dst else None, dst=dst, group=group)
        return full[:height] if rank == dst else None
    frame_c = gather_rows(part_c, (w, 3), torch.uint8)
    frame_z = gather_rows(part_z, (w,), torch.float32) if want_depth else None
    if rank != dst:
        return None
    return frame_c, frame_z


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
