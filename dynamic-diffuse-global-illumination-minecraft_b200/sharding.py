"""Probe-row sharding of the probe field across the GPUs of one box.

The tile of probe p sits at (p mod X*Z, p div X*Z) (assets/shaders/probe_pass.comp:139-145),
so probe row y is texture rows [y*ry, (y+1)*ry): rank r's slab of probe rows is ONE
contiguous byte range of each row-major RGBA8 plane, and the per-frame exchange is an
in-place all-gather of that range.  The reference has no multi-GPU path to mirror.
"""
from __future__ import annotations


def probe_row_shard(probe_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slab [y0, y1) of probe rows owned by `rank`: rows split as evenly as
    possible, the first `probe_rows % world` ranks take one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of world {world}")
    base, extra = divmod(probe_rows, world)
    y0 = rank * base + min(rank, extra)
    return y0, y0 + base + (1 if rank < extra else 0)


def shard_byte_ranges(probe_rows: int, world: int, row_bytes: int) -> list[tuple[int, int]]:
    """Byte range of every rank's slab inside one texture plane; `row_bytes` = bytes of one
    PROBE row = W * 4 * ry."""
    out = []
    for g in range(world):
        a, b = probe_row_shard(probe_rows, g, world)
        out.append((a * row_bytes, b * row_bytes))
    return out


def allgather_probe_rows(plane, probe_rows: int, row_bytes: int, rank: int, world: int, group=None):
    """In-place all-gather of one texture plane (a flat uint8 torch tensor over the whole
    plane, on any device the process group supports).  Even splits use ONE
    all_gather_into_tensor whose send buffer aliases the rank's own slab; ragged splits
    (probe_rows % world != 0: slabs differ in size, which all_gather does not take on
    every backend) broadcast each slab from its owner."""
    import torch.distributed as dist

    if world == 1:
        return
    ranges = shard_byte_ranges(probe_rows, world, row_bytes)
    a, b = ranges[rank]
    if probe_rows % world == 0:
        dist.all_gather_into_tensor(plane, plane[a:b], group=group)
    else:
        for g, (x, y) in enumerate(ranges):
            if y > x:
                dist.broadcast(plane[x:y], src=dist.get_global_rank(group, g) if group is not None else g, group=group)
