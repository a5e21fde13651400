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


def cyclic_block(probe_rows: int, world: int, max_groups: int = 8) -> int:
    """Block size for block-cyclic ownership: 1 row unless that needs more than `max_groups`
    collectives per plane (one per group of world*block rows)."""
    block = 1
    while probe_rows > max_groups * world * block:
        block *= 2
    return block


def probe_row_blocks(probe_rows: int, rank: int, world: int, block: int = 1) -> list[tuple[int, int]]:
    """Block-cyclic ownership (ddgi_set_probe_rows_cyclic): the row ranges [y0, y1) with
    (y // block) % world == rank, in increasing order.  Spreads expensive regions of the
    field over all ranks, which contiguous slabs do not."""
    if world < 1 or not 0 <= rank < world or block < 1:
        raise ValueError(f"rank {rank} of world {world}, block {block}")
    out = []
    y = rank * block
    while y < probe_rows:
        out.append((y, min(y + block, probe_rows)))
        y += world * block
    return out


def probe_owner(probe: int, world: int, block: int = 1) -> int:
    """Rank that updates `probe` under probe-cyclic ownership (ddgi_set_probes_cyclic): probes dealt
    round-robin in blocks of `block` — the finest balance; paired with the fused exchange because a
    rank's texels are then scattered tiles, not a byte range."""
    if world < 1 or block < 1:
        raise ValueError(f"world {world}, block {block}")
    return (probe // block) % world


def frame_band_rows(screen_height: int, rank: int, world: int) -> tuple[int, int]:
    """Pixel rows [y0, y1) of the frame band `rank` renders (ddgi_set_frame_band): the reference
    dispatches floor(h/16) rows of 16x16 workgroups (src/rvpt/rvpt.cpp:1139-1140); they are split as
    evenly as integer arithmetic allows, in order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of world {world}")
    groups = screen_height // 16
    return 16 * (groups * rank // world), 16 * (groups * (rank + 1) // world)


def allgather_probe_rows_cyclic(plane, probe_rows: int, row_bytes: int, rank: int, world: int, block: int = 1, group=None):
    """In-place exchange of one texture plane under block-cyclic ownership: every group of
    world*block consecutive probe rows is one all_gather_into_tensor (rank r's block is the
    r-th chunk of the group); a ragged last group is broadcast block by block."""
    import torch.distributed as dist

    if world == 1:
        return
    span = world * block
    full_groups = probe_rows // span
    for g in range(full_groups):
        base = g * span * row_bytes
        mine = base + rank * block * row_bytes
        dist.all_gather_into_tensor(plane[base:base + span * row_bytes], plane[mine:mine + block * row_bytes], group=group)
    y = full_groups * span
    owner = 0
    while y < probe_rows:
        y1 = min(y + block, probe_rows)
        src = dist.get_global_rank(group, owner) if group is not None else owner
        dist.broadcast(plane[y * row_bytes:y1 * row_bytes], src=src, group=group)
        y = y1
        owner += 1


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
