"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` for the plumbing.

* large image  -> contiguous block-row shards; blocks are independent (ImageEncoder.cpp:121-126), so the only exchange
                  is ONE u64 per rank (the shard's bit total): all-gather -> exclusive scan -> every rank knows where
                  its bits start in the single output stream.  Each rank then re-aligns its shard to the 128-bit chunk
                  grid of the global stream (`ie_stream_shift_dev`), after which shards concatenate by plain byte copies
                  with an OR at the one chunk two neighbours share.
* image batch  -> whole images per rank, no communication.
* video        -> whole GOPs per rank (I-frame at idx % gop == 0, VideoBase.hpp:32; P-frames only reference their in-GOP
                  predecessor), stitched bit-contiguously like image shards.

The offset logic is plain integer arithmetic and is tested on CPU with gloo (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass

CHUNK_BITS = 128
CHUNK_BYTES = 16


def shard_block_rows(height: int, block: int, world: int, rank: int) -> tuple[int, int]:
    """Pixel-row range [y0, y1) owned by `rank`: contiguous block rows, remainder spread over the first ranks."""
    rows = height // block
    base, extra = divmod(rows, world)
    r0 = rank * base + min(rank, extra)
    r1 = r0 + base + (1 if rank < extra else 0)
    return r0 * block, r1 * block


def shard_items(count: int, world: int, rank: int) -> tuple[int, int]:
    """Index range [i0, i1) of whole items (images of a batch, GOPs of a video) owned by `rank`."""
    base, extra = divmod(count, world)
    i0 = rank * base + min(rank, extra)
    return i0, i0 + base + (1 if rank < extra else 0)


def shard_gops(frames: int, gop: int, world: int, rank: int) -> tuple[int, int]:
    """Frame range [f0, f1) owned by `rank`: whole GOPs (20 GOPs over 8 ranks -> 3,3,3,3,2,2,2,2; ideal speed-up 6.67x)."""
    gop = max(1, gop)
    ngops = (frames + gop - 1) // gop
    g0, g1 = shard_items(ngops, world, rank)
    return min(frames, g0 * gop), min(frames, g1 * gop)


@dataclass
class ShardPlacement:
    """Where a shard's bits land in the single output stream."""
    global_bit: int        # first bit of the shard in the global stream
    nbits: int
    shift: int             # global_bit % 128: leading zero bits of the re-aligned shard buffer
    byte_offset: int       # byte offset of the re-aligned shard buffer in the global stream (multiple of 16)
    nbytes: int            # bytes of the re-aligned shard buffer that carry data (multiple of 16)
    shares_first_chunk: bool   # first chunk also holds the tail of the previous shard -> OR-merge instead of copy


def place_shards(bit_totals: list[int], prefix_bits: int = 0) -> list[ShardPlacement]:
    """Exclusive scan of the per-rank bit totals.  `prefix_bits`: bits in front of rank 0's first block that are NOT in
    its total (0 when rank 0's total already includes the header it wrote)."""
    out = []
    pos = prefix_bits
    for n in bit_totals:
        shift = pos % CHUNK_BITS
        nbytes = (n + shift + CHUNK_BITS - 1) // CHUNK_BITS * CHUNK_BYTES if n else 0
        out.append(ShardPlacement(pos, n, shift, pos // CHUNK_BITS * CHUNK_BYTES, nbytes, shift != 0 and pos != 0))
        pos += n
    return out


def total_bytes(placements: list[ShardPlacement]) -> int:
    if not placements:
        return 0
    last = placements[-1]
    return (last.global_bit + last.nbits + 7) // 8


def merge_shard_into(stream: bytearray, shard: bytes, pl: ShardPlacement) -> None:
    """Host-side merge of one re-aligned shard buffer into the global stream (the writer side of the stitch)."""
    if pl.nbytes == 0:
        return
    end = pl.byte_offset + pl.nbytes
    if len(stream) < end:
        stream.extend(b"\0" * (end - len(stream)))
    start = 0
    if pl.shares_first_chunk:
        for i in range(CHUNK_BYTES):
            stream[pl.byte_offset + i] |= shard[i]
        start = CHUNK_BYTES
    stream[pl.byte_offset + start:end] = shard[start:pl.nbytes]


# ------------------------------------------------------------------------------------------------------------------
# device side (needs torch + NCCL or gloo)
# ------------------------------------------------------------------------------------------------------------------
def exchange_bit_totals(local_bits, group=None):
    """All-gather of one int64 per rank; returns (totals tensor [world], exclusive offsets tensor [world]) on the
    device of `local_bits`.  This is the only collective of a sharded encode."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        totals = local_bits.reshape(1).clone()
    else:
        totals = torch.empty(world, dtype=local_bits.dtype, device=local_bits.device)
        dist.all_gather_into_tensor(totals, local_bits.reshape(1), group=group)
    offsets = torch.cumsum(totals, 0) - totals
    return totals, offsets


class ShardedImageEncoder:
    """Encodes this rank's block-row shard of one large image straight into its place in the single output stream.

    tile kernel (blocks -> packed tile images in scratch, shard bit total) -> all-gather of one u64 per rank -> copy-out
    kernel that writes the shard at its global bit offset.  The collective sits between the two kernels; there is no
    re-alignment pass over the output."""

    def __init__(self, width: int, shard_height: int, block: int, full_height: int | None = None):
        import torch

        from . import device
        from ._lib import lib

        self.width, self.shard_height, self.block = width, shard_height, block
        self.sess = device.Session(device.Session.IMAGE_ENCODE, width, shard_height, block)
        if full_height:
            from ._lib import check
            check(lib().ie_session_set_header_height(self.sess.h, full_height))
        self.cap = int(lib().ie_max_encoded_bytes(width, shard_height, block, 1))
        self.d_aligned = torch.empty(self.cap + 32, dtype=torch.uint8, device="cuda")
        self.d_total = torch.zeros(1, dtype=torch.int64, device="cuda")     # this shard's bits (header included on rank 0)
        self.d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")      # (first % 128) + this shard's bits
        self.d_first = torch.zeros(1, dtype=torch.int64, device="cuda")     # first bit of this shard in the global stream
        self.d_totals = None

    def encode(self, d_raw, quant, rle: bool, rank: int, lead_bit: bool = True, group=None):
        """Asynchronous.  Returns the device tensor of all shards' bit totals; self.d_aligned then holds this rank's bytes
        of the global stream starting at byte (self.d_first // 128) * 16."""
        import torch
        import torch.distributed as dist

        from . import device

        device.encode_image_begin_dev(self.sess, d_raw, quant, rle, self.d_total, lead_bit=lead_bit, write_header=(rank == 0),
                                      width=self.width, height=self.shard_height)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            totals = self.d_total
        else:
            if self.d_totals is None or self.d_totals.numel() != world:
                self.d_totals = torch.empty(world, dtype=torch.int64, device="cuda")
            totals = self.d_totals
            dist.all_gather_into_tensor(totals, self.d_total, group=group)
        device.encode_image_end_dev(self.sess, totals, rank, self.d_aligned, self.d_bits, self.d_first)
        return totals
