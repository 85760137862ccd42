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


ABSENT = (1 << 64) - 1        # first-occurrence value of a byte that does not occur (ie_byte_histogram_dev)


def shard_byte_range(placements: list[ShardPlacement], rank: int) -> tuple[int, int]:
    """Bytes [b0, b1) of the global plain stream whose Huffman codes `rank` produces: a byte belongs to the shard that holds
    its first bit (the byte two shards share goes to the left one, which gets the right one's bits of it)."""
    b0 = 0 if rank == 0 else (placements[rank].global_bit + 7) // 8
    last = rank == len(placements) - 1
    end_bit = placements[rank].global_bit + placements[rank].nbits
    b1 = (end_bit + 7) // 8                     # the last shard ends the stream; otherwise: ceil(first bit of rank+1 / 8)
    return b0, max(b0, b1) if last or placements[rank].nbits else b0


def reduce_histograms(hists, firsts, byte_offsets):
    """What the all-reduce computes: sum of the shard histograms, min of the shard-local first occurrences moved to
    global byte positions (Huffman.cpp:236-243 feeds its containers in first-occurrence order)."""
    import numpy as np
    hist = np.zeros(256, dtype=np.uint64)
    first = np.full(256, ABSENT, dtype=np.uint64)
    for h, f, off in zip(hists, firsts, byte_offsets):
        hist += np.asarray(h, dtype=np.uint64)
        f = np.asarray(f, dtype=np.uint64)
        g = np.where(f == ABSENT, f, f + np.uint64(off))
        first = np.minimum(first, g)
    return hist.astype(np.uint32), first


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


class Comm:
    """ie_comm (csrc/comm.cu): the library's own multi-GPU exchange -- peer-mapped mailboxes written over NVLink, no NCCL on the
    data path.  `torch.distributed` is used ONCE, at set-up, to swap the CUDA IPC handles (any all-gather of bytes would do)."""

    def __init__(self, rank: int, world: int, stitch_bytes: int = 0, group=None):
        import ctypes as C

        import torch.distributed as dist

        from ._lib import check, lib
        self.rank, self.world = rank, world
        self.h = C.c_void_p()
        check(lib().ie_comm_create(C.byref(self.h), rank, world, stitch_bytes if rank == 0 else 0))
        n = int(lib().ie_comm_handle_bytes())
        blob = C.create_string_buffer(n)
        check(lib().ie_comm_export(self.h, blob))
        blobs = [bytes(blob.raw)]
        if world > 1:
            blobs = [None] * world
            dist.all_gather_object(blobs, bytes(blob.raw), group=group)
        allb = C.create_string_buffer(b"".join(blobs), n * world)
        check(lib().ie_comm_connect(self.h, allb))

    def close(self):
        from ._lib import lib
        if self.h:
            lib().ie_comm_destroy(self.h)
            self.h = None

    def encode_image_shard(self, sess, d_raw, width, height_shard, height_total, quant, rle, d_out, d_bits, d_first, lead_bit=True):
        """block-row shard of one image stream: tile kernel -> mailbox exchange of the shard totals -> copy-out at the global
        offset (ie_encode_image_shard_dev).  Asynchronous on the current stream."""
        from . import device
        from ._lib import check, lib
        _q, qp = device._q(quant)
        check(lib().ie_encode_image_shard_dev(sess.h, self.h, device._dp(d_raw), width, height_shard, height_total, qp, int(bool(rle)),
                                              int(bool(lead_bit)), device._dp(d_out), d_out.numel(), device._dp(d_bits), device._dp(d_first),
                                              device._stream()))

    def exchange(self, d_total):
        """all-gather of one int64 per rank through the mailboxes (ie_comm_exchange_totals_dev); returns a device tensor [world].
        Asynchronous on the current stream."""
        import torch

        from . import device
        from ._lib import check, lib
        out = torch.empty(self.world, dtype=torch.int64, device="cuda")
        check(lib().ie_comm_exchange_totals_dev(self.h, device._dp(d_total), device._dp(out), device._stream()))
        return out

    def totals(self):
        """the bit totals of the last exchange: torch int64 tensor [world] on the device (a copy, stream-ordered)"""
        import torch

        from . import device
        from ._lib import check, lib
        out = torch.empty(self.world, dtype=torch.int64, device="cuda")
        check(lib().ie_comm_copy_totals(self.h, device._dp(out), device._stream()))
        return out

    def stitch(self, d_shard, d_bits, d_first):
        """every rank: its chunks -> rank 0's stitch buffer over NVLink (ie_comm_stitch_dev).  Asynchronous."""
        from . import device
        from ._lib import check, lib
        check(lib().ie_comm_stitch_dev(self.h, device._dp(d_shard), device._dp(d_bits), device._dp(d_first), device._stream()))

    def stitched(self, nbytes: int):
        """rank 0: the first nbytes of the stitched stream as a device tensor (a copy)"""
        import torch

        from . import device
        from ._lib import check, lib
        out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        check(lib().ie_comm_copy_stitched(self.h, device._dp(out), nbytes, device._stream()))
        return out


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
        # more ranks than block rows leave some ranks without rows: they take part in the all-gather with 0 bits
        self.sess = device.Session(device.Session.IMAGE_ENCODE, width, shard_height, block) if shard_height else None
        if full_height and self.sess is not None:
            from ._lib import check
            check(lib().ie_session_set_header_height(self.sess.h, full_height))
        self.cap = int(lib().ie_max_encoded_bytes(width, max(shard_height, block), block, 1))
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

        if self.sess is None:
            if rank == 0:
                raise ValueError("rank 0 writes the header and needs at least one block row")
            self.d_total.zero_()
        else:
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
        if self.sess is None:
            self.d_bits.zero_()
            self.d_first.copy_(totals[:rank].sum().reshape(1))
        else:
            device.encode_image_end_dev(self.sess, totals, rank, self.d_aligned, self.d_bits, self.d_first)
        return totals


class ShardedHuffmanStage:
    """Huffman stage of a block-row sharded encode (BASELINE config 3): every rank codes the bytes of the plain stream it
    holds with the dictionary of the WHOLE stream.

    exchange 1 (all-gather, one byte per rank): the bits a rank has of the byte it shares with its left neighbour
    exchange 2 (all-reduce): 256-bin histogram (sum) and first-occurrence positions (min)
    exchange 3 (all-gather, one u64 per rank): code bits per shard -> offsets of the Huffman shards
    The tree is built on every rank from identical inputs (same libstdc++ containers as the reference, csrc/huffman.cu).
    """

    def __init__(self, enc: ShardedImageEncoder):
        import torch

        from . import device
        self.enc = enc
        self.sess = device.Session(device.Session.IMAGE_ENCODE, enc.width, enc.shard_height, enc.block)
        self.d_plain = torch.zeros(enc.cap + 64, dtype=torch.uint8, device="cuda")
        self.d_code = torch.zeros(enc.cap + (1 << 16), dtype=torch.uint8, device="cuda")
        self.d_out = torch.zeros(enc.cap + (1 << 16) + 32, dtype=torch.uint8, device="cuda")
        self.d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.d_params = torch.zeros(2, dtype=torch.int64, device="cuda")

    # -- step 1: what the left neighbour needs ----------------------------------------------------------------------
    def head_byte(self, pl: list[ShardPlacement], rank: int) -> int:
        p = pl[rank]
        if rank == 0 or p.global_bit % 8 == 0 or p.nbits == 0:
            return 0
        return int(self.enc.d_aligned[p.global_bit // 8 - p.byte_offset].item())

    # -- step 2: this rank's bytes, then their histogram --------------------------------------------------------------
    def local_histogram(self, pl: list[ShardPlacement], rank: int, heads: list[int]):
        from . import device
        p = pl[rank]
        self.b0, self.b1 = shard_byte_range(pl, rank)
        self.n = self.b1 - self.b0
        if self.n:
            # the last byte of this rank may be shared with the ranks behind it -- with several of them when their shards are
            # shorter than a byte (a one-block-wide image of all-zero blocks: 4 bits per block row): OR in the bits of every rank
            # that starts inside that byte
            last = self.b1 - 1
            r = rank + 1
            while r < len(pl) and pl[r].global_bit // 8 == last and pl[r].global_bit % 8:
                if pl[r].nbits:
                    self.enc.d_aligned[last - p.byte_offset] |= heads[r]
                r += 1
        if self.n == 0:
            import numpy as np
            return np.zeros(256, np.uint32), np.full(256, ABSENT, np.uint64)
        self.d_plain[: self.n].copy_(self.enc.d_aligned[self.b0 - p.byte_offset: self.b1 - p.byte_offset])
        return device.byte_histogram_dev(self.d_plain, self.n)

    # -- step 3: code this rank's bytes ---------------------------------------------------------------------------------
    def encode(self, hist, first, rank: int) -> int:
        import torch

        from . import device
        if self.n == 0:
            self.d_bits.zero_()
            return 0
        device.huffman_encode_shard_dev(self.sess, self.d_plain, self.n, hist, first, rank == 0, self.d_code, self.d_bits)
        torch.cuda.synchronize()
        return int(self.d_bits.item())

    # -- steps 1-4 on the device: nothing here reads a value back.  dev_head / dev_histogram / dev_encode / dev_place are the
    #    four steps between the three exchanges (tests with emulated ranks call them rank by rank); run_device strings them
    #    together with the caller's collectives.
    def dev_head(self, pl: list[ShardPlacement], rank: int):
        """int64 device tensor [1]: the bits this rank has of the byte it shares with its left neighbour(s)"""
        import torch
        p = pl[rank]
        self.b0, self.b1 = shard_byte_range(pl, rank)
        self.n = self.b1 - self.b0
        if rank > 0 and p.global_bit % 8 and p.nbits:
            i = p.global_bit // 8 - p.byte_offset
            return self.enc.d_aligned[i:i + 1].to(torch.int64)
        return torch.zeros(1, dtype=torch.int64, device="cuda")

    def dev_histogram(self, pl: list[ShardPlacement], rank: int, heads):
        """heads: int64 device tensor [world].  Returns (histogram, first occurrences in stream coordinates; 2^62 = absent),
        int64 device tensors [256]: what the sum / min all-reduce takes"""
        import torch

        from . import device
        world, p, n = len(pl), pl[rank], self.n
        BIG = 1 << 62
        if n:
            last = self.b1 - 1
            r = rank + 1
            while r < world and pl[r].global_bit // 8 == last and pl[r].global_bit % 8:
                if pl[r].nbits:
                    self.enc.d_aligned[last - p.byte_offset] |= heads[r].to(torch.uint8)
                r += 1
        if not hasattr(self, "d_hist"):
            self.d_hist = torch.zeros(256, dtype=torch.int32, device="cuda")
            self.d_first = torch.zeros(256, dtype=torch.int64, device="cuda")
            self.d_cb = torch.zeros(1, dtype=torch.int64, device="cuda")
            self.d_params2 = torch.zeros(2, dtype=torch.int64, device="cuda")
        if not n:
            return torch.zeros(256, dtype=torch.int64, device="cuda"), torch.full((256,), BIG, dtype=torch.int64, device="cuda")
        self.d_plain[:n].copy_(self.enc.d_aligned[self.b0 - p.byte_offset: self.b1 - p.byte_offset])
        device.byte_histogram_async_dev(self.sess, self.d_plain, n, self.d_hist, self.d_first)
        f = torch.where(self.d_first < 0, torch.full_like(self.d_first, BIG), self.d_first + self.b0)
        return self.d_hist.to(torch.int64), f

    def dev_encode(self, rank: int, hist_g, first_g):
        """codes this rank's bytes with the dictionary of the whole stream (built by a host callback in stream order); returns its
        code bit count, int64 device tensor [1]"""
        import torch

        from . import device
        if self.n:
            device.huffman_encode_shard_async_dev(self.sess, self.d_plain, self.n, hist_g.to(torch.int32).contiguous(), first_g.contiguous(),
                                                  rank == 0, self.d_code, self.d_cb)
        else:
            self.d_cb.zero_()
        return self.d_cb

    def dev_place(self, pl: list[ShardPlacement], rank: int, cb) -> None:
        """cb: the ranks' code bit counts, int64 device tensor [world].  Shifts this rank's shard to its place in self.d_out --
        or, by the revert rule over all shards (Huffman.cpp:329-341), its raw bytes behind the '0' bit"""
        import torch

        from . import device
        excl = torch.cumsum(cb, 0) - cb
        reverted = (cb.sum() + 7) // 8 > total_bytes(pl)
        zero2 = torch.zeros(2, dtype=torch.int64, device="cuda")
        par_code = torch.stack([cb[rank], excl[rank]])
        par_rev = torch.tensor([8 * self.n, 1 + 8 * self.b0], dtype=torch.int64, device="cuda")
        # both shifts are enqueued; the one that does not apply gets (0 bits at offset 0) and writes nothing
        self.d_params.copy_(torch.where(reverted, zero2, par_code))
        device.stream_shift_dev(self.d_code, self.d_params, self.d_out)
        self.d_params2.copy_(torch.where(reverted, par_rev, zero2))
        device.stream_shift_dev(self.d_plain, self.d_params2, self.d_out)

    def run_device(self, pl: list[ShardPlacement], rank: int, gather=None, reduce_=None):
        """`gather(out, t)`: all-gather of an int64 tensor [1] per rank into out [world]; `reduce_(t, "sum" | "min")`: in-place
        all-reduce of an int64 tensor [256]; both None for a single rank.  Returns the ranks' code bit counts (int64 device
        tensor [world]); the shifted shard is in self.d_out."""
        import torch
        world = len(pl)

        def all_gather(t):
            out = torch.empty(world, dtype=torch.int64, device="cuda")
            if gather:
                gather(out, t)
            else:
                out.copy_(t)
            return out

        heads = all_gather(self.dev_head(pl, rank))                         # exchange 1
        h, f = self.dev_histogram(pl, rank, heads)
        if reduce_:                                                         # exchange 2
            reduce_(h, "sum")
            reduce_(f, "min")
        cb = all_gather(self.dev_encode(rank, h, f))                        # exchange 3
        self.dev_place(pl, rank, cb)
        return cb

    # -- step 4: place the shard (or, by the reference's revert rule, the raw bytes behind a '0' bit) -----------------------
    def place(self, code_bits: list[int], plain_pl: list[ShardPlacement], rank: int) -> tuple[list[ShardPlacement], bool]:
        from . import device
        total_code_bytes = (sum(code_bits) + 7) // 8
        reverted = total_bytes(plain_pl) < total_code_bytes                        # Huffman.cpp:329-341: no gain
        if reverted:
            # '0' + the plain bytes: rank r's bytes start at bit 1 + 8 * b0(r); nothing more to exchange
            ranges = [shard_byte_range(plain_pl, r) for r in range(len(plain_pl))]
            bits = [8 * (b1 - b0) + (1 if r == 0 else 0) for r, (b0, b1) in enumerate(ranges)]
            pl = place_shards(bits)
            self.d_params[0] = 8 * self.n
            self.d_params[1] = 1 + 8 * self.b0
            device.stream_shift_dev(self.d_plain, self.d_params, self.d_out)
            return pl, True
        pl = place_shards(code_bits)
        self.d_params[0] = code_bits[rank]
        self.d_params[1] = pl[rank].global_bit
        device.stream_shift_dev(self.d_code, self.d_params, self.d_out)
        return pl, False


def sharded_image_encode_huffman_dev(enc: ShardedImageEncoder, stage: ShardedHuffmanStage, d_raw, quant, rle: bool, rank: int, group=None):
    """Block-row sharded encode + Huffman stage over `torch.distributed`, device-resident: the host reads the shards' plain bit
    totals once (every size below follows from them) and the code bit totals once at the end (the caller needs the placements);
    the three exchanges in between run on device tensors, the dictionary is built by a stream-ordered host callback
    (ie_huffman_encode_shard_async_dev).  Same bytes as sharded_image_encode_huffman.  Returns (placements of the Huffman
    shards, this rank's aligned shard bytes as a device tensor view)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    totals = enc.encode(d_raw, quant, rle, rank, lead_bit=False, group=group)
    pl = place_shards([int(b) for b in totals.cpu().tolist()])                  # synchronisation 1 of 2
    gather = (lambda out, t: dist.all_gather_into_tensor(out, t, group=group)) if world > 1 else None
    reduce_ = (lambda t, op: dist.all_reduce(t, op=(dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MIN), group=group)) if world > 1 else None
    cb = stage.run_device(pl, rank, gather, reduce_)
    code_bits = [int(x) for x in cb.cpu().tolist()]                             # synchronisation 2 of 2
    reverted = total_bytes(pl) < (sum(code_bits) + 7) // 8
    if reverted:
        ranges = [shard_byte_range(pl, r) for r in range(len(pl))]
        hpl = place_shards([8 * (b1 - b0) + (1 if r == 0 else 0) for r, (b0, b1) in enumerate(ranges)])
    else:
        hpl = place_shards(code_bits)
    return hpl, stage.d_out[: hpl[rank].nbytes]


def sharded_image_encode_huffman(enc: ShardedImageEncoder, stage: ShardedHuffmanStage, d_raw, quant, rle: bool, rank: int, group=None):
    """Block-row sharded encode + Huffman stage over `torch.distributed`, host-orchestrated (every exchange visits the host: eight
    synchronisations per image).  For ONE image at a time this is the faster of the two orchestrations (config 3: 1.55 / 1.17 ms
    at 2 / 8 GPUs against 1.63 / 1.36 ms of sharded_image_encode_huffman_dev, whose stream-ordered dictionary callback and small
    device operations cost more than the synchronisations they replace); the device-resident one never blocks the host between
    its two read-backs.  Returns (placements of the Huffman shards, this
    rank's aligned shard bytes as a device tensor view)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    totals = enc.encode(d_raw, quant, rle, rank, lead_bit=False, group=group)
    pl = place_shards([int(b) for b in totals.cpu().tolist()])
    heads = [stage.head_byte(pl, rank)]
    if world > 1:
        t = torch.tensor(heads, dtype=torch.int64, device="cuda")
        g = torch.empty(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(g, t, group=group)
        heads = [int(x) for x in g.cpu().tolist()]
    hist, first = stage.local_histogram(pl, rank, heads)
    h = torch.from_numpy(hist.astype(np.int64)).cuda()
    f = np.where(first == ABSENT, np.uint64((1 << 62)), first + np.uint64(stage.b0)).astype(np.int64)
    f = torch.from_numpy(f).cuda()
    if world > 1:
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(f, op=dist.ReduceOp.MIN, group=group)
    hist_g = h.cpu().numpy().astype(np.uint32)
    f_g = f.cpu().numpy().astype(np.uint64)
    first_g = np.where(f_g >= np.uint64(1 << 62), np.uint64(ABSENT), f_g)
    bits = stage.encode(hist_g, first_g, rank)
    code_bits = [bits]
    if world > 1:
        t = torch.tensor(code_bits, dtype=torch.int64, device="cuda")
        g = torch.empty(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(g, t, group=group)
        code_bits = [int(x) for x in g.cpu().tolist()]
    hpl, _ = stage.place(code_bits, pl, rank)
    return hpl, stage.d_out[: hpl[rank].nbytes]


class ShardedImageDecoder:
    """Decode of ONE image stream on `world` ranks (ImageDecoder.cpp:88-112): every rank holds the stream, walks
    1 / world of its speculative parse grid (ie_decode_image_shard_begin_dev), the per-group results (16 bytes per group) are
    all-gathered -- the one exchange step of this path -- and every rank decodes its own block rows
    (ie_decode_image_shard_end_dev).  `gather` is the collective: a callable (tensor_out, tensor_in) -> None, by default
    torch.distributed.all_gather_into_tensor on `group`; tests with emulated ranks on one GPU pass their own."""

    def __init__(self, block: int, world: int, rank: int):
        from . import device
        self.block, self.world, self.rank = block, world, rank
        self.sess = device.Session(device.Session.IMAGE_DECODE, 0, 0, block)
        self.d_spec = None

    def begin(self, hdr, d_enc, enc_bytes: int):
        import torch

        from . import device
        total, chunk = device.decode_shard_spec_bytes(enc_bytes, self.block, self.world)
        if self.d_spec is None or self.d_spec.numel() < total:
            self.d_spec = torch.empty(total, dtype=torch.uint8, device="cuda")
        self.chunk, self.total = chunk, total
        device.decode_image_shard_begin_dev(self.sess, hdr, d_enc, enc_bytes, self.rank, self.world, self.d_spec)

    def halves(self):
        """the two arrays to all-gather: [(whole half, this rank's chunk of it)] x 2, views of the spec buffer"""
        half = self.world * self.chunk
        out = []
        for k in range(2):
            whole = self.d_spec[k * half:(k + 1) * half]
            out.append((whole, whole[self.rank * self.chunk:(self.rank + 1) * self.chunk]))
        return out

    def end(self, hdr, d_enc, enc_bytes: int, d_rows_out=None):
        import torch

        from . import device
        r0, r1 = shard_block_rows(hdr.height, self.block, self.world, self.rank)
        rows0, rows1 = r0 // self.block, r1 // self.block
        if d_rows_out is None:
            d_rows_out = torch.empty((r1 - r0) * hdr.width, dtype=torch.uint8, device="cuda")
        if rows0 == rows1:                      # more ranks than block rows: nothing to decode here
            return d_rows_out
        device.decode_image_shard_end_dev(self.sess, hdr, d_enc, enc_bytes, self.world, self.d_spec, rows0, rows1, d_rows_out)
        return d_rows_out

    def decode(self, hdr, d_enc, enc_bytes: int, d_rows_out=None, group=None):
        """this rank's rows of the image (device tensor); asynchronous on the current stream"""
        import torch.distributed as dist
        self.begin(hdr, d_enc, enc_bytes)
        if self.world > 1:
            for whole, mine in self.halves():
                dist.all_gather_into_tensor(whole, mine.clone(), group=group)
        return self.end(hdr, d_enc, enc_bytes, d_rows_out)


class ShardedVideoEncoder:
    """Encodes this rank's whole GOPs of a clip (BASELINE config 5) and re-aligns the result for the single output stream.
    GOPs are independent (VideoBase.hpp:32), so the only exchange is the all-gather of one u64 per rank."""

    def __init__(self, width: int, height: int, total_frames: int, gop: int, world: int, rank: int):
        import torch

        from . import device
        from ._lib import check, lib
        self.width, self.height, self.gop = width, height, max(1, gop)
        self.f0, self.f1 = shard_gops(total_frames, self.gop, world, rank)
        self.frames = self.f1 - self.f0
        self.sess = device.Session(device.Session.VIDEO_ENCODE, width, height, 4, max(1, self.frames))
        check(lib().ie_session_set_video_shard(self.sess.h, total_frames, int(rank == 0)))
        cap = int(lib().ie_max_encoded_bytes(width, height, 4, max(1, self.frames))) + 4096
        self.d_local = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        self.d_aligned = torch.zeros(cap + 32, dtype=torch.uint8, device="cuda")
        self.d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.d_params = torch.zeros(2, dtype=torch.int64, device="cuda")

    def encode(self, d_yuv_shard, quant, rle: bool, merange: int, rank: int, lead_bit: bool = True, group=None):
        """d_yuv_shard: this rank's frames [f0, f1) (YUV420, rebuilt in place).  Returns the device tensor of all ranks' bit
        totals; self.d_aligned holds this rank's bytes of the global stream from byte (offset // 128) * 16."""
        import torch
        import torch.distributed as dist

        from . import device
        if self.frames or rank == 0:
            device.encode_video_dev(self.sess, d_yuv_shard, self.width, self.height, quant, rle, self.gop, merange, self.d_local,
                                    self.d_bits, lead_bit=lead_bit)
        else:
            self.d_bits.zero_()
        totals, offsets = exchange_bit_totals(self.d_bits, group)
        self.d_params[0] = self.d_bits[0]
        self.d_params[1] = offsets[rank]
        device.stream_shift_dev(self.d_local, self.d_params, self.d_aligned)
        return totals
