"""CPU tests of the multi-GPU host logic: shard ranges, the offset scan, and a world_size-2 gloo run that exchanges
bit totals with the same all-gather the NCCL path uses and stitches two shard streams into the oracle's single stream."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import INPUTS

ROOT = Path(__file__).resolve().parents[1]


def test_shard_ranges():
    from imageencoder_b200.parallel import shard_block_rows, shard_gops, shard_items
    for H, N, world in ((16384, 8, 8), (936, 4, 3), (8, 8, 4)):
        ranges = [shard_block_rows(H, N, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == H
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all((y1 - y0) % N == 0 for y0, y1 in ranges)
    assert [shard_items(1024, 8, r) for r in (0, 7)] == [(0, 128), (896, 1024)]
    sizes = [(lambda f: (f[1] - f[0]) // 12)(shard_gops(240, 12, 8, r)) for r in range(8)]
    assert sizes == [3, 3, 3, 3, 2, 2, 2, 2]                  # ideal speed-up 6.67x at 8 GPUs
    assert shard_gops(10, 4, 2, 1) == (8, 10)


def test_place_shards():
    from imageencoder_b200.parallel import place_shards, total_bytes
    pl = place_shards([549 + 1000, 300, 0, 77])
    assert [p.global_bit for p in pl] == [0, 1549, 1849, 1849]
    assert pl[1].shift == 1549 % 128 and pl[1].byte_offset == 1549 // 128 * 16 and pl[1].shares_first_chunk
    assert not pl[0].shares_first_chunk and pl[2].nbytes == 0
    assert total_bytes(pl) == (1849 + 77 + 7) // 8


def _bits_of(data: bytes, start: int, n: int) -> np.ndarray:
    return np.unpackbits(np.frombuffer(data, np.uint8))[start:start + n]


def _aligned_shard(bits: np.ndarray, shift: int) -> bytes:
    padded = np.concatenate([np.zeros(shift, np.uint8), bits])
    padded = np.concatenate([padded, np.zeros((-len(padded)) % 128, np.uint8)])
    return np.packbits(padded).tobytes()


def _worker(rank: int, world: int, port: int, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import oracle
    from imageencoder_b200.parallel import exchange_bit_totals, merge_shard_into, place_shards, shard_block_rows, total_bytes
    from imageencoder_b200.synth import synth_image
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, N = 128, 96, 8
        quant = oracle.read_matrix(INPUTS / "matrix8_1.txt")
        img = synth_image(W, H, 5, flat=True)
        full, full_bits, coef, bl, lf = oracle.image_encode_plain(img, W, H, N, quant, True, True, stages=True)
        hdr = oracle.header_bits(N, quant, True)
        per_block = 4 + bl.astype(np.int64) + lf.astype(np.int64) * bl.astype(np.int64)
        y0, y1 = shard_block_rows(H, N, world, rank)
        b0, b1 = (y0 // N) * (W // N), (y1 // N) * (W // N)
        start = hdr + int(per_block[:b0].sum())
        nbits = int(per_block[b0:b1].sum()) + (hdr if rank == 0 else 0)      # rank 0's shard stream carries the header
        if rank == 0:
            start = 0
        totals, offsets = exchange_bit_totals(torch.tensor([nbits], dtype=torch.int64))
        assert int(offsets[rank]) == start, (rank, int(offsets[rank]), start)
        pl = place_shards([int(t) for t in totals])
        shard = _aligned_shard(_bits_of(full, start, nbits), pl[rank].shift)
        # stitch on rank 0 (byte copies + one OR-merged chunk per boundary)
        gathered = [None] * world
        dist.all_gather_object(gathered, shard)
        if rank == 0:
            stream = bytearray()
            for r in range(world):
                merge_shard_into(stream, gathered[r], pl[r])
            out = bytes(stream[:total_bytes(pl)])
            q.put(("ok", out == full, len(out), len(full)))
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_offset_scan_and_stitch():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert res[0] == "ok", res
    assert res[1], f"stitched stream differs from the oracle's ({res[2]} vs {res[3]} bytes)"


def _hist_first(b: np.ndarray):
    hist = np.bincount(b, minlength=256).astype(np.int64)
    first = np.full(256, 1 << 62, dtype=np.int64)
    idx = np.arange(len(b), dtype=np.int64)
    for v in np.unique(b):
        first[v] = idx[b == v][0]
    return hist, first


def _huff_worker(rank: int, world: int, port: int, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import oracle
    from imageencoder_b200.parallel import place_shards, shard_block_rows, shard_byte_range
    from imageencoder_b200.synth import synth_image
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, N = 128, 96, 8
        quant = oracle.read_matrix(INPUTS / "matrix8_2.txt")
        img = synth_image(W, H, 6, flat=True)
        full, full_bits, coef, bl, lf = oracle.image_encode_plain(img, W, H, N, quant, True, False, stages=True)   # no lead bit
        hdr = oracle.header_bits(N, quant, False)
        per_block = 4 + bl.astype(np.int64) + lf.astype(np.int64) * bl.astype(np.int64)
        y0, y1 = shard_block_rows(H, N, world, rank)
        b0, b1 = (y0 // N) * (W // N), (y1 // N) * (W // N)
        nbits = int(per_block[b0:b1].sum()) + (hdr if rank == 0 else 0)
        t = torch.tensor([nbits], dtype=torch.int64)
        totals = torch.empty(world, dtype=torch.int64)
        dist.all_gather_into_tensor(totals, t)
        pl = place_shards([int(x) for x in totals])
        lo, hi = shard_byte_range(pl, rank)
        mine = np.frombuffer(full, np.uint8)[lo:hi]                       # the bytes this rank would Huffman-code
        hist, first = _hist_first(mine)
        first = np.where(first == (1 << 62), first, first + lo)
        h, f = torch.from_numpy(hist), torch.from_numpy(first)
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        gh, gf = _hist_first(np.frombuffer(full, np.uint8))
        ranges = [shard_byte_range(pl, r) for r in range(world)]
        cover = ranges[0][0] == 0 and ranges[-1][1] == len(full) and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        if rank == 0:
            q.put(("ok", bool(np.array_equal(h.numpy(), gh)) and bool(np.array_equal(f.numpy(), gf)), cover))
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_huffman_histogram_exchange():
    """The exchange of the sharded Huffman stage: byte ranges tile the plain stream; all-reduced histogram (sum) and
    first-occurrence positions (min) equal those of the whole stream -> every rank builds the reference's dictionary."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_huff_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert res[0] == "ok", res
    assert res[1], "reduced histogram / first occurrences differ from the whole stream's"
    assert res[2], "byte ranges do not tile the stream"


def test_random_shard_placement_and_merge_roundtrip():
    """place_shards / merge_shard_into / shard_byte_range on random bit strings cut at random bit positions: the stitched
    bytes are the original stream, and the byte ranges of the Huffman stage tile it exactly"""
    from imageencoder_b200.parallel import merge_shard_into, place_shards, shard_byte_range, total_bytes
    rng = np.random.default_rng(11)
    for trial in range(40):
        world = int(rng.integers(1, 7))
        nbits = int(rng.integers(200 * world, 6000))
        bits = rng.integers(0, 2, nbits).astype(np.uint8)
        cuts = np.sort(rng.choice(np.arange(1, nbits), size=world - 1, replace=False)) if world > 1 else np.array([], int)
        edges = [0, *[int(c) for c in cuts], nbits]
        totals = [edges[i + 1] - edges[i] for i in range(world)]
        pl = place_shards(totals)
        stream = bytearray()
        for r in range(world):
            assert pl[r].global_bit == edges[r]
            merge_shard_into(stream, _aligned_shard(bits[edges[r]:edges[r + 1]], pl[r].shift), pl[r])
        want = np.packbits(np.concatenate([bits, np.zeros((-nbits) % 8, np.uint8)])).tobytes()
        assert bytes(stream[: total_bytes(pl)]) == want
        ranges = [shard_byte_range(pl, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == len(want)
        assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        assert all(b1 >= b0 for b0, b1 in ranges)
