"""Worker of tests/test_multigpu_gpu.py: one rank of a real multi-GPU sharded encode (NCCL), checked against the oracle."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib
    from imageencoder_b200.parallel import (ShardedHuffmanStage, ShardedImageEncoder, merge_shard_into, place_shards, shard_block_rows,
                                            sharded_image_encode_huffman, total_bytes)
    from imageencoder_b200.synth import synth_image
    _lib.check(ie.lib().ie_init(local))
    W, H, N = 1024, 1024, 8
    q = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / "matrix8_2.txt")
    img = synth_image(W, H, 4242)
    y0, y1 = shard_block_rows(H, N, world, rank)
    d_raw = torch.from_numpy(img[y0:y1].copy()).cuda().reshape(-1)
    ok = True
    msg = []
    # plain stream
    enc = ShardedImageEncoder(W, y1 - y0, N, H)
    totals = enc.encode(d_raw, q, True, rank)
    torch.cuda.synchronize()
    pl = place_shards([int(b) for b in totals.cpu().tolist()])
    mine = enc.d_aligned[: pl[rank].nbytes].cpu().numpy().tobytes()
    shards = [None] * world
    dist.all_gather_object(shards, mine)
    # Huffman stream
    enc2 = ShardedImageEncoder(W, y1 - y0, N, H)
    stage = ShardedHuffmanStage(enc2)
    hpl, d_h = sharded_image_encode_huffman(enc2, stage, d_raw, q, True, rank)
    torch.cuda.synchronize()
    hshards = [None] * world
    dist.all_gather_object(hshards, d_h.cpu().numpy().tobytes())
    if rank == 0:
        import oracle
        for name, parts, places, huff in (("plain", shards, pl, False), ("huffman", hshards, hpl, True)):
            stream = bytearray()
            for r in range(world):
                merge_shard_into(stream, parts[r], places[r])
            got = bytes(stream[: total_bytes(places)])
            want = oracle.image_encode(img, W, H, N, q, True, huff)
            if got != want:
                ok = False
                msg.append(f"{name}: {len(got)} vs {len(want)} bytes")
        print("MGPU_RESULT", "OK" if ok else "FAIL " + "; ".join(msg), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
