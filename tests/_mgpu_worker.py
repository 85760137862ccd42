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
                                            sharded_image_encode_huffman, sharded_image_encode_huffman_dev, total_bytes)
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
    # the same stage device-resident (exchanges on device tensors, dictionary by stream-ordered host callback)
    enc3 = ShardedImageEncoder(W, y1 - y0, N, H)
    stage3 = ShardedHuffmanStage(enc3)
    for _ in range(2):
        hpl3, d_h3 = sharded_image_encode_huffman_dev(enc3, stage3, d_raw, q, True, rank)
    torch.cuda.synchronize()
    hshards3 = [None] * world
    dist.all_gather_object(hshards3, d_h3.cpu().numpy().tobytes())
    # ---- the library's own exchange (ie_comm, csrc/comm.cu): P2P mailboxes instead of the NCCL all-gather, and the single
    #      output stream assembled ON THE DEVICE in rank 0's stitch buffer (no host merge).  Several sizes and both block sizes;
    #      repeated calls on one communicator exercise the epoch / parity logic of the mailboxes.
    from imageencoder_b200 import device
    from imageencoder_b200.parallel import Comm
    comm = Comm(rank, world, stitch_bytes=int(ie.lib().ie_max_encoded_bytes(2048, 2048, 8, 1)))
    comm_results = []
    for (cw, ch, cn, mat, seed) in ((1024, 1024, 8, "matrix8_2.txt", 4242), (2048, 2048, 8, "matrix8_1.txt", 77), (512, 256, 4, "matrix4_2.txt", 78),
                                    (64, 8 * world, 8, "matrix8_1.txt", 79), (1024, 1024, 8, "matrix8_2.txt", 4242)):
        cq = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / mat)
        cimg = synth_image(cw, ch, seed)
        cy0, cy1 = shard_block_rows(ch, cn, world, rank)
        d_r = torch.from_numpy(cimg[cy0:cy1].copy()).cuda().reshape(-1)
        sess = device.Session(device.Session.IMAGE_ENCODE, cw, cy1 - cy0, cn)
        cap = int(ie.lib().ie_max_encoded_bytes(cw, cy1 - cy0, cn, 1)) + 64
        d_o = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        d_b = torch.zeros(1, dtype=torch.int64, device="cuda")
        d_f = torch.zeros(1, dtype=torch.int64, device="cuda")
        for rep in range(3):
            comm.encode_image_shard(sess, d_r, cw, cy1 - cy0, ch, cq, True, d_o, d_b, d_f)
            comm.stitch(d_o, d_b, d_f)
        tot = comm.totals()
        torch.cuda.synchronize()
        dist.barrier()                     # every rank's stitch kernel has finished: the root's buffer is complete
        nbytes = (int(tot.sum().item()) + 7) // 8
        if rank == 0:
            got = comm.stitched(nbytes).cpu().numpy().tobytes()
            comm_results.append((cw, ch, cn, mat, seed, got))
        dist.barrier()
        sess.close()
    comm.close()
    # ---- sharded decode of ONE stream: every rank walks its share of the parse grid, NCCL all-gather of the per-group results,
    #      every rank decodes its own block rows; the bands are gathered and compared with the oracle's decode on rank 0
    from imageencoder_b200.parallel import ShardedImageDecoder
    dec_results = []
    for (cw, ch, cn, mat, seed) in ((1024, 1024, 8, "matrix8_2.txt", 4242), (2048, 1024, 8, "matrix8_1.txt", 77), (512, 256, 4, "matrix4_2.txt", 78)):
        cq = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / mat)
        stream_b = ie.encode_image(synth_image(cw, ch, seed), cw, ch, cq, True, False)      # every rank: the same bytes
        hdr = device.parse_image_header(stream_b[:160], cn)
        d_enc = torch.from_numpy(np.frombuffer(stream_b + bytes(32), np.uint8).copy()).cuda()
        sd = ShardedImageDecoder(cn, world, rank)
        for rep in range(2):
            band = sd.decode(hdr, d_enc, len(stream_b))
        torch.cuda.synchronize()
        bands = [None] * world
        dist.all_gather_object(bands, band.cpu().numpy().tobytes())
        dec_results.append((cw, ch, cn, stream_b, b"".join(bands)))
    if rank == 0:
        import oracle
        for (cw, ch, cn, stream_b, got) in dec_results:
            want = np.asarray(oracle.image_decode(stream_b, cn)[0]).tobytes()
            if got != want:
                ok = False
                msg.append(f"sharded decode {cw}x{ch} {cn}x{cn}: pixels differ")
        for (cw, ch, cn, mat, seed, got) in comm_results:
            cq = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / mat)
            want = oracle.image_encode(synth_image(cw, ch, seed), cw, ch, cn, cq, True, False)
            if got != want:
                ok = False
                msg.append(f"ie_comm stitch {cw}x{ch} {cn}x{cn}: {len(got)} vs {len(want)} bytes, first diff at "
                           f"{next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), -1)}")
        for name, parts, places, huff in (("plain", shards, pl, False), ("huffman", hshards, hpl, True), ("huffman, device-resident", hshards3, hpl3, True)):
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
