#!/usr/bin/env python
"""BASELINE configs 4 and 5 on N GPUs (one process per GPU), with the reference's sha256 checked:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py video|batch [...]
    python tools/bench_sharded.py video|batch                                   (N = 1)

video  config 5: 1920x1088, 240 frames, GOP 12, merange 16, matrix.txt.  Whole GOPs per rank (20 GOPs: 3,3,3,3,2,2,2,2 at 8
       ranks -> ideal speed-up 6.67x), device-resident.  Exchange of one u64 per rank through ie_comm (P2P mailboxes), every
       rank re-aligns its shard (ie_stream_shift_dev) and stores it into rank 0's stitch buffer over NVLink; the stitched
       stream's sha256 must be the reference's (tests/golden/golden_configs.json).  Timed: encode of the rank's GOPs (+ the
       exchange and the re-alignment), max over ranks; the stitch separately.
batch  config 4: images of 4096x4096, 4x4 blocks, matrix4_2.txt, `--images` per rank (default 128 = 1024 over 8 ranks), encode +
       decode round trip through the HOST entry points ie_encode_images / ie_decode_images from pinned buffers (the copies
       are inside the region: this is an end-to-end figure).  No communication.  Eight distinct images (seeds 2000..2007)
       repeat cyclically; the first three are checked against the reference's sha256, every decoded image against its source
       stream's decode.
Prints one JSON object (rank 0)."""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


def pin_to_gpu_numa_node(local: int):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["video", "batch"])
    ap.add_argument("--images", type=int, default=128, help="batch: images per rank")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib, device
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.check(ie.lib().ie_init(local))
    L = ie.lib()
    gold = json.loads((ROOT / "tests" / "golden" / "golden_configs.json").read_text())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.what == "video":
        from imageencoder_b200.parallel import Comm, shard_gops
        from imageencoder_b200.synth import synth_video
        g = gold["C5|1920x1088x240|gop12|mer16|matrix|seed4000"]
        W, H, F, gop, mer = g["W"], g["H"], g["frames"], g["gop"], g["merange"]
        q = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / g["matrix"])
        fsz = W * H * 3 // 2
        f0, f1 = shard_gops(F, gop, world, rank)
        nf = f1 - f0
        yuv = synth_video(W, H, F, g["seed"])            # every rank builds the clip (seeded), keeps its frames
        if rank == 0:
            assert sha(yuv) == g["input_sha256"], "synthetic clip differs from the one the reference encoded"
        d_src = torch.from_numpy(np.ascontiguousarray(yuv[f0 * fsz: f1 * fsz])).cuda()
        del yuv
        d_yuv = d_src.clone()
        sess = device.Session(device.Session.VIDEO_ENCODE, W, H, 4, max(1, nf))
        _lib.check(L.ie_session_set_video_shard(sess.h, F, int(rank == 0)))
        cap = int(L.ie_max_encoded_bytes(W, H, 4, max(1, nf))) + 4096
        d_local = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        d_aligned = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
        d_params = torch.zeros(2, dtype=torch.int64, device="cuda")
        d_sbits = torch.zeros(1, dtype=torch.int64, device="cuda")
        d_first = torch.zeros(1, dtype=torch.int64, device="cuda")
        comm = Comm(rank, world, stitch_bytes=int(L.ie_max_encoded_bytes(W, H, 4, F)) + 4096)

        def encode():
            d_yuv.copy_(d_src)                           # the encoder rebuilds the frames in place
            device.encode_video_dev(sess, d_yuv, W, H, q, True, gop, mer, d_local, d_bits, lead_bit=True)
            tot = comm.exchange(d_bits)                  # P2P mailboxes: one u64 per rank
            off = (torch.cumsum(tot, 0) - tot)[rank:rank + 1]
            d_params[0:1] = d_bits
            d_params[1:2] = off
            device.stream_shift_dev(d_local, d_params, d_aligned)
            d_first.copy_(off)
            d_sbits.copy_(d_bits + off % 128)
            return tot

        for _ in range(2):
            tot = encode()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d_yuv.copy_(d_src)
        b.record()
        torch.cuda.synchronize()
        copy_ms = a.elapsed_time(b)
        ts = []
        for _ in range(args.reps):
            barrier()
            a.record()
            tot = encode()
            b.record()
            torch.cuda.synchronize()
            ts.append(allmax(a.elapsed_time(b) - copy_ms))
        enc_ms = float(np.median(ts))
        # the encode alone (ie_encode_video_dev of the rank's GOPs), without the exchange of the shard sizes and the shift onto the
        # stream's bit offset that only a sharded run needs
        tb = []
        for _ in range(args.reps):
            barrier()
            a.record()
            d_yuv.copy_(d_src)
            device.encode_video_dev(sess, d_yuv, W, H, q, True, gop, mer, d_local, d_bits, lead_bit=True)
            b.record()
            torch.cuda.synchronize()
            tb.append(allmax(a.elapsed_time(b) - copy_ms))
        bare_ms = float(np.median(tb))
        tot = encode()
        barrier()
        a.record()
        comm.stitch(d_aligned, d_sbits, d_first)
        b.record()
        barrier()
        stitch_ms = allmax(a.elapsed_time(b))
        nbytes = (int(tot.sum().item()) + 7) // 8
        out = None
        if rank == 0:
            got = comm.stitched(nbytes).cpu().numpy().tobytes()
            ok = sha(got) == g["enc_sha256"] and len(got) == g["enc_bytes"]
            # per I-frame W*H + S_f, per P-frame 2 W*H + S_f + W*H written back
            npf = F - (F + gop - 1) // gop
            alg = (F - npf) * W * H + npf * 3 * W * H + nbytes
            out = {"config": "5: 1920x1088, 240 frames, GOP 12, merange 16, matrix.txt, GOPs sharded", "n_gpus": world,
                   "encode_ms": enc_ms, "encode_gpx_s": W * H * F / enc_ms / 1e6, "encode_ms_kernels_only": bare_ms,
                   "encode_gpx_s_kernels_only": W * H * F / bare_ms / 1e6, "encoded_bytes": nbytes,
                   "algorithmic_GB_s": alg / enc_ms / 1e6, "parity_sha_ok": bool(ok), "stitch_ms": stitch_ms,
                   "gops_per_rank": [(shard_gops(F, gop, world, r)[1] - shard_gops(F, gop, world, r)[0]) // gop for r in range(world)],
                   "ideal_speedup": (F // gop) / max((shard_gops(F, gop, world, r)[1] - shard_gops(F, gop, world, r)[0]) // gop
                                                        for r in range(world))}
        barrier()
        if rank == 0:
            # decode of the stitched stream on one GPU (frame-sequential: the frame boundaries are a chain through the stream)
            d_enc = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
            d_enc[:nbytes] = comm.stitched(nbytes)
            d_dec = torch.empty(fsz * F, dtype=torch.uint8, device="cuda")
            sd = device.Session(device.Session.VIDEO_DECODE, W, H, 4, F)
            device.decode_video_dev(sd, d_enc, nbytes, d_dec, True)
            torch.cuda.synchronize()
            t = time.perf_counter()
            device.decode_video_dev(sd, d_enc, nbytes, d_dec, True)
            torch.cuda.synchronize()
            dec_ms = (time.perf_counter() - t) * 1e3
            out["decode_ms_one_gpu"] = dec_ms
            out["decode_gpx_s_one_gpu"] = W * H * F / dec_ms / 1e6
            out["decode_sha_ok"] = bool(sha(d_dec.cpu().numpy().tobytes()) == g["dec_mc1_sha256"])
            if world == 1:
                # end to end through the host entry points (VideoEncoder / VideoDecoder drop-in boundary): pinned host buffers,
                # the copies inside the timed region (GOP batches go up / come down next to the kernels)
                import ctypes as C
                del d_enc, d_dec
                qa = np.ascontiguousarray(q, dtype=np.uint16).reshape(-1)
                qp = qa.ctypes.data_as(C.POINTER(C.c_uint16))
                h_yuv = torch.from_numpy(np.ascontiguousarray(synth_video(W, H, F, g["seed"])).reshape(-1)).pin_memory()
                cap = int(L.ie_max_encoded_bytes(W, H, 4, F))
                h_enc = torch.empty(cap, dtype=torch.uint8).pin_memory()
                h_dec = torch.empty(h_yuv.numel(), dtype=torch.uint8).pin_memory()
                nb, nby = C.c_size_t(0), C.c_size_t(0)
                w_, h_, f_ = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
                te, td = [], []
                for _ in range(1 + args.reps):
                    t = time.perf_counter()
                    _lib.check(L.ie_encode_video(C.c_void_p(h_yuv.data_ptr()), h_yuv.numel(), W, H, qp, 1, gop, mer, 0,
                                                 C.c_void_p(h_enc.data_ptr()), cap, C.byref(nb)))
                    te.append((time.perf_counter() - t) * 1e3)
                    t = time.perf_counter()
                    _lib.check(L.ie_decode_video(C.c_void_p(h_enc.data_ptr()), nb.value, 1, C.c_void_p(h_dec.data_ptr()), h_dec.numel(),
                                                 C.byref(nby), C.byref(w_), C.byref(h_), C.byref(f_)))
                    td.append((time.perf_counter() - t) * 1e3)
                e_ms, d_ms = float(np.median(te[1:])), float(np.median(td[1:]))
                out["e2e"] = {"api": "ie_encode_video / ie_decode_video (C-ABI, pinned host buffers)",
                              "encode_ms": e_ms, "decode_ms": d_ms, "encode_gpx_s": W * H * F / e_ms / 1e6, "decode_gpx_s": W * H * F / d_ms / 1e6,
                              "h2d_bytes_encode": h_yuv.numel(), "d2h_bytes_encode": nb.value,
                              "parity_sha_ok": bool(sha(h_enc[:nb.value].numpy().tobytes()) == g["enc_sha256"]
                                                    and sha(h_dec.numpy().tobytes()) == g["dec_mc1_sha256"])}
            print(json.dumps(out))
        barrier()
        comm.close()
    else:
        from imageencoder_b200.synth import synth_image
        W = H = 4096
        N = 4
        q = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / "matrix4_2.txt")
        qa = np.ascontiguousarray(q, dtype=np.uint16).reshape(-1)
        import ctypes as C
        qp = qa.ctypes.data_as(C.POINTER(C.c_uint16))
        n = args.images
        npx = W * H
        distinct = min(8, n)
        base = [synth_image(W, H, 2000 + i) for i in range(distinct)]
        h_raw = torch.empty(n * npx, dtype=torch.uint8).pin_memory()
        hr = h_raw.numpy().reshape(n, H, W)
        for i in range(n):
            hr[i] = base[i % distinct]
        slot = int(L.ie_max_encoded_bytes(W, H, N, 1))
        # encoded streams of this content are ~0.4 byte per pixel: slots of 1/2 of the worst case are ample (checked below)
        es = (slot // 2 + 15) // 16 * 16
        h_enc = torch.empty(n * es, dtype=torch.uint8).pin_memory()
        h_dec = torch.empty(n * npx, dtype=torch.uint8).pin_memory()
        sizes = (C.c_size_t * n)()
        wv, hv = C.c_uint32(0), C.c_uint32(0)

        def enc():
            _lib.check(L.ie_encode_images(C.c_void_p(h_raw.data_ptr()), n, W, H, N, qp, 1, 0, C.c_void_p(h_enc.data_ptr()), es, sizes))

        def dec():
            _lib.check(L.ie_decode_images(C.c_void_p(h_enc.data_ptr()), es, sizes, n, N, C.c_void_p(h_dec.data_ptr()), npx, C.byref(wv), C.byref(hv)))

        enc(); dec()
        barrier()
        te, td = [], []
        for _ in range(args.reps):
            barrier()
            t = time.perf_counter(); enc(); te.append(allmax(time.perf_counter() - t))
            barrier()
            t = time.perf_counter(); dec(); td.append(allmax(time.perf_counter() - t))
        e_s, d_s = float(np.median(te)), float(np.median(td))
        ok = True
        he = h_enc.numpy()
        hd = h_dec.numpy().reshape(n, H, W)
        for i in range(min(3, n)):
            gi = gold[f"C4|4096x4096|matrix4_2|seed{2000 + i}"]["plain"]
            ok = ok and sha(he[i * es: i * es + sizes[i]].tobytes()) == gi["enc_sha256"] and sha(hd[i].tobytes()) == gi["dec_sha256"]
        for i in range(distinct, n):                       # repeats of the same image: same stream, same pixels
            j = i % distinct
            ok = ok and sizes[i] == sizes[j] and np.array_equal(hd[i], hd[j])
        oks = [ok]
        if world > 1:
            oks = [None] * world
            dist.all_gather_object(oks, ok)
        if rank == 0:
            tot = n * world
            enc_bytes = sum(int(sizes[i]) for i in range(n))
            print(json.dumps({"config": f"4: {tot} images of 4096x4096, 4x4 blocks (matrix4_2.txt), {n} per rank, host API round trip",
                              "n_gpus": world, "encode_s": e_s, "decode_s": d_s, "encode_gpx_s_e2e": tot * npx / e_s / 1e9,
                              "decode_gpx_s_e2e": tot * npx / d_s / 1e9, "roundtrip_gpx_s_e2e": tot * npx / (e_s + d_s) / 1e9,
                              "h2d_GB_s_per_rank_encode": n * npx / e_s / 1e9, "d2h_GB_s_per_rank_decode": n * npx / d_s / 1e9,
                              "encoded_bytes_per_rank": enc_bytes, "parity_sha_ok": bool(all(oks)),
                              "cpu_affinity": (f"{ncpu} cores next to each rank's GPU" if ncpu else "not set")}))
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
