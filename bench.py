#!/usr/bin/env python
"""Benchmark of the block-transform hot path (BASELINE.json): image encode Mpixels/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one encode of one synthetic raw image shard (per rank) through the fused sm_100a kernel.  Workload at N=1:
BASELINE config 2 -- synthetic 8192x8192, 8x8 blocks, matrix8_1.txt, RLE on, Huffman off.  At N>1 every rank encodes
its own 8192x8192 block-row stripe of an 8192 x (8192*N) image (weak scaling; block rows are independent, the only
exchange is the all-gather of one bit total per rank).  Prints ONE JSON line (contract in the task).

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified reference compiled by
oracle/build_ref.sh, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H, BLOCK = 8192, 8192, 8
MATRIX = "matrix8_1.txt"
SEED = 1234
RING = 4                       # distinct input/output buffers cycled through (> L2 in total)
WORKLOAD = "config2: synthetic 8192x8192 raw image, 8x8 blocks (matrix8_1.txt), RLE on, Huffman off"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def quant_matrix():
    from imageencoder_b200 import read_matrix
    return read_matrix(ROOT / "tests" / "golden" / "inputs" / MATRIX)


def shard_image(rank: int):
    """Rows [rank*H, (rank+1)*H) of the 8192 x (8192*N) synthetic image."""
    from imageencoder_b200.synth import synth_rows
    return synth_rows(W, rank * H, (rank + 1) * H, SEED)


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference (oracle/_ref) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def run_reference_sample(rows: int, reps: int = 1):
    """Encodes the first `rows` pixel rows of the workload image with the compiled reference (all host threads).
    Returns (Mpixels/s by process() time, threads, per-rep ms)."""
    import oracle
    from imageencoder_b200.synth import synth_rows
    img = synth_rows(W, 0, rows, SEED)
    workdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    _, res = oracle.ref_image_encode(img, W, rows, BLOCK, quant_matrix(), True, False, threads=os.cpu_count(), reps=reps,
                                     workdir=workdir)
    ms = res["process_ms"]
    return [W * rows / (m / 1e3) / 1e6 for m in ms], res["threads"], ms


def run_port_sample(rows: int, reps: int = 1):
    """The oracle port (oracle/oracle_block.c, one thread) on the first `rows` pixel rows of the workload image: used only
    when the compiled reference (oracle/_ref) is missing.  Returns (Mpixels/s per rep, 1, per-rep ms)."""
    import oracle
    from imageencoder_b200.synth import synth_rows
    img = synth_rows(W, 0, rows, SEED)
    q = quant_matrix()
    ms = []
    for _ in range(reps):
        t = time.perf_counter()
        oracle.image_encode(img, W, rows, BLOCK, q, True, False)
        ms.append((time.perf_counter() - t) * 1e3)
    return [W * rows / (m / 1e3) / 1e6 for m in ms], 1, ms


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    kind = "reference" if oracle.ref_available(BLOCK, False) else "port"
    sample = run_reference_sample if kind == "reference" else run_port_sample
    # size the sample so that warmup + steps end within a few minutes: probe with 256 rows first
    probe, threads, pms = sample(256)
    mpx = probe[0]
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    rows = int(min(H, max(256, (mpx * 1e6 * min(budget_s, 20.0)) / W)) // 8 * 8)
    vals, threads, ms = sample(rows, reps=args.steps + args.warmup)
    vals, ms = vals[args.warmup:], ms[args.warmup:]
    value = W * rows * len(ms) / (sum(ms) / 1e3) / 1e6
    what = ("reference ImageEncoder::process() (OpenMP)" if kind == "reference"
            else "oracle port (oracle_block.c, single thread; oracle/_ref missing)")
    line = {
        "impl": "reference", "metric": "image encode Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / len(ms), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"first {rows} of {H} pixel rows per step, process() only"},
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": threads, "kind": kind,
                         "sample": f"{W}x{rows} stripe of the workload image, {len(ms)} reps, {what}"},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist

    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib, device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.check(ie.lib().ie_init(local))
    q = quant_matrix()

    # ---- inputs: this rank's stripe, RING distinct variants (rolled by whole block rows) so that successive steps
    #      never find their input in L2 (ring footprint >> 126 MB)
    base = shard_image(rank)
    cap = int(ie.lib().ie_max_encoded_bytes(W, H, BLOCK, 1))
    d_raw = [torch.from_numpy(np.roll(base, 8 * 37 * i, axis=0).copy()).cuda().reshape(-1) for i in range(RING)]
    d_out = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(RING)]
    d_bits = torch.zeros(RING, dtype=torch.int64, device="cuda")
    sess = device.Session(device.Session.IMAGE_ENCODE, W, H, BLOCK)
    sharded = None
    if world > 1:
        # block-row shards of one 8192 x (8192*world) image: encode, all-gather of ONE u64 per rank, offset scan,
        # re-alignment of the shard to the chunk grid of the single output stream (imageencoder_b200/parallel.py)
        from imageencoder_b200.parallel import ShardedImageEncoder
        if H * world > 32767:
            full_h = None          # > 15-bit header field: the shards are still encoded/placed, only the header height saturates
        else:
            full_h = H * world
        sharded = [ShardedImageEncoder(W, H, BLOCK, full_h) for _ in range(RING)]

    def step(i):
        k = i % RING
        if sharded is None:
            device.encode_image_dev(sess, d_raw[k], q, True, d_out[k], d_bits[k:k + 1])
        else:
            sharded[k].encode(d_raw[k], q, True, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(RING):          # set-up, not warm-up: every ring slot allocates its session scratch on first use
        step(k)
    barrier()
    for i in range(args.warmup):
        step(i)
    barrier()
    if sharded is not None:
        d_bits = torch.cat([sh.d_total for sh in sharded])
    out_bytes = [int((int(b) + 7) // 8) for b in d_bits.cpu().tolist()]
    if 0 in out_bytes[: min(RING, args.warmup)]:
        raise SystemExit("warm-up produced an empty stream")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ie.launch_count()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
    t1.record()
    barrier()
    total_ms = t0.elapsed_time(t1)
    launches = ie.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- decode of the same stream (second half of BASELINE's "encode/decode Mpixels/s"), device-resident, rank-local
    dec_ms = None
    try:
        sess_d = device.Session(device.Session.IMAGE_DECODE, W, H, BLOCK)
        d_dec = torch.empty(W * H, dtype=torch.uint8, device="cuda")
        # (multi-GPU runs: a shard is not a stream of its own, the decode figure is reported by the N=1 run)
        if sharded is None:
            d_stream, nb = d_out[0], out_bytes[0]
            for _ in range(2):
                device.decode_image_dev(sess_d, d_stream, nb, d_dec, 1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            a.record()
            for _ in range(reps):
                device.decode_image_dev(sess_d, d_stream, nb, d_dec, 1)
            b.record()
            torch.cuda.synchronize()
            dec_ms = a.elapsed_time(b) / reps
    except Exception as e:      # decode is an extra figure; never fail the encode bench on it
        dec_ms = None
        dec_err = str(e)

    # ---- e2e: the public host-buffer call (pinned host memory in, pinned host memory out), copies inside the region
    h_raw = torch.from_numpy(base).reshape(-1).pin_memory()
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_raw_np, h_out_np = h_raw.numpy(), h_out.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    n_e2e = ie.encode_image(h_raw_np, W, H, q, True, False, out=h_out_np)      # warm-up (allocates staging)
    ie.encode_image(h_raw_np, W, H, q, True, False, out=h_out_np)
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        n_e2e = ie.encode_image(h_raw_np, W, H, q, True, False, out=h_out_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - te
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src, peaks = load_peaks()
        px = W * H
        value = px * world * args.steps / (total_ms / 1e3) / 1e6
        kernel_ms = statistics.mean(step_ms)
        s_out = statistics.mean(out_bytes[: min(RING, max(1, args.warmup))]) if args.warmup else out_bytes[0]
        alg_bytes = px + s_out                                   # W*H in + stream out
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("encode_step_dram_bytes")
            except Exception:
                traffic = None
        line = {
            "metric": "image encode Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_rank": f"{W}x{H} block-row stripe of an {W}x{H * world} image",
                       "l2": f"ring of {RING} distinct input/output buffers ({RING * (px + cap) >> 20} MiB) > L2, no flush needed",
                       "encoded_bytes_per_image": int(s_out), "parallelism": f"block-row shards x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "encode step = encode_tiles_kernel<8,1,0,1,2> + tile_copyout_fast_kernel<4,8> (the tile kernel is "
                                   "~80 % of it, profiles/r1_launches_v9.csv)",
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": kernel_ms},
            "e2e": {"value": px * world * e2e_steps / e2e_s / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": px,
                    "d2h_bytes_per_step": int(n_e2e) + 16, "steps": e2e_steps,
                    "api": "ie_encode_image (C-ABI, pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "decode": ({"value": px / (dec_ms / 1e3) / 1e6, "unit": "Mpixels/s", "ms": dec_ms, "n_gpus": 1,
                        "what": "ie_decode_image_dev of rank 0's stream (parallel parse + guarded inverse transform), HBM-resident, "
                                "includes one 160-byte header read-back"} if dec_ms else None),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                import oracle
                if oracle.ref_available(BLOCK, False):
                    probe, threads, _ = run_reference_sample(256)
                    rows = int(min(H, max(256, probe[0] * 1e6 * 12.0 / W)) // 8 * 8)
                    vals, threads, ms = run_reference_sample(rows)
                    line["cpu_baseline"] = {"value": vals[0], "unit": "Mpixels/s", "cores": threads, "kind": "reference",
                                            "sample": f"{W}x{rows} stripe of the workload image, reference ImageEncoder::process() "
                                                      f"(OpenMP, {ms[0]:.0f} ms)"}
                else:       # oracle/_ref did not travel: the oracle port, one thread, on a smaller stripe
                    probe, threads, _ = run_port_sample(64)
                    rows = int(min(H, max(64, probe[0] * 1e6 * 12.0 / W)) // 8 * 8)
                    vals, threads, ms = run_port_sample(rows)
                    line["cpu_baseline"] = {"value": vals[0], "unit": "Mpixels/s", "cores": threads, "kind": "port",
                                            "sample": f"{W}x{rows} stripe of the workload image, oracle port (oracle_block.c, "
                                                      f"single thread, {ms[0]:.0f} ms; oracle/_ref missing)"}
            except Exception as e:      # the baseline is a reported number; never fail the bench on it
                line["cpu_baseline"] = {"value": None, "unit": "Mpixels/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
