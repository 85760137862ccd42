#!/usr/bin/env python
"""Benchmark of the block-transform hot path (BASELINE.json): image encode Mpixels/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1  BASELINE config 2 (the configuration the metric is quoted on for one B200): synthetic 8192x8192, 8x8 blocks,
       matrix8_1.txt, RLE on, Huffman off.  A step = one encode of one image (pixels in HBM -> final stream in HBM);
       successive images alternate between two sessions / CUDA streams, so one image's copy-out kernel overlaps the next
       image's tile kernel.  The line also carries the decode half of the metric, the end-to-end figures through the
       host-buffer C-ABI calls, the reference's CPU time and `parity_sha_ok`: the sha256 of the encoded stream and of the
       decoded pixels against what the UNMODIFIED reference wrote for the same input (tests/golden/golden_configs.json).
N > 1  BASELINE config 3 (the sharded configuration): synthetic 16384x16384, 8x8 blocks, matrix8_2.txt, block rows sharded
       over the ranks, STRONG scaling.  A step = every rank encodes its rows straight into its place of the ONE output
       stream: tile kernel -> exchange of one u64 per rank through peer-mapped mailboxes over NVLink (ie_comm, no NCCL on the
       data path) -> copy-out at the global bit offset.  `value` is the plain (pre-Huffman) stage; the device-side stitch of
       the shards into rank 0's buffer and the Huffman stage are timed separately (extra keys), and the stitched stream's
       sha256 is checked against the reference's once, outside the timed region.

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified reference compiled by
oracle/build_ref.sh, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import hashlib
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

BLOCK = 8
RING = 4                       # distinct input/output buffers cycled through (> L2 in total)
CONFIGS = {
    2: dict(W=8192, H=8192, matrix="matrix8_1.txt", seed=1234, golden="C2|8192x8192|matrix8_1|seed1234",
            workload="config2: synthetic 8192x8192 raw image, 8x8 blocks (matrix8_1.txt), RLE on, Huffman off"),
    3: dict(W=16384, H=16384, matrix="matrix8_2.txt", seed=1235, golden="C3|16384x16384|matrix8_2|seed1235",
            workload="config3: synthetic 16384x16384 raw image, 8x8 blocks (matrix8_2.txt), RLE on, block rows sharded over the "
                     "ranks; value = plain stage, Huffman stage timed separately"),
}
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


def golden(cfg):
    p = ROOT / "tests" / "golden" / "golden_configs.json"
    return json.loads(p.read_text())[cfg["golden"]] if p.exists() else None


def bench_config(cfg, world):
    """The `config` object of the JSON line: identical for both arms."""
    g = golden(cfg)
    px = cfg["W"] * cfg["H"]
    enc = g["plain"]["enc_bytes"] if g else None
    if world == 1:
        return {"workload": cfg["workload"], "image": f"{cfg['W']}x{cfg['H']}", "encoded_bytes_per_image": enc,
                "l2": f"ring of {RING} distinct input/output buffers ({RING * (px + (enc or 0)) >> 20} MiB) > L2, no flush needed",
                "parallelism": "1 GPU, two sessions on two CUDA streams"}
    return {"workload": cfg["workload"], "image": f"{cfg['W']}x{cfg['H']}", "encoded_bytes_per_image": enc,
            "l2": f"ring of {RING} distinct input/output buffers per rank ({RING * (px + (enc or 0)) // world >> 20} MiB) > L2, no flush needed",
            "parallelism": f"block-row shards x{world} ({cfg['H'] // world} pixel rows per rank), ie_comm mailboxes over NVLink"}


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


def quant_matrix(cfg):
    from imageencoder_b200 import read_matrix
    return read_matrix(ROOT / "tests" / "golden" / "inputs" / cfg["matrix"])


def pin_to_gpu_numa_node(local: int):
    """CPU affinity of this rank = the cores next to its GPU, so that pinned staging memory is first touched on that NUMA
    node (eight ranks copying through one node's memory controller is what capped the end-to-end figure in round 1)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference (oracle/_ref) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def run_reference_sample(cfg, rows: int, reps: int = 1):
    """Encodes the first `rows` pixel rows of the workload image with the compiled reference (all host threads).
    Returns (Mpixels/s by process() time per rep, threads, per-rep ms)."""
    import oracle
    from imageencoder_b200.synth import synth_rows
    W = cfg["W"]
    img = synth_rows(W, 0, rows, cfg["seed"])
    workdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    _, res = oracle.ref_image_encode(img, W, rows, BLOCK, quant_matrix(cfg), True, False, threads=os.cpu_count(), reps=reps,
                                     workdir=workdir)
    ms = res["process_ms"]
    run_reference_sample.last_io_ms = float(res.get("ctor_ms", 0.0)) + float(res.get("save_ms", 0.0))     # file read + file write
    return [W * rows / (m / 1e3) / 1e6 for m in ms], res["threads"], ms


def run_port_sample(cfg, rows: int, reps: int = 1):
    """The oracle port (oracle/oracle_block.c, one thread): used only when the compiled reference (oracle/_ref) is missing."""
    import oracle
    from imageencoder_b200.synth import synth_rows
    W = cfg["W"]
    img = synth_rows(W, 0, rows, cfg["seed"])
    q = quant_matrix(cfg)
    ms = []
    for _ in range(reps):
        t = time.perf_counter()
        oracle.image_encode(img, W, rows, BLOCK, q, True, False)
        ms.append((time.perf_counter() - t) * 1e3)
    return [W * rows / (m / 1e3) / 1e6 for m in ms], 1, ms


def reference_decode_sample(cfg, rows: int):
    """reference ImageDecoder::process() (ImageDecoder.cpp:55-122) on the stream of the first `rows` rows; (Mpixels/s, threads, ms)"""
    import oracle
    from imageencoder_b200.synth import synth_rows
    W = cfg["W"]
    img = synth_rows(W, 0, rows, cfg["seed"])
    enc = oracle.image_encode(img, W, rows, BLOCK, quant_matrix(cfg), True, False)
    workdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    _, res = oracle.ref_image_decode(enc, BLOCK, W, rows, threads=os.cpu_count(), workdir=workdir)
    ms = res["process_ms"][-1]
    reference_decode_sample.last_io_ms = float(res.get("ctor_ms", 0.0)) + float(res.get("save_ms", 0.0))
    return W * rows / (ms / 1e3) / 1e6, res["threads"], ms


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    import oracle
    cfg = CONFIGS[2 if world == 1 else 3]
    W, H = cfg["W"], cfg["H"]
    kind = "reference" if oracle.ref_available(BLOCK, False) else "port"
    sample = run_reference_sample if kind == "reference" else run_port_sample
    # a step = the whole image when warmup + steps of it end within a few minutes, else its first rows (probe with 256 rows)
    probe, threads, pms = sample(cfg, 256)
    reps = max(1, args.steps + args.warmup)
    budget_s = 200.0 / reps
    rows = int(min(H, max(256, probe[0] * 1e6 * budget_s / W)) // 8 * 8)
    vals, threads, ms = sample(cfg, rows, reps=reps)
    vals, ms = vals[args.warmup:] or vals, ms[args.warmup:] or ms
    value = W * rows * len(ms) / (sum(ms) / 1e3) / 1e6
    what = ("reference ImageEncoder::process() (OpenMP)" if kind == "reference"
            else "oracle port (oracle_block.c, single thread; oracle/_ref missing)")
    whole = "the whole image" if rows == H else f"the first {rows} of {H} pixel rows"
    line = {
        "impl": "reference", "metric": "image encode Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(ms) / len(ms), "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(cfg, world),
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": threads, "kind": kind,
                         "sample": f"{whole} per step, {len(ms)} reps, process() only, {what}"},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_objects(cfg):
    """encode (and decode) of a bounded stripe by the compiled reference on the box's host cores: ~10-30 s of CPU work"""
    out = {}
    try:
        import oracle
        W = cfg["W"]
        if oracle.ref_available(BLOCK, False):
            probe, threads, _ = run_reference_sample(cfg, 256)
            rows = int(min(cfg["H"], max(256, probe[0] * 1e6 * 12.0 / W)) // 8 * 8)
            vals, threads, ms = run_reference_sample(cfg, rows)
            io = getattr(run_reference_sample, "last_io_ms", 0.0)
            out["cpu_baseline"] = {"value": vals[0], "unit": "Mpixels/s", "cores": threads, "kind": "reference",
                                   "sample": f"{W}x{rows} stripe of the workload image, reference ImageEncoder::process() "
                                             f"(OpenMP, {ms[0]:.0f} ms)",
                                   # SURVEY 8d timing (i): what the reference's own "Elapsed time" covers (main.cpp:68, 111-112):
                                   # constructor (file read) + process() + saveResult() (file write), files on /dev/shm
                                   "value_incl_file_io": W * rows / ((ms[0] + io) / 1e3) / 1e6, "file_io_ms": io}
            drows = int(min(cfg["H"], max(256, rows // 2)) // 8 * 8)
            dv, dthreads, dms = reference_decode_sample(cfg, drows)
            dio = getattr(reference_decode_sample, "last_io_ms", 0.0)
            out["decode_cpu_baseline"] = {"value": dv, "unit": "Mpixels/s", "cores": dthreads, "kind": "reference",
                                          "sample": f"stream of a {W}x{drows} stripe of the workload image, reference "
                                                    f"ImageDecoder::process() (OpenMP, {dms:.0f} ms)",
                                          "value_incl_file_io": W * drows / ((dms + dio) / 1e3) / 1e6, "file_io_ms": dio}
        else:       # oracle/_ref did not travel: the oracle port, one thread, on a smaller stripe
            probe, threads, _ = run_port_sample(cfg, 64)
            rows = int(min(cfg["H"], max(64, probe[0] * 1e6 * 12.0 / W)) // 8 * 8)
            vals, threads, ms = run_port_sample(cfg, rows)
            out["cpu_baseline"] = {"value": vals[0], "unit": "Mpixels/s", "cores": threads, "kind": "port",
                                   "sample": f"{W}x{rows} stripe of the workload image, oracle port (oracle_block.c, "
                                             f"single thread, {ms[0]:.0f} ms; oracle/_ref missing)"}
    except Exception as e:      # the baseline is a reported number; never fail the bench on it
        out.setdefault("cpu_baseline", {"value": None, "unit": "Mpixels/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"})
    return out


# ----------------------------------------------------------------------------------------------------------------
# our arm, one GPU: config 2
# ----------------------------------------------------------------------------------------------------------------
def ours_single(args):
    import torch

    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib, device
    from imageencoder_b200.synth import synth_image

    cfg = CONFIGS[2]
    W, H = cfg["W"], cfg["H"]
    px = W * H
    torch.cuda.set_device(0)
    _lib.check(ie.lib().ie_init(0))
    q = quant_matrix(cfg)
    g = golden(cfg)

    # inputs: RING distinct variants (rolled by whole block rows) so that successive steps never find their input in L2
    base = synth_image(W, H, cfg["seed"])
    cap = int(ie.lib().ie_max_encoded_bytes(W, H, BLOCK, 1))
    d_raw = [torch.from_numpy(np.roll(base, 8 * 37 * i, axis=0).copy()).cuda().reshape(-1) for i in range(RING)]
    d_out = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(RING)]
    d_bits = torch.zeros(RING, dtype=torch.int64, device="cuda")
    NS = 2                                                             # sessions / streams alternated by successive images
    sess = [device.Session(device.Session.IMAGE_ENCODE, W, H, BLOCK) for _ in range(NS)]
    streams = [torch.cuda.Stream() for _ in range(NS)]

    # one C-ABI call per step, arguments prepared once (the GPU step is ~0.1 ms: Python overhead per call must stay far below)
    import ctypes as C
    L = ie.lib()
    qa, qp = device._q(q)
    calls = {}
    for k in range(RING):
        for j in range(NS):
            calls[(k, j)] = (sess[j].h, C.c_void_p(d_raw[k].data_ptr()), W, H, qp, 1, 1, 1, 0, C.c_void_p(d_out[k].data_ptr()),
                             C.c_size_t(cap), C.c_void_p(d_bits[k:k + 1].data_ptr()), C.c_void_p(streams[j].cuda_stream))

    def step(i):
        rc = L.ie_encode_image_dev(*calls[(i % RING, i % NS)])
        if rc:
            _lib.check(rc)

    def sync():
        torch.cuda.synchronize()

    for k in range(2 * RING):      # set-up, not warm-up: sessions allocate their scratch on first use
        step(k)
    sync()
    for i in range(args.warmup):
        step(i)
    sync()
    out_bytes = [int((int(b) + 7) // 8) for b in d_bits.cpu().tolist()]
    if 0 in out_bytes:
        raise SystemExit("warm-up produced an empty stream")

    # ---- parity: slot 0 holds the un-rolled workload image -> its stream and its decoded pixels against the reference's
    parity = {"parity_sha_ok": None}
    stream0 = d_out[0][: out_bytes[0]].cpu().numpy().tobytes()
    if g:
        dec = ie.decode_image(stream0, BLOCK)
        parity = {"parity_sha_ok": bool(sha(stream0) == g["plain"]["enc_sha256"] and sha(dec.tobytes()) == g["plain"]["dec_sha256"]
                                        and sha(base) == g["input_sha256"]),
                  "parity": {"encoded_sha256": sha(stream0), "reference_sha256": g["plain"]["enc_sha256"],
                             "decoded_matches_reference": bool(sha(dec.tobytes()) == g["plain"]["dec_sha256"]),
                             "source": "tests/golden/golden_configs.json (unmodified reference, oracle/_ref)"}}

    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ie.launch_count()
    sync()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(streams[0])
    for s in streams[1:]:
        s.wait_stream(streams[0])                                      # nothing starts before t0
    for i in range(args.steps):
        step(args.warmup + i)
    for s in streams[1:]:
        streams[0].wait_stream(s)
    t1.record(streams[0])
    sync()
    total_ms = t0.elapsed_time(t1)
    launches = ie.launch_count() - launches0

    # one image at a time on one stream: the duration of the two kernels of a step without any overlap
    iso = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(streams[0])
        _lib.check(L.ie_encode_image_dev(*calls[(i % RING, 0)]))
        b.record(streams[0])
        sync()
        iso.append(a.elapsed_time(b))
    iso_ms = statistics.median(iso)

    # ---- decode of the same stream (second half of BASELINE's "encode/decode Mpixels/s"), device-resident
    #      the RING streams alternate between two sessions / streams like the encoder: the latency-bound parse kernels of one
    #      image overlap the block decode of the previous one
    sess_d = [device.Session(device.Session.IMAGE_DECODE, W, H, BLOCK) for _ in range(NS)]
    d_dec = [torch.empty(px, dtype=torch.uint8, device="cuda") for _ in range(NS)]
    # the header is parsed once on the host (ie_parse_image_header, as a file reader would); the decodes are then fully asynchronous
    hdrs = [device.parse_image_header(d_out[k][:160].cpu().numpy().tobytes(), BLOCK) for k in range(RING)]
    dcalls = {(k, j): (sess_d[j].h, C.byref(hdrs[k]), C.c_void_p(d_out[k].data_ptr()), C.c_size_t(out_bytes[k]), C.c_void_p(d_dec[j].data_ptr()),
                       C.c_size_t(px), C.c_void_p(streams[j].cuda_stream)) for k in range(RING) for j in range(NS)}

    def dstep(i, j=None):
        rc = L.ie_decode_image_with_header_dev(*dcalls[(i % RING, i % NS if j is None else j)])
        if rc:
            _lib.check(rc)

    for i in range(2 * RING):
        dstep(i)
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dreps = 16
    a.record(streams[0])
    streams[1].wait_stream(streams[0])
    for i in range(dreps):
        dstep(i)
    streams[0].wait_stream(streams[1])
    b.record(streams[0])
    sync()
    dec_ms = a.elapsed_time(b) / dreps
    a.record(streams[0])
    for i in range(dreps):
        dstep(i, 0)
    b.record(streams[0])
    sync()
    dec_iso_ms = a.elapsed_time(b) / dreps

    # ---- e2e: the public host-buffer calls (pinned host memory in, pinned host memory out), copies inside the region
    h_raw = torch.from_numpy(base).reshape(-1).pin_memory()
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_raw_np, h_out_np = h_raw.numpy(), h_out.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        n_e2e = ie.encode_image(h_raw_np, W, H, q, True, False, out=h_out_np)      # warm-up (allocates staging)
    sync()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        n_e2e = ie.encode_image(h_raw_np, W, H, q, True, False, out=h_out_np)
    e2e_s = time.perf_counter() - te
    h_dec = torch.empty(px, dtype=torch.uint8).pin_memory()
    h_dec_np = h_dec.numpy()
    enc_np = h_out_np[:n_e2e]
    for _ in range(2):
        ie.decode_image(enc_np, BLOCK, out=h_dec_np)
    te = time.perf_counter()
    for _ in range(e2e_steps):
        ie.decode_image(enc_np, BLOCK, out=h_dec_np)
    e2e_dec_s = time.perf_counter() - te
    clocks = sampler.stop()

    peak, peak_src, _ = load_peaks()
    value = px * args.steps / (total_ms / 1e3) / 1e6
    step_ms = total_ms / args.steps
    s_out = out_bytes[0]
    alg_bytes = px + s_out                                   # W*H in + stream out
    achieved = alg_bytes / (step_ms / 1e3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("encode_step_dram_bytes")
        except Exception:
            traffic = None
    dec_alg = s_out + px
    line = {
        "metric": "image encode Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(cfg, 1),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "encode step = encode_tiles_kernel<8,1,0,1,2> + tile_copyout_words_kernel<8,128>; achieved = algorithmic "
                               "bytes / device time per step with the copy-out of one image overlapping the tile kernel of the next "
                               "(two streams); kernel_ms_isolated = the same two kernels for one image alone on one stream",
                     "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": step_ms, "kernel_ms_isolated": iso_ms,
                     "frac_isolated": alg_bytes / (iso_ms / 1e3) / 1e9 / peak},
        "e2e": {"value": px * e2e_steps / e2e_s / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": px,
                "d2h_bytes_per_step": int(n_e2e), "steps": e2e_steps,
                "api": "ie_encode_image (C-ABI, pinned host buffers)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "decode": {"value": px / (dec_ms / 1e3) / 1e6, "unit": "Mpixels/s", "ms": dec_ms, "n_gpus": 1,
                   "ms_one_stream": dec_iso_ms,
                   "what": "ie_decode_image_dev of the workload streams (parallel parse + guarded inverse transform), HBM-resident, two "
                           "sessions / streams, header parsed once on the host (ie_decode_image_with_header_dev)",
                   "roofline": {"bound": "hbm", "achieved": dec_alg / (dec_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": dec_alg / (dec_ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(dec_alg),
                                "kernel": "parse_spec_* (4 launches) + decode_blocks_lean_kernel<8>"},
                   "e2e": {"value": px * e2e_steps / e2e_dec_s / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": int(n_e2e),
                           "d2h_bytes_per_step": px, "steps": e2e_steps, "api": "ie_decode_image (C-ABI, pinned host buffers)"}},
    }
    line.update(parity)
    if not args.no_cpu_baseline:
        cb = cpu_baseline_objects(cfg)
        line["cpu_baseline"] = cb.get("cpu_baseline")
        if "decode_cpu_baseline" in cb:
            line["decode"]["cpu_baseline"] = cb["decode_cpu_baseline"]
    if not args.no_other_configs:
        # the other BASELINE configurations on this GPU, outside everything timed above (own processes, bounded): config 5
        # (video, device-resident, sha256 of stream and frames checked) and a 16-image sample of config 4 (host entry points, end
        # to end).  Extra evidence only: a failure or a time-out never fails the bench.
        del d_raw, d_out
        torch.cuda.empty_cache()
        line["other_configs"] = {"c5_video": run_tool(["tools/bench_sharded.py", "video", "--reps", "3"], 150),
                                 "c4_batch_sample": run_tool(["tools/bench_sharded.py", "batch", "--images", "16", "--reps", "2"], 120)}
    print(json.dumps(line))


def run_tool(argv, timeout_s):
    """last line of a tool's stdout as JSON (the tools print one object), or why there is none"""
    import subprocess
    try:
        r = subprocess.run([sys.executable] + argv, cwd=str(ROOT), capture_output=True, text=True, timeout=timeout_s)
        lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"failed": (r.stderr or r.stdout)[-300:]}
        return json.loads(lines[-1])
    except Exception as e:
        return {"failed": repr(e)}


# ----------------------------------------------------------------------------------------------------------------
# our arm, N GPUs: config 3, block rows sharded, strong scaling
# ----------------------------------------------------------------------------------------------------------------
def ours_sharded(args):
    import torch
    import torch.distributed as dist

    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib, device
    from imageencoder_b200.parallel import (Comm, ShardedHuffmanStage, ShardedImageEncoder, merge_shard_into, shard_block_rows,
                                            sharded_image_encode_huffman, total_bytes)
    from imageencoder_b200.synth import synth_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.check(ie.lib().ie_init(local))
    cfg = CONFIGS[3]
    W, H = cfg["W"], cfg["H"]
    q = quant_matrix(cfg)
    g = golden(cfg)
    y0, y1 = shard_block_rows(H, BLOCK, world, rank)
    hs = y1 - y0
    base = synth_rows(W, y0, y1, cfg["seed"])
    cap = int(ie.lib().ie_max_encoded_bytes(W, hs, BLOCK, 1)) + 64
    d_raw = [torch.from_numpy(np.roll(base, 8 * 37 * i, axis=0).copy() if i else base).cuda().reshape(-1) for i in range(RING)]
    d_out = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(RING)]
    d_bits = torch.zeros(RING, dtype=torch.int64, device="cuda")
    d_first = torch.zeros(RING, dtype=torch.int64, device="cuda")
    NS = 2                                                             # sessions / streams alternated by successive steps
    sessions = [device.Session(device.Session.IMAGE_ENCODE, W, hs, BLOCK) for _ in range(NS)]
    sess = sessions[0]
    streams = [torch.cuda.Stream() for _ in range(NS)]
    comm = Comm(rank, world, stitch_bytes=int(ie.lib().ie_max_encoded_bytes(W, H, BLOCK, 1)))
    # one C-ABI call per step, arguments prepared once.  Successive steps alternate between two sessions / streams (every rank
    # in the same order), so one shard's copy-out overlaps the next step's tile kernel; the mailboxes are double-buffered
    import ctypes as C
    L = ie.lib()
    qa, qp = device._q(q)
    calls = {}
    for k in range(RING):
        for j in range(NS):
            calls[(k, j)] = (sessions[j].h, comm.h, C.c_void_p(d_raw[k].data_ptr()), W, hs, H, qp, 1, 1, C.c_void_p(d_out[k].data_ptr()),
                             C.c_size_t(cap), C.c_void_p(d_bits[k:k + 1].data_ptr()), C.c_void_p(d_first[k:k + 1].data_ptr()),
                             C.c_void_p(streams[j].cuda_stream))

    def step(i):
        rc = L.ie_encode_image_shard_dev(*calls[(i % RING, i % NS)])
        if rc:
            _lib.check(rc)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for k in range(RING):
        step(k)
    barrier()
    for i in range(args.warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ie.launch_count()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(streams[0])
    for st in streams[1:]:
        st.wait_stream(streams[0])                                     # nothing starts before t0
    for i in range(args.steps):
        step(args.warmup + i)
    for st in streams[1:]:
        streams[0].wait_stream(st)
    t1.record(streams[0])
    barrier()
    total_ms = t0.elapsed_time(t1)
    launches = ie.launch_count() - launches0
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    # ---- the ONE output stream, outside the timed region: slot 0 (the un-rolled image) encoded once more, every rank stores
    #      its chunks into rank 0's buffer over NVLink, rank 0 hashes the result against the reference's file
    comm.encode_image_shard(sess, d_raw[0], W, hs, H, q, True, d_out[0], d_bits[0:1], d_first[0:1])       # torch's current stream
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    comm.stitch(d_out[0], d_bits[0:1], d_first[0:1])
    b.record()
    tot = comm.totals()
    barrier()
    stitch_ms = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
    dist.all_reduce(stitch_ms, op=dist.ReduceOp.MAX)
    total_bits = int(tot.sum().item())
    nbytes = (total_bits + 7) // 8
    parity_ok = None
    if rank == 0 and g:
        got = comm.stitched(nbytes).cpu().numpy().tobytes()
        parity_ok = bool(sha(got) == g["plain"]["enc_sha256"] and len(got) == g["plain"]["enc_bytes"])
    barrier()

    # ---- sharded DECODE of that one stream (the other half of the metric): the stitched stream is handed to every rank (file
    #      reader work, outside the timed region); timed: every rank walks its share of the parse grid, one NCCL all-gather of
    #      the per-group results, every rank decodes its own block rows.  The gathered pixels are hashed against the reference's.
    dec = None
    try:
        from imageencoder_b200.parallel import ShardedImageDecoder
        d_stream = torch.zeros((nbytes + 15) // 16 * 16 + 64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            d_stream[:nbytes].copy_(comm.stitched(nbytes))
        dist.broadcast(d_stream, src=0)
        hdr = device.parse_image_header(d_stream[:160].cpu().numpy().tobytes(), BLOCK)
        sd = ShardedImageDecoder(BLOCK, world, rank)
        band = torch.empty(W * hs, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            sd.decode(hdr, d_stream, nbytes, band)
        barrier()
        dreps = 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(dreps):
            sd.decode(hdr, d_stream, nbytes, band)
        b.record()
        barrier()
        dms = torch.tensor([a.elapsed_time(b) / dreps], device="cuda", dtype=torch.float64)
        dist.all_reduce(dms, op=dist.ReduceOp.MAX)
        bands = [torch.empty_like(band) for _ in range(world)] if rank == 0 else None
        dist.gather(band, bands, dst=0)
        if rank == 0:
            import hashlib
            hh = hashlib.sha256()
            for t in bands:
                hh.update(t.cpu().numpy().tobytes())
            # the same stream on ONE GPU (rank 0 alone): the base of the decode's strong scaling
            s1d = device.Session(device.Session.IMAGE_DECODE, 0, 0, BLOCK)
            full_out = torch.empty(W * H, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                device.decode_image_with_header_dev(s1d, hdr, d_stream, nbytes, full_out)
            torch.cuda.synchronize()
            a.record()
            for _ in range(5):
                device.decode_image_with_header_dev(s1d, hdr, d_stream, nbytes, full_out)
            b.record()
            torch.cuda.synchronize()
            ms1d = a.elapsed_time(b) / 5
            dec = {"value": W * H / (float(dms.item()) / 1e3) / 1e6, "unit": "Mpixels/s", "ms": float(dms.item()), "n_gpus": world,
                   "parity_sha_ok": (bool(hh.hexdigest() == g["plain"]["dec_sha256"]) if g else None),
                   "one_gpu": {"ms": ms1d, "value": W * H / (ms1d / 1e3) / 1e6},
                   "what": "ie_decode_image_shard_begin_dev / _end_dev (parallel.ShardedImageDecoder): each rank walks 1/N of the "
                           "speculative parse grid, NCCL all-gather of 16 bytes per group, each rank decodes its own block rows; "
                           "one stream (sessions alternate in the encode figure, not here)"}
            del full_out, s1d
        del d_stream, band, sd
    except Exception as e:      # an extra figure; never fail the encode bench on it
        dec = {"failed": repr(e)}
    barrier()

    # ---- Huffman stage of the sharded stream (BASELINE config 3 asks for RLE + Huffman): parallel.py's exchange (byte shared by
    #      two shards, all-reduce of histogram / first occurrences, per-rank coding with the global dictionary), wall clock
    huff = None
    try:
        enc2 = ShardedImageEncoder(W, hs, BLOCK, H)
        stage = ShardedHuffmanStage(enc2)
        from imageencoder_b200.parallel import sharded_image_encode_huffman_dev
        hreps = 3
        h_ms_dev = None
        for fn in (sharded_image_encode_huffman_dev, sharded_image_encode_huffman):       # the host-orchestrated one last: its shards are hashed
            hpl, d_h = fn(enc2, stage, d_raw[0], q, True, rank)       # warm-up
            barrier()
            th = time.perf_counter()
            for _ in range(hreps):
                hpl, d_h = fn(enc2, stage, d_raw[0], q, True, rank)
            barrier()
            h_ms = (time.perf_counter() - th) / hreps * 1e3
            if h_ms_dev is None:
                h_ms_dev = h_ms
                hshards = [None] * world
                dist.all_gather_object(hshards, d_h.cpu().numpy().tobytes())
                dev_ok = None
                if rank == 0:
                    stream = bytearray()
                    for r in range(world):
                        merge_shard_into(stream, hshards[r], hpl[r])
                    dev_ok = bool(g and sha(bytes(stream[: total_bytes(hpl)])) == g["huff"]["enc_sha256"])
        hshards = [None] * world
        dist.all_gather_object(hshards, d_h.cpu().numpy().tobytes())
        if rank == 0:
            stream = bytearray()
            for r in range(world):
                merge_shard_into(stream, hshards[r], hpl[r])
            hgot = bytes(stream[: total_bytes(hpl)])
            huff = {"ms_per_image_plain_plus_huffman": h_ms, "ms_per_image_device_resident": h_ms_dev, "device_resident_parity_sha_ok": dev_ok,
                    "encoded_bytes": len(hgot),
                    "parity_sha_ok": (bool(sha(hgot) == g["huff"]["enc_sha256"]) if g else None),
                    "what": "block-row sharded encode + Huffman stage over the global histogram, wall clock per image, one image at a "
                            "time.  ms_per_image_plain_plus_huffman: parallel.sharded_image_encode_huffman (host-orchestrated: NCCL "
                            "all-reduce of 256 bins + first occurrences, host tree, eight synchronisations); "
                            "ms_per_image_device_resident: sharded_image_encode_huffman_dev (the three exchanges on device tensors, "
                            "dictionary by stream-ordered host callback, two synchronisations)"}
        del enc2, stage
    except Exception as e:      # an extra figure; never fail the encode bench on it
        huff = {"failed": str(e)}
    barrier()

    # ---- strong-scaling base: the same image on ONE GPU (rank 0 alone, the others idle), same two-stream schedule
    base1 = None
    if rank == 0:
        try:
            from imageencoder_b200.synth import synth_image
            full = torch.from_numpy(synth_image(W, H, cfg["seed"])).cuda().reshape(-1)
            cap1 = int(ie.lib().ie_max_encoded_bytes(W, H, BLOCK, 1))
            o1 = [torch.empty(cap1, dtype=torch.uint8, device="cuda") for _ in range(NS)]
            b1 = torch.zeros(NS, dtype=torch.int64, device="cuda")
            s1 = [device.Session(device.Session.IMAGE_ENCODE, W, H, BLOCK) for _ in range(NS)]
            c1 = [(s1[j].h, C.c_void_p(full.data_ptr()), W, H, qp, 1, 1, 1, 0, C.c_void_p(o1[j].data_ptr()), C.c_size_t(cap1),
                   C.c_void_p(b1[j:j + 1].data_ptr()), C.c_void_p(streams[j].cuda_stream)) for j in range(NS)]
            for i in range(4):
                _lib.check(L.ie_encode_image_dev(*c1[i % NS]))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n1 = 6
            e0.record(streams[0])
            streams[1].wait_stream(streams[0])
            for i in range(n1):
                _lib.check(L.ie_encode_image_dev(*c1[i % NS]))
            streams[0].wait_stream(streams[1])
            e1.record(streams[0])
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / n1
            ok1 = bool(g and sha(o1[0][: (int(b1[0].item()) + 7) // 8].cpu().numpy().tobytes()) == g["plain"]["enc_sha256"])
            base1 = {"n_gpus": 1, "value": W * H / (ms1 / 1e3) / 1e6, "unit": "Mpixels/s", "ms_per_step": ms1, "parity_sha_ok": ok1,
                     "what": "the same 16384x16384 image encoded on rank 0's GPU alone (ie_encode_image_dev, two sessions / streams): "
                             "the base of this strong-scaling series (bench.py --gpus 1 measures config 2)"}
            del full, o1, s1
        except Exception as e:
            base1 = {"failed": str(e)}
    barrier()

    # ---- e2e: pinned host rows in -> shard encode -> this rank's bytes of the stream back in pinned host memory
    h_raw = torch.from_numpy(base).reshape(-1).pin_memory()
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    d_in = torch.empty(W * hs, dtype=torch.uint8, device="cuda")
    d_bits_h = torch.zeros(1, dtype=torch.int64).pin_memory()

    def e2e_step():
        d_in.copy_(h_raw, non_blocking=True)
        comm.encode_image_shard(sess, d_in, W, hs, H, q, True, d_out[0], d_bits[0:1], d_first[0:1])
        d_bits_h.copy_(d_bits[0:1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n = (int(d_bits_h.item()) + 127) // 128 * 16
        h_out[:n].copy_(d_out[0][:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return n

    n_e2e = e2e_step()
    barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        n_e2e = e2e_step()
    e2e_s = time.perf_counter() - te
    t = torch.tensor([e2e_s, float(n_e2e)], device="cuda", dtype=torch.float64)
    tm = t.clone()
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    e2e_s = float(tm[0].item())
    d2h_total = int(t[1].item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src, _ = load_peaks()
        px = W * H
        step_ms = total_ms / args.steps
        value = px * args.steps / (total_ms / 1e3) / 1e6
        alg_bytes = px + nbytes                                   # whole image in + whole stream out, all ranks
        achieved = alg_bytes / (step_ms / 1e3) / 1e9
        line = {
            "metric": "image encode Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(cfg, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "kernel": "per rank: encode_tiles_kernel + tile_totals_kernel + shard_exchange_kernel (P2P mailboxes) + "
                                   "stream_init_shard_kernel + tile_copyout_words_kernel",
                         "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": step_ms},
            "e2e": {"value": px * e2e_steps / e2e_s / 1e6, "unit": "Mpixels/s", "h2d_bytes_per_step": px,
                    "d2h_bytes_per_step": d2h_total, "steps": e2e_steps,
                    "api": "ie_encode_image_shard_dev per rank around pinned-host copies of its rows / its bytes of the stream",
                    "cpu_affinity": (f"{ncpu} cores next to each rank's GPU" if ncpu else "not set")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_sha_ok": parity_ok,
            "stitch": {"ms": float(stitch_ms.item()), "bytes": nbytes,
                       "what": "ie_comm_stitch_dev: every rank's chunks -> rank 0's buffer over NVLink, outside the timed region; the "
                               "sha256 of that buffer is what parity_sha_ok compares with the reference's file"},
            "decode": dec,
            "huffman_stage": huff,
            "strong_scaling_base": base1,
        }
        print(json.dumps(line))
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="N = 1: skip the config-4 / config-5 figures appended to the line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
        return
    try:
        import torch
        ok = torch.cuda.is_available()
    except Exception:
        ok = False
    if not ok:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        ours_sharded(args)
    else:
        ours_single(args)


if __name__ == "__main__":
    main()
