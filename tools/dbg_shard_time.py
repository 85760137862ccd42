"""Where the time of a sharded encode step goes (run under torchrun, 2+ GPUs): python -m torch.distributed.run ... tools/dbg_shard_time.py"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.parallel import ShardedImageEncoder
from imageencoder_b200.synth import synth_image
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
_lib.check(ie.lib().ie_init(local))
size = 8192
q = ie.read_matrix('tests/golden/inputs/matrix8_1.txt')
img = synth_image(size, size, 1234 + rank)
d_raw = [torch.from_numpy(np.roll(img, 8 * 37 * i, axis=0).copy()).cuda().reshape(-1) for i in range(4)]
enc = [ShardedImageEncoder(size, size, 8, 32760) for _ in range(4)]
def timeit(fn, n=20):
    for i in range(4): fn(i)
    torch.cuda.synchronize(); 
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for i in range(n): fn(i)
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (t1 - t0) / n * 1e3
def full(i): enc[i % 4].encode(d_raw[i % 4], q, True, rank)
def nocoll(i):
    e = enc[i % 4]
    device.encode_image_begin_dev(e.sess, d_raw[i % 4], q, True, e.d_total, write_header=(rank == 0), width=size, height=size)
    device.encode_image_end_dev(e.sess, e.d_totals if e.d_totals is not None else e.d_total, 0, e.d_aligned, e.d_bits, e.d_first)
tot = torch.zeros(world, dtype=torch.int64, device="cuda"); one = torch.zeros(1, dtype=torch.int64, device="cuda")
def coll_only(i):
    if world > 1: dist.all_gather_into_tensor(tot, one)
r = {"full (gpu ms, host enqueue ms)": timeit(full), "begin+end only": timeit(nocoll), "all_gather only": timeit(coll_only)}
if rank == 0:
    for k, v in r.items(): print(k, v)
if world > 1: dist.destroy_process_group()
