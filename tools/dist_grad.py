"""Optimiser evaluation (gpx_fit + gpx_lml_grad) on a sharded handle, timed (torchrun): python tools/dist_grad.py C3 40000"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from thesis_code_b200 import _gpx, synthetic, distributed as gdist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = _gpx.Handle(lr)
gdist.init_handle_comm(h, dist, rank, world)
cfg, N = sys.argv[1], int(sys.argv[2])
c = synthetic.make_config(cfg, N=N, Ns=8)
kw = c['model']['kwargs']; kk = kw['kernel']['kwargs']
ls = np.atleast_1d(np.array(kk['lengthscale'], dtype=float))
ard = bool(kk.get('ARD', False))
if ard and ls.size == 1:
    ls = np.full(c["d"], ls[0])
kid = _gpx.KERNEL_IDS[kw['kernel']['name']]
for rep in range(3):
    dist.barrier(); t0 = time.perf_counter()
    h.fit(c['X'], c['Y'], kid, kk.get('variance', 1.0), ls, kw['noise_prior'], ARD=ard)
    dist.barrier(); t1 = time.perf_counter()
    g = h.lml_grad(ls.size)
    dist.barrier(); t2 = time.perf_counter()
if rank == 0:
    print("[%d GPUs] %s N=%d: fit %.1f ms, lml_grad %.1f ms, evaluation %.2f TFLOP/s (N^3 flop); grad[:3] = %s" % (
        world, cfg, N, (t1 - t0) * 1e3, (t2 - t1) * 1e3, N ** 3 / (t2 - t0) / 1e12, g[:3]), flush=True)
dist.destroy_process_group()
