"""Multi-GPU parity check (torchrun): block-cyclic fit + sharded predict vs the oracle and vs timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch, torch.distributed as dist
from thesis_code_b200 import _gpx, synthetic, distributed as gdist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
h = _gpx.Handle(lr)
if os.environ.get("GPX_LO_CTAS"): h.set_option("nccl_lo_ctas", int(os.environ["GPX_LO_CTAS"]))
gdist.init_handle_comm(h, dist, rank, world)
cases = [("C5", 5000, 1000, 4), ("C2", 3000, 700, 1), ("C3", 2500, 333, 3), ("C1", 1000, 1000, 4), ("C4", 4000, 5000, 2)]
if len(sys.argv) > 1:
    cases = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))]
for cfg, N, Ns, W in cases:
    c = synthetic.make_config(cfg, N=N, Ns=Ns)
    kw = c['model']['kwargs']; kk = kw['kernel']['kwargs']
    ls = np.atleast_1d(np.array(kk['lengthscale'], dtype=float))
    kid = _gpx.KERNEL_IDS[kw['kernel']['name']]
    h.set_option("outer_tiles", W)
    if os.environ.get("GPX_MAX_CHUNK"): h.set_option("gemm_max_chunk", int(os.environ["GPX_MAX_CHUNK"]))
    if os.environ.get("GPX_CACHE_MB"): h.set_option("cache_mb", int(os.environ["GPX_CACHE_MB"]))
    if os.environ.get("GPX_GROUP"): h.set_option("group_panels", int(os.environ["GPX_GROUP"]))
    if os.environ.get("GPX_TRACE"): h.set_option("trace", 1)
    if os.environ.get("GPX_LOOKAHEAD"): h.set_option("lookahead", int(os.environ["GPX_LOOKAHEAD"]))
    for rep in range(2):
        dist.barrier(); t0 = time.time()
        h.fit(c['X'], c['Y'], kid, kk.get('variance', 1.0), ls, kw['noise_prior'], ARD=kk.get('ARD', False))
        dist.barrier(); t1 = time.time()
        mean, var = h.predict(c['Xs'])
        dist.barrier(); t2 = time.time()
    ph = h.phase_ms()
    msg = "[rank %d/%d] %s N=%d Ns=%d W=%d fit %.3fs predict %.3fs chol %.1f ms (%.2f TF/s aggregate) trsm %.1f ms solve %.1f ms lml %.10e" % (
        rank, world, cfg, N, Ns, W, t1 - t0, t2 - t1, ph['cholesky'], (N ** 3 / 3) / (ph['cholesky'] * 1e-3) / 1e12, ph['predict_trsm'], ph['solve'], h.lml()[0])
    if N <= 12000:
        from oracle import gp_oracle as o
        if rank == 0:
            om = o.OracleGPModel(kernel=kw['kernel'], noise_prior=kw['noise_prior']); om.init(c['X'], c['Y'])
            omean, ovar = om.get_statistics(c['Xs'], full_cov=False)
            alpha, _ = h.export(N, 1)
            nrm = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
            msg += " | vs oracle: mean %.2e var %.2e lml %.2e alpha %.2e" % (nrm(mean, omean[0]), nrm(var, ovar[0, :, 0]), abs(h.lml()[0] - om.state['lml']) / abs(om.state['lml']), nrm(alpha, om.state['alpha']))
    print(msg, flush=True)
    if os.environ.get("GPX_TRACE"):
        tr = h.trace()
        np.save("gpurun_out/trace_rank%d.npy" % rank, tr)
dist.destroy_process_group()
