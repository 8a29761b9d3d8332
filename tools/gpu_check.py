"""Quick on-GPU parity check of the CUDA path against the oracle (development aid; the real tests are tests/)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from thesis_code_b200 import _gpx, GPModel, synthetic

def nrm(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

h = _gpx.Handle(0)
rng = np.random.RandomState(0)
# K-build parity
for name in ["GPyRBF", "GPyMatern52", "GPyMatern32", "GPyExponential"]:
    for (N, d, ARD) in [(300, 3, True), (129, 1, False), (1000, 8, True)]:
        X = rng.rand(N, d) * 3; X2 = rng.rand(77, d) * 3
        ls = (0.5 + rng.rand(d)) if ARD else np.array([0.7])
        K = h.kernel_matrix(X, None, _gpx.KERNEL_IDS[name], 1.3, ls, 0.01)
        Ko = np.tril(o.kernel_matrix(name, X, None, 1.3, ls, ARD) + 0.01 * np.eye(N))
        Kc = h.kernel_matrix(X, X2, _gpx.KERNEL_IDS[name], 1.3, ls)
        Kco = o.kernel_matrix(name, X, X2, 1.3, ls, ARD)
        print("kbuild", name, N, d, "sym err", np.abs(K - Ko).max(), "cross err", np.abs(Kc - Kco).max())

for cfgname, N, Ns in [("C1", 1000, 1000), ("C2", 3000, 700), ("C3", 2500, 333), ("C4", 4000, 5000), ("C5", 5000, 1000)]:
    c = synthetic.make_config(cfgname, N=N, Ns=Ns)
    kw = c['model']['kwargs']
    m = GPModel.from_config(kw)
    t0 = time.time(); m.init(c['X'], c['Y']); t1 = time.time()
    mean, var = m.get_statistics(c['Xs'], full_cov=False); t2 = time.time()
    om = o.OracleGPModel(kernel=kw['kernel'], noise_prior=kw['noise_prior'])
    om.init(c['X'], c['Y'])
    omean, ovar = om.get_statistics(c['Xs'], full_cov=False)
    alpha, L = m._get_handle().export(N, 1, want_L=True)
    lml = m.get_marginal_log_likelihood()
    print(cfgname, N, "fit %.3fs pred %.3fs" % (t1 - t0, t2 - t1), "L", nrm(L, om.state['L']), "alpha", nrm(alpha, om.state['alpha']),
          "mean", nrm(mean, omean), "var", nrm(var, ovar), "lml", abs(lml - om.state['lml']) / abs(om.state['lml']))
    print("   phases", {k: round(v, 3) for k, v in m.phase_times_ms().items()}, "launches", m._get_handle().launch_count())
