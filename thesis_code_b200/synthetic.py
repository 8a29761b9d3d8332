"""Synthetic inputs for the five BASELINE.json configs.  All fp64, seeded, hyper-parameters
fixed (do_optimize=False).  The generators restate the reference's environments where one is named:
``Sinc`` (src/environments/smooth.py:14-21), ``GenzOscillatory`` (src/environments/genz1984.py:105-128) and the
Runner's own draws (src/experiment/runner.py:104,115 with ``random_hypercube_samples``, src/utils.py:77-101)."""
import numpy as np


def _hypercube(n, bounds, rng):
    """random_hypercube_samples: uniform (dims, n) draw scaled to the bounds, returned as (n, dims)."""
    bounds = np.asarray(bounds, dtype=np.float64)
    a = rng.uniform(0, 1, (bounds.shape[0], n))
    s = a * np.abs(bounds[:, 1] - bounds[:, 0])[:, None] + bounds[:, 0][:, None]
    return np.clip(s.T, bounds[:, 0], bounds[:, 1])


def sinc(x):
    x = np.asanyarray(x)
    y = np.where(x == 0, 1.0e-20, x)
    return np.sin(y) / y


def genz_oscillatory(X, u=0.5, a=5.0):
    X = np.asarray(X)
    return np.cos(2 * np.pi * u + a * X.sum(axis=1))[:, None]


def _zscore(A):
    return (A - A.mean(axis=0)) / A.std(axis=0)


def make_config(name, N=None, Ns=None):
    """Return dict(X, Y, Xs, model=<GPModel config dict>, name, N, d, Ns).  N / Ns override the config's sizes
    (used by the parity tests to run the same generator at oracle-feasible sizes)."""
    name = name.upper()
    if name == "C1":  # GPModel RBF on Sinc 1-D, N=1000 / 1000 test
        N = N or 1000; Ns = Ns or 1000
        bounds = np.array([[-20.0, 20.0]])
        X = _hypercube(N, bounds, np.random.RandomState(0))
        Y = sinc(X)
        Xs = _hypercube(Ns, bounds, np.random.RandomState(1))
        kernel = {'name': 'GPyRBF', 'kwargs': {'variance': 1.0, 'lengthscale': 1.0}}
        noise = 1e-2
    elif name == "C2":  # ARD-RBF on Genz oscillatory d=10, N=20k
        N = N or 20000; Ns = Ns or 2500
        bounds = np.array([[0.0, 1.0]] * 10)
        X = _hypercube(N, bounds, np.random.RandomState(0))
        Y = genz_oscillatory(X)
        Xs = _hypercube(Ns, bounds, np.random.RandomState(1))
        kernel = {'name': 'GPyRBF', 'kwargs': {'variance': 1.0, 'lengthscale': 0.6, 'ARD': True}}
        noise = 1e-2
    elif name == "C3":  # Matern-5/2 ARD on kin40k-shaped synthetic d=8, N=40k
        N = N or 40000; Ns = Ns or 4000
        d = 8
        rng = np.random.RandomState(0)
        Xall = rng.randn(N + Ns, d)
        rff = np.random.RandomState(2)
        Wf = rff.randn(d, 512) / 1.5
        bf = rff.uniform(0, 2 * np.pi, 512)
        wf = rff.randn(512)
        F = np.sqrt(2.0 / 512) * np.cos(Xall @ Wf + bf) @ wf
        Yall = F + 0.1 * np.random.RandomState(3).randn(N + Ns)
        Xall = _zscore(Xall)
        Yall = _zscore(Yall[:, None])
        X, Xs, Y = Xall[:N], Xall[N:], Yall[:N]
        kernel = {'name': 'GPyMatern52', 'kwargs': {'variance': 1.0, 'lengthscale': list(np.linspace(0.8, 2.0, d)), 'ARD': True}}
        noise = 1e-2
    elif name == "C4":  # 1-D natural-sound-shaped, N=60k with 100k test points
        N = N or 60000; Ns = Ns or 100000
        grid = int(round(N * 65000 / 60000))
        rng = np.random.RandomState(0)
        idx = np.sort(rng.choice(grid, size=N, replace=False))
        X = (idx / float(grid))[:, None]
        sr = np.random.RandomState(4)
        freqs = sr.uniform(40.0, 1500.0, 20) * (grid / 65000.0)
        mods = sr.uniform(1.0, 8.0, 20)
        phases = sr.uniform(0, 2 * np.pi, 20)
        amps = sr.uniform(0.2, 1.0, 20)

        def f(x):
            x = x[:, 0]
            s = np.zeros_like(x)
            for a, fr, m, p in zip(amps, freqs, mods, phases):
                s += a * (1 + 0.5 * np.sin(2 * np.pi * m * x)) * np.sin(2 * np.pi * fr * x + p)
            return s
        Yraw = f(X) + 0.01 * np.random.RandomState(5).randn(N)
        Y = _zscore(Yraw[:, None])
        Xs = np.random.RandomState(1).uniform(0, 1, (Ns, 1))
        kernel = {'name': 'GPyRBF', 'kwargs': {'variance': 1.0, 'lengthscale': 5e-4 * 65000.0 / grid}}
        noise = 1e-3
    elif name == "C5":  # ARD-RBF d=8, N=200k fp64, block-cyclic over 2/4/8 GPUs
        N = N or 200000; Ns = Ns or 2500
        d = 8
        bounds = np.array([[0.0, 1.0]] * d)
        X = _hypercube(N, bounds, np.random.RandomState(0))
        Y = genz_oscillatory(X) + 0.1 * np.random.RandomState(6).randn(N, 1)
        Xs = _hypercube(Ns, bounds, np.random.RandomState(1))
        kernel = {'name': 'GPyRBF', 'kwargs': {'variance': 1.0, 'lengthscale': list(0.6 * (1 + 0.1 * np.arange(d))), 'ARD': True}}
        noise = 1e-2
    else:
        raise KeyError(name)
    model = {'name': 'GPModel', 'kwargs': {'kernel': kernel, 'noise_prior': noise, 'do_optimize': False}}
    return dict(name=name, X=np.ascontiguousarray(X), Y=np.ascontiguousarray(Y), Xs=np.ascontiguousarray(Xs),
                model=model, N=X.shape[0], d=X.shape[1], Ns=Xs.shape[0])
