"""ctypes binding of the C-ABI shared library ``libgpx.so`` (include/gpx.h).

This is the only bridge between the Python model API and the sm_100a CUDA kernels.  There is no CPU fallback:
if the library is missing, cannot be loaded, or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPX_LIB") or os.path.join(_HERE, "libgpx.so")

KERNEL_IDS = {"GPyRBF": 0, "GPyMatern52": 1, "GPyMatern32": 2, "GPyExponential": 3}
MEM_HOST, MEM_DEVICE = 0, 1
PHASES = ("kbuild", "cholesky", "solve", "lml", "predict_kstar", "predict_trsm", "predict_reduce", "h2d", "d2h")

_c_double_p = ctypes.POINTER(ctypes.c_double)
_lib = None


class GpxError(RuntimeError):
    """A negative return code from libgpx (bad argument / CUDA / NCCL error)."""


class NotPositiveDefiniteError(np.linalg.LinAlgError):
    """LAPACK-style info > 0 from the Cholesky: first non-positive pivot (1-based row)."""

    def __init__(self, info, msg):
        super().__init__(msg)
        self.info = info


class AppendUnsupported(RuntimeError):
    """gpx_append returned -8: this handle cannot append in place (sharded or normalised); the caller refits."""


# every symbol include/gpx.h declares: (restype, argtypes)
SIGNATURES = {
    "gpx_version": (ctypes.c_int, []),
    "gpx_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "gpx_destroy": (None, [ctypes.c_void_p]),
    "gpx_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gpx_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "gpx_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "gpx_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_int, ctypes.c_int, ctypes.c_double, _c_double_p, ctypes.c_int, ctypes.c_int,
                               ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    "gpx_set_normalizer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "gpx_get_normalizer": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p]),
    "gpx_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "gpx_append": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]),
    "gpx_predict_moments": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "gpx_selfcheck": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int]),
    "gpx_lml": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, _c_double_p]),
    "gpx_lml_grad": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int]),
    "gpx_predict": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "gpx_predict_cov": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "gpx_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "gpx_kernel_matrix": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                         _c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]),
    "gpx_phase_ms": (ctypes.c_double, [ctypes.c_void_p, ctypes.c_int]),
    "gpx_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "gpx_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "gpx_bytes_allocated": (ctypes.c_int64, [ctypes.c_void_p]),
    "gpx_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "gpx_trace": (ctypes.c_int64, [ctypes.c_void_p, _c_double_p, ctypes.c_int64]),
    "gpx_gemm_stats": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_int64)]),
}


def load_library(path=None):
    """Load libgpx.so and attach the prototypes.  Raises if the library was not built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise GpxError("{} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)".format(path))
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class Handle:
    """One gpx handle = one GPU's factor storage, workspaces and stream."""

    def __init__(self, device=0):
        self.lib = load_library()
        hp = ctypes.c_void_p()
        rc = self.lib.gpx_create(ctypes.byref(hp), int(device))
        if rc != 0 or not hp.value:
            reason = {-4: "no CUDA device visible", -5: "device is not sm_100 (B200)", -1: "bad device index"}.get(rc, "CUDA error")
            raise GpxError("gpx_create(device={}) failed with code {} ({}); the exact-GP path has no CPU fallback"
                           .format(device, rc, reason))
        self._h = hp
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.gpx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == 0:
            return
        msg = self.lib.gpx_last_error(self._h).decode("utf-8", "replace")
        if rc > 0:
            raise NotPositiveDefiniteError(rc, msg)
        raise GpxError("libgpx error {}: {}".format(rc, msg))

    def set_option(self, name, value):
        self._check(self.lib.gpx_set_option(self._h, name.encode(), int(value)))

    def comm_init(self, rank, world, uid):
        self._check(self.lib.gpx_comm_init(self._h, int(rank), int(world), uid))
        self.rank, self.world = int(rank), int(world)

    def set_normalizer(self, normalize_input, normalize_output):
        """Fused NormalizerModel: z-score X / Y on the device in every later fit, de-normalise in every predict."""
        self._check(self.lib.gpx_set_normalizer(self._h, int(bool(normalize_input)), int(bool(normalize_output))))

    def get_normalizer(self, d, O):
        """(x_mean, x_std, y_mean, y_std) of the last fit (zeros / ones where a z-score is off)."""
        xm, xs, ym, ys = np.zeros(d), np.ones(d), np.zeros(O), np.ones(O)
        self._check(self.lib.gpx_get_normalizer(self._h, xm.ctypes.data_as(_c_double_p), xs.ctypes.data_as(_c_double_p),
                                                ym.ctypes.data_as(_c_double_p), ys.ctypes.data_as(_c_double_p)))
        return xm, xs, ym, ys

    def set_stream(self, stream_ptr):
        """Order every later call after the work already enqueued on this cudaStream_t (None switches it off)."""
        self._check(self.lib.gpx_set_stream(self._h, stream_ptr or None, 0 if stream_ptr is None else 1))

    @staticmethod
    def _ard(lengthscale, d, ARD):
        """ARD flag for the ABI: explicit, or inferred from the lengthscale count (d == 1 with one value: isotropic)."""
        ls = _f64(np.atleast_1d(lengthscale))
        if ARD is None:
            ARD = ls.size > 1
        if ARD and ls.size == 1 and d > 1:
            ls = np.full(d, ls[0])
        return ls, int(bool(ARD))

    # ---- host (numpy) entry points -----------------------------------------------------------------------
    def fit(self, X, Y, kernel_id, variance, lengthscale, noise, jitter=1e-8, mean_const=0.0, ARD=None):
        X, Y = _f64(X), _f64(Y)
        N, d = X.shape
        ls, ard = self._ard(lengthscale, d, ARD)
        assert Y.shape[0] == N and Y.ndim == 2
        rc = self.lib.gpx_fit(self._h, X.ctypes.data, N, d, Y.ctypes.data, Y.shape[1], int(kernel_id), float(variance),
                              ls.ctypes.data_as(_c_double_p), ls.size, ard, float(noise), float(jitter), float(mean_const),
                              MEM_HOST)
        self._check(rc)

    def append(self, X_new, Y_new):
        """Rank-k append to the resident factor (gpx_append).  Raises AppendUnsupported when the handle cannot."""
        X_new, Y_new = _f64(X_new), _f64(Y_new)
        rc = self.lib.gpx_append(self._h, X_new.ctypes.data, X_new.shape[0], Y_new.ctypes.data, MEM_HOST)
        if rc == -8:
            raise AppendUnsupported(self.lib.gpx_last_error(self._h).decode("utf-8", "replace"))
        self._check(rc)

    def predict(self, Xs, want_var=True, include_noise=True, output_dim=1, var_cols=1):
        Xs = _f64(Xs)
        Ns = Xs.shape[0]
        mean = np.empty((Ns, output_dim), dtype=np.float64)
        var = np.empty((Ns,) if var_cols == 1 else (Ns, var_cols), dtype=np.float64) if want_var else None
        rc = self.lib.gpx_predict(self._h, Xs.ctypes.data, Ns, mean.ctypes.data,
                                  var.ctypes.data if want_var else None, int(var_cols), int(bool(include_noise)), MEM_HOST)
        self._check(rc)
        return mean, var

    def predict_cov(self, Xs, include_noise=True, output_dim=1):
        Xs = _f64(Xs)
        Ns = Xs.shape[0]
        mean = np.empty((Ns, output_dim), dtype=np.float64)
        cov = np.empty((Ns, Ns), dtype=np.float64)
        rc = self.lib.gpx_predict_cov(self._h, Xs.ctypes.data, Ns, mean.ctypes.data, cov.ctypes.data,
                                      int(bool(include_noise)), MEM_HOST)
        self._check(rc)
        return mean, cov

    def predict_moments(self, Xs, order=2):
        """S0 (Ns), S1 (Ns, d), S2 (Ns, d, d) of gpx_predict_moments (S2 only for order 2)."""
        Xs = _f64(Xs)
        Ns, d = Xs.shape
        S0 = np.empty(Ns); S1 = np.empty((Ns, d)); S2 = np.empty((Ns, d, d)) if order >= 2 else None
        self._check(self.lib.gpx_predict_moments(self._h, Xs.ctypes.data, Ns, S0.ctypes.data, S1.ctypes.data,
                                                 S2.ctypes.data if S2 is not None else None, MEM_HOST))
        return S0, S1, S2

    def selfcheck(self):
        """dict(residual, factor, kv_norm, y_norm) of gpx_selfcheck (collective on a sharded handle)."""
        out = np.zeros(4)
        self._check(self.lib.gpx_selfcheck(self._h, out.ctypes.data_as(_c_double_p), 4))
        return {"solve_residual": float(out[0]), "factor_residual": float(out[1]), "kv_norm": float(out[2]), "y_norm": float(out[3])}

    # ---- device-pointer entry points (inputs already resident in HBM, e.g. torch tensors) -------------------
    def fit_device(self, X_ptr, N, d, Y_ptr, O, kernel_id, variance, lengthscale, noise, jitter=1e-8, mean_const=0.0,
                   ARD=None):
        ls, ard = self._ard(lengthscale, d, ARD)
        rc = self.lib.gpx_fit(self._h, X_ptr, int(N), int(d), Y_ptr, int(O), int(kernel_id), float(variance),
                              ls.ctypes.data_as(_c_double_p), ls.size, ard, float(noise), float(jitter), float(mean_const),
                              MEM_DEVICE)
        self._check(rc)

    def predict_device(self, Xs_ptr, Ns, mean_ptr, var_ptr, include_noise=True, var_cols=1):
        rc = self.lib.gpx_predict(self._h, Xs_ptr, int(Ns), mean_ptr, var_ptr, int(var_cols), int(bool(include_noise)),
                                  MEM_DEVICE)
        self._check(rc)

    # ---- results ------------------------------------------------------------------------------------------
    def lml(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.gpx_lml(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def lml_grad(self, n_ls):
        """d LML / d [variance, lengthscale..., noise, mean_const] of the last fit."""
        g = np.zeros(n_ls + 3, dtype=np.float64)
        rc = self.lib.gpx_lml_grad(self._h, g.ctypes.data_as(_c_double_p), g.size)
        if rc < 0:
            self._check(rc)
        return g

    def export(self, N, O, want_L=False):
        alpha = np.empty((N, O), dtype=np.float64)
        L = np.empty((N, N), dtype=np.float64) if want_L else None
        self._check(self.lib.gpx_export(self._h, alpha.ctypes.data, L.ctypes.data if want_L else None))
        return alpha, L

    def kernel_matrix(self, X, X2, kernel_id, variance, lengthscale, diag_add=0.0, ARD=None):
        X = _f64(X)
        N, d = X.shape
        ls, ard = self._ard(lengthscale, d, ARD)
        if X2 is None:
            M = N
            out = np.empty((N, N), dtype=np.float64)
            x2p = None
        else:
            X2 = _f64(X2)
            M = X2.shape[0]
            out = np.empty((N, M), dtype=np.float64)
            x2p = X2.ctypes.data
        rc = self.lib.gpx_kernel_matrix(self._h, X.ctypes.data, N, d, x2p, M, int(kernel_id), float(variance),
                                        ls.ctypes.data_as(_c_double_p), ls.size, ard, float(diag_add), out.ctypes.data)
        self._check(rc)
        return out

    def phase_ms(self):
        return {name: self.lib.gpx_phase_ms(self._h, i) for i, name in enumerate(PHASES)}

    def stream_ptr(self):
        return int(self.lib.gpx_stream(self._h) or 0)

    def gemm_stats(self):
        """(device ms, issued fp64 flops, launches) of the DMMA tile-GEMM kernel since the last call."""
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        self._check(self.lib.gpx_gemm_stats(self._h, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)))
        return ms.value, fl.value, n.value

    def trace(self):
        """(npanels, 6) array of per-panel timeline stamps (ms) of the last fit; needs set_option("trace", 1)."""
        n = int(self.lib.gpx_trace(self._h, None, 0))
        out = np.zeros(max(n, 0), dtype=np.float64)
        if n > 0:
            self.lib.gpx_trace(self._h, out.ctypes.data_as(_c_double_p), n)
        return out.reshape(-1, 6)

    def launch_count(self):
        return int(self.lib.gpx_launch_count(self._h))

    def bytes_allocated(self):
        return int(self.lib.gpx_bytes_allocated(self._h))
