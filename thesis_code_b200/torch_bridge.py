"""Thin PyTorch bridge over the C ABI: CUDA tensors in, CUDA tensors out, ordered after the caller's current stream.

``north_star`` names "a thin C-ABI PyTorch extension" as the only bridge; PyTorch is plumbing here (device memory, streams),
so the bridge is deliberately nothing but pointer unwrapping: it checks dtype / contiguity / device, hands
``torch.cuda.current_stream()`` to ``gpx_set_stream`` (the library then waits on the device for whatever the caller has
enqueued there -- no host synchronisation) and calls the same ``gpx_fit`` / ``gpx_predict`` entry points with
``GPX_MEM_DEVICE``.  The calls return after the handle's streams have drained, so the results are visible to any stream.
"""
import torch

from . import _gpx


def _check(t, name, device_index, ndim):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError("{} must be a CUDA tensor".format(name))
    if t.dtype != torch.float64 or not t.is_contiguous() or t.dim() != ndim:
        raise ValueError("{} must be a contiguous float64 tensor with {} dimensions".format(name, ndim))
    if t.device.index != device_index:
        raise ValueError("{} lives on cuda:{} but the handle is bound to cuda:{}".format(name, t.device.index, device_index))


def _order_after_current_stream(handle):
    handle.set_stream(torch.cuda.current_stream(torch.device("cuda", handle.device)).cuda_stream)


def fit(handle, X, Y, kernel_id, variance, lengthscale, noise, jitter=1e-8, mean_const=0.0, ARD=None):
    """gpx_fit on device tensors X (N, d), Y (N, O)."""
    _check(X, "X", handle.device, 2)
    _check(Y, "Y", handle.device, 2)
    if X.shape[0] != Y.shape[0]:
        raise ValueError("X and Y has to match size. It was {} and {} respectively".format(X.shape[0], Y.shape[0]))
    _order_after_current_stream(handle)
    handle.fit_device(X.data_ptr(), X.shape[0], X.shape[1], Y.data_ptr(), Y.shape[1], kernel_id, variance, lengthscale, noise,
                      jitter=jitter, mean_const=mean_const, ARD=ARD)


def predict(handle, Xs, output_dim=1, want_var=True, include_noise=True, var_cols=1, out=None):
    """gpx_predict on a device tensor Xs (N*, d) -> (mean (N*, O), var (N*,) | (N*, O) | None) device tensors."""
    _check(Xs, "Xs", handle.device, 2)
    Ns = Xs.shape[0]
    if out is not None:
        mean, var = out
    else:
        mean = torch.empty((Ns, output_dim), dtype=torch.float64, device=Xs.device)
        var = torch.empty((Ns,) if var_cols == 1 else (Ns, var_cols), dtype=torch.float64, device=Xs.device) if want_var else None
    _order_after_current_stream(handle)
    handle.predict_device(Xs.data_ptr(), Ns, mean.data_ptr(), var.data_ptr() if want_var else None, include_noise, var_cols=var_cols)
    return mean, var
