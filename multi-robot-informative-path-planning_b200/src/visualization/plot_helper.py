"""Metrics half of the reference's src/visualization/plot_helper.py.  Figures are outside the
accelerated path (and matplotlib is not a dependency here): `helper_plot` is a no-op hook."""
import numpy as np


def calculate_differential_entropy(std_devs: np.ndarray) -> float:
    """plot_helper.py:383-398: sum of Gaussian differential entropies, in bits; zero std -> 1e-6
    (in place, like the reference)."""
    if np.any(std_devs == 0):
        std_devs[std_devs == 0] = 1e-6
    return np.sum(np.log(std_devs * np.sqrt(2 * (np.pi * np.e)))) / (np.log(2))


def field_metrics_device(z_pred, std, z_true):
    """(RMSE, weighted RMSE of log10(Z + 1), differential entropy) of main.py:113-116 in one fused pass over
    device-resident tensors (ipp_grid_metrics): a metrics-only caller never reads the posterior back."""
    from mripp_b200 import _lib

    torch = _lib.require_cuda()
    lib = _lib.load()
    z, s, t = (a.contiguous().view(-1) for a in (z_pred, std, z_true))
    G = z.numel()
    partial = torch.empty(4 * 592, dtype=torch.float64, device="cuda")
    out = torch.empty(4, dtype=torch.float64, device="cuda")
    _lib.check(lib.ipp_grid_metrics(_lib.ptr(z), _lib.ptr(s), _lib.ptr(t), G, _lib.ptr(partial), _lib.ptr(out),
                                    _lib.stream_ptr()), "ipp_grid_metrics")
    r = out.cpu().numpy()
    return float(r[0]), float(r[1]), float(r[2])


def helper_plot(*args, **kwargs):
    return None
