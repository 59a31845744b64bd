"""Metrics half of the reference's src/visualization/plot_helper.py.  Figures are outside the
accelerated path (and matplotlib is not a dependency here): `helper_plot` is a no-op hook."""
import numpy as np


def calculate_differential_entropy(std_devs: np.ndarray) -> float:
    """plot_helper.py:383-398: sum of Gaussian differential entropies, in bits; zero std -> 1e-6
    (in place, like the reference)."""
    if np.any(std_devs == 0):
        std_devs[std_devs == 0] = 1e-6
    return np.sum(np.log(std_devs * np.sqrt(2 * (np.pi * np.e)))) / (np.log(2))


def helper_plot(*args, **kwargs):
    return None
