"""B200-native GP estimation + information-gain scoring for multi-robot informative path planning.

Drop-in for the hot path of franfrancisco9/Multi-Robot-Informative-Path-Planning: the estimator and
strategy classes keep the reference's Python API (``src.point_source.point_source.PointSourceField``,
``src.rrt.rrt.*``, ``src.boustrophedon.boustrophedon.*``, ``main``) while the dense math runs in
hand-written sm_100a CUDA kernels behind the C ABI declared in ``include/ipp_b200.h``.
There is no CPU fallback: GPU entry points raise if the extension or a CUDA device is missing.
"""
__version__ = "0.1.0"
