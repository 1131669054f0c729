"""starfish_b200 — B200-native (sm_100a) implementation of Starfish's per-step log-likelihood hot path.

Public surface (mirrors the reference names on the path):

* :class:`starfish_b200.engine.LikelihoodEngine` — the batched stage boundary over the C ABI
* :func:`starfish_b200.kernels.global_covariance_matrix`, :func:`local_covariance_matrix` — function seam
* :class:`starfish_b200.spectrum_model.SpectrumModel` — drop-in model object (``log_likelihood``,
  ``log_likelihood_batch``)
* :class:`starfish_b200.emulator.Emulator`, :class:`starfish_b200.spectrum.Spectrum` — host-side producers
  of the path's inputs

Importing the package does not load CUDA; the shared library is loaded on first use and its absence is
an error (no CPU fallback).
"""
__version__ = "0.1.0"

from .constants import c_kms, JITTER  # noqa: F401
