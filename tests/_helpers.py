import copy
import warnings

from starfish_b200 import synth


def make_model(n_pix=256, walker=0, wave=None, mus=None, **over):
    """Our SpectrumModel on the seeded synthetic set-up (same arrays the golden generator fed the reference)."""
    from starfish_b200.emulator import Emulator
    from starfish_b200.spectrum import Spectrum
    from starfish_b200.spectrum_model import SpectrumModel

    emu = Emulator(**copy.deepcopy(synth.make_emulator_arrays()))
    emu._trained = True
    w, f, s = synth.make_data(n_pix, wave=wave)
    grid, p = synth.walker_params(walker)
    if mus is not None:
        for k, mu in enumerate(mus):
            p["local_cov"][k]["mu"] = float(mu)
    p.update(over)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return SpectrumModel(emu, Spectrum(w, f, sigmas=s, name="synthetic"), grid_params=grid, **p)


def make_model_params(wave, grid, params, **kw):
    """Our SpectrumModel for an explicit (wave, grid params, parameter dict) — the upstream variants."""
    from starfish_b200.emulator import Emulator
    from starfish_b200.spectrum import Spectrum
    from starfish_b200.spectrum_model import SpectrumModel

    emu = Emulator(**copy.deepcopy(synth.make_emulator_arrays()))
    emu._trained = True
    w, f, s = synth.make_data(len(wave), wave=wave)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return SpectrumModel(emu, Spectrum(w, f, sigmas=s, name="synthetic"), grid_params=grid,
                             **copy.deepcopy(params), **kw)


def ls_for_half_bandwidth(wave, b):
    """Length scale (km/s) of the global kernel for which the covariance band has half-width exactly b pixels on
    this (strictly increasing) grid: the support test is r <= 6·ls with r = (c/2)|λi−λk|/(λi+λk)
    (Starfish/models/kernels.py:27-33), so 6·ls is put midway between the smallest r at offsets b and b+1."""
    import numpy as np

    c = 2.99792458e5

    def rmin(d):
        return float(np.min(c / 2 * np.abs((wave[:-d] - wave[d:]) / (wave[:-d] + wave[d:]))))

    return (rmin(b) + rmin(b + 1)) / 12.0
