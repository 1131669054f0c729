"""Host-side logic of the drop-in SpectrumModel/Emulator (no GPU): parameter bookkeeping mirrors the
reference's tests (tests/test_models/test_models.py) and the upstream transforms reproduce the stage
inputs the reference produced (fixtures captured from inside the reference's __call__)."""
import os

import numpy as np
import pytest

from starfish_b200 import synth
from starfish_b200.paramtree import ParamTree

from _helpers import make_model


@pytest.fixture
def model():
    return make_model(256, 0, wave=synth.log_uniform_wave(256, 5092.0, 5108.0), mus=(5098.0, 5103.0))


def test_paramtree_semantics():
    t = ParamTree({"a": 1, "g": {"x": 2, "y": 3}, "l": [{"mu": 1.0}, {"mu": 2.0}]})
    assert t.keys() == ["a", "g:x", "g:y", "l:0:mu", "l:1:mu"]
    assert t["g:x"] == 2 and t["g"]["y"] == 3 and t["l"]["1"]["mu"] == 2.0
    assert "g" in t and "g:x" in t and "g:z" not in t and "mu" in t["l"]["0"]
    t["g:z"] = 5
    t["new:deep:key"] = 7
    assert t.as_dict()["l"] == [{"mu": 1.0}, {"mu": 2.0}]
    assert t.as_dict()["new"] == {"deep": {"key": 7}}
    del t["g"]
    assert "g:x" not in t and t == {"a": 1, "l": [{"mu": 1.0}, {"mu": 2.0}], "new": {"deep": {"key": 7}}}
    with pytest.raises(KeyError):
        t["nope"]


def test_labels_order_and_vector_roundtrip(model):
    assert model.labels == ("vsini", "vz", "log_scale", "global_cov:log_amp", "global_cov:log_ls",
                            "local_cov:0:mu", "local_cov:0:log_amp", "local_cov:0:log_sigma",
                            "local_cov:1:mu", "local_cov:1:log_amp", "local_cov:1:log_sigma",
                            "cheb:1", "cheb:2", "T", "logg", "Z")
    P0 = model.get_param_vector()
    model.set_param_vector(P0)
    assert np.array_equal(P0, model.get_param_vector())
    P0[2] = 7
    model.set_param_vector(P0)
    assert model[model.labels[2]] == 7
    with pytest.raises(ValueError):
        model.set_param_vector(np.append(P0, 1.0))


def test_item_access_errors(model):
    for bad in ("garbage", "global_cov:not quite", "global_cov:garbage", "local_cov:garbage"):
        with pytest.raises(KeyError):
            model[bad] = -4
    with pytest.raises(KeyError):
        model["cheb:0"] = 1
    model["cheb"] = [-0.2, 0.1]
    assert model["cheb:1"] == -0.2 and model["cheb:2"] == 0.1
    model["cheb:4"] = 0.05
    assert list(model.cheb) == [-0.2, 0.1, 0, 0.05]


def test_freeze_thaw(model):
    model.freeze("logg")
    assert "logg" not in model.get_param_dict() and "T" in model.get_param_dict()
    P = model.get_param_dict()
    model.freeze("Z")
    P["Z"] = 7
    model.set_param_dict(P)
    assert model["Z"] != 7
    model.thaw(["logg", "Z"])
    labels = model.labels
    model.freeze("all")
    assert set(labels + ("global_cov", "local_cov", "cheb")) == set(model.frozen)
    model.thaw("all")
    assert set(labels) == set(model.labels)
    for group in ("global_cov", "local_cov", "cheb"):
        members = [l for l in model.labels if l.startswith(group)]
        model.freeze(group)
        assert group in model.frozen and all(m in model.frozen for m in members)
        model.thaw(group)
        assert group not in model.frozen and not any(m in model.frozen for m in members)
    before = list(model.frozen)
    model.freeze("pinguino")
    model.thaw("pinguino")
    assert model.frozen == before


def test_delete_group_clears_cache_and_frozen(model):
    model.freeze("global_cov")
    model._kernel_hyper()
    assert model._glob_cov is not None
    del model["global_cov"]
    assert "global_cov" not in model.params and "global_cov" not in model.frozen and model._glob_cov is None
    with pytest.raises(KeyError):
        del model["global_cov"]


def test_frozen_kernel_cache_semantics(model):
    """spectrum_model.py:341-363: thawed groups are re-read every call; frozen groups keep the cached kernel;
    freeze() itself clears the cache so the next call recomputes once."""
    assert model._glob_cov is None and model._loc_cov is None
    g0, l0 = model._kernel_hyper()
    assert model._glob_cov.shape == model._loc_cov.shape == (256, 256)
    model["global_cov:log_amp"] = model["global_cov:log_amp"] + 1.0
    g1, _ = model._kernel_hyper()
    assert np.isclose(g1[0], g0[0] * np.e)
    model.freeze("global_cov")
    assert model._glob_cov is None
    g2, _ = model._kernel_hyper()
    model.params["global_cov:log_amp"] = 0.0      # edited behind the freeze: cached kernel still used
    g3, _ = model._kernel_hyper()
    assert g3 == g2 == g1


def test_save_load_roundtrip(model, tmp_path):
    path = os.path.join(tmp_path, "model.toml")
    model.freeze(["logg", "vsini", "global_cov"])
    model.set_param_vector(model.get_param_vector())  # numpy scalars must survive TOML
    P0, f0 = model.params.as_dict(), list(model.frozen)
    model.save(path, metadata={"note": "x"})
    assert "[metadata]\n" in open(path).readlines()
    model.load(path)
    assert model.params == P0 and model.frozen == f0


def test_multi_order_rejected():
    from starfish_b200.emulator import Emulator
    from starfish_b200.spectrum import Spectrum
    from starfish_b200.spectrum_model import SpectrumModel

    emu = Emulator(**synth.make_emulator_arrays())
    w, f, s = synth.make_data(256)
    two = Spectrum(w.reshape(2, -1), f.reshape(2, -1), s.reshape(2, -1))
    with pytest.raises(ValueError):
        SpectrumModel(emu, two, grid_params=[6100, 4.5, 0.0])


def test_train_validates_priors(model):
    import scipy.stats as st

    with pytest.raises(ValueError):
        model.train({"penguin": st.uniform(5900, 6700)}, options={"maxiter": 1})
    with pytest.raises(ValueError):
        model.train({"T": lambda x: 1 / x}, options={"maxiter": 1})


@pytest.mark.parametrize("walker", [0, 1])
def test_transform_mirrors_reproduce_reference_stage_inputs(golden_dir, walker):
    """The public numpy mirrors of Starfish/transforms.py + Emulator.__call__, chained as spectrum_model.py:287-332
    chains them, reproduce the stage inputs recorded inside the reference (the model itself runs these stages on
    the device; this pins the host-side API functions)."""
    from starfish_b200 import transforms as T

    g = dict(np.load(os.path.join(golden_dir, f"model_n256_w{walker}.npz")))
    m = make_model(256, walker, wave=g["wave"], mus=(5098.0, 5103.0))
    assert tuple(g["labels"]) == m.labels
    assert np.array_equal(g["param_vector"], m.get_param_vector())
    fluxes = T.rotational_broaden(m.min_dv_wave, m.bulk_fluxes, m["vsini"])
    fluxes = T.resample(T.doppler_shift(m.min_dv_wave, m["vz"]), fluxes, m.data.wave)
    fluxes = T.chebyshev_correct(m.data.wave, fluxes, [1, *m.cheb])
    weights, wcov = m.emulator(m.grid_params)
    *eig, mean, std = fluxes
    X = T.rescale(eig * std, np.exp(m["log_scale"]))
    flux = T.rescale(weights @ (eig * std) + mean, np.exp(m["log_scale"]))
    assert np.abs(flux - g["model_flux"]).max() <= 1e-13
    assert np.abs(X - g["X"]).max() <= 1e-15
    # Σ_w = v22 − v21·v11⁻¹·v12 cancels 1e4 down to O(1): its fp64 noise floor is ~1e-11 relative and
    # depends on the LAPACK build (numpy's dgesv vs scipy's getrf/getrs differ at 4e-12), see DESIGN.md
    assert np.abs(wcov - g["weights_cov"]).max() <= 1e-9 * np.abs(g["weights_cov"]).max()
    glob, loc = m._kernel_hyper()
    assert np.allclose(glob, g["glob"], rtol=1e-15) and np.allclose(loc, g["loc"], rtol=1e-15)


def test_emulator_call_contract():
    from starfish_b200.emulator import Emulator

    emu = Emulator(**synth.make_emulator_arrays())
    with pytest.warns(UserWarning):
        mu, cov = emu([6100, 4.5, 0.0])
    assert mu.shape == (6,) and cov.shape == (6, 6) and np.allclose(cov, cov.T)
    emu._trained = True
    with pytest.raises(ValueError):
        emu([5000, 4.5, 0.0])
    with pytest.raises(ValueError):
        emu([[6100, 4.5, 0.0]], full_cov=True, reinterpret_batch=True)
    mu_b, cov_b = emu.predict_batch([[6100, 4.5, 0.0], [9000, 4.5, 0.0]])
    assert np.allclose(mu_b[0], mu) and np.isnan(mu_b[1]).all()
    # at a grid point the GP mean reproduces ŵ closely (λ_ξ = 1 regularisation keeps it approximate)
    assert emu.bulk_fluxes.shape == (8, emu.wl.size)
    assert np.isfinite(emu.log_likelihood())
    P = emu.get_param_vector()
    emu.set_param_vector(P)
    assert np.allclose(emu.get_param_vector(), P)


def test_emulator_train_improves_likelihood_and_respects_bounds():
    """Emulator.train mirrors emulator.py:484-524: Nelder-Mead on the hyper-parameters, lengthscales kept
    above twice the grid separation, state restored when the optimiser does not converge."""
    import copy

    from starfish_b200 import synth
    from starfish_b200.emulator import Emulator

    emu = Emulator(**copy.deepcopy(synth.make_emulator_arrays(n_comp=2)))
    before = emu.log_likelihood()
    soln = emu.train(options={"maxiter": 60})
    if soln.success:
        assert emu._trained and emu.log_likelihood() >= before
    else:
        assert abs(emu.log_likelihood() - before) <= 1e-9 * abs(before)
    assert np.all(emu.lengthscales >= 2 * emu._grid_sep - 1e-12)


def test_batch_parameter_packing_is_pure_host_logic():
    """theta / hyper-parameter packing of log_likelihood_batch (no GPU involved): column order, frozen values,
    frozen-group sharing, vectorised priors."""
    import scipy.stats as st

    from _helpers import make_model

    m = make_model(256, 0, wave=np.linspace(5092.0, 5108.0, 256), mus=(5098.0, 5103.0), vz=3.0)
    labels = list(m.labels)
    P0 = m.get_param_vector()
    P = np.tile(P0, (3, 1))
    P[1, labels.index("vsini")] = 7.5
    P[2, labels.index("cheb:2")] = 0.03
    P[2, labels.index("local_cov:1:log_sigma")] += 0.25
    B, cols = m._columns(P)
    th = m._theta(B, cols)
    D = 3
    assert th.shape == (3, D + 4 + 2)
    assert np.array_equal(th[:, :D], np.tile(m.grid_params, (3, 1)))
    assert th[:, D].tolist() == [5.0, 7.5, 5.0] and np.all(th[:, D + 1] == 3.0) and np.all(th[:, D + 3] == 1.0)
    assert th[:, D + 4].tolist() == [0.01] * 3 and th[:, D + 5].tolist() == [-0.01, -0.01, 0.03]
    glob, nloc, loc, shared = m._hyper_rows(B, cols)
    assert not shared and glob.shape == (3, 2) and nloc.tolist() == [2, 2, 2]
    assert np.allclose(loc[2, 1, 2], np.exp(cols["local_cov:1:log_sigma"][2])) and loc[0, 1, 2] != loc[2, 1, 2]
    assert np.allclose(glob[:, 0], np.exp(m["global_cov:log_amp"]))
    # frozen parameters keep the model's value whatever P holds for the thawed ones
    m.freeze("vz")
    labels2 = list(m.labels)
    assert "vz" not in labels2
    B2, cols2 = m._columns(np.tile(m.get_param_vector(), (2, 1)))
    assert np.all(cols2["vz"] == 3.0)
    # frozen kernel groups: one shared hyper-parameter row
    m.freeze(["global_cov", "local_cov"])
    B3, cols3 = m._columns(np.tile(m.get_param_vector(), (4, 1)))
    g3, n3, l3, shared3 = m._hyper_rows(B3, cols3)
    assert shared3 and g3.shape == (1, 2) and l3.shape[0] == 1 and n3.tolist() == [2]
    # priors are evaluated on whole columns; scalar-only prior objects fall back to per-row calls
    class ScalarOnly:
        def logpdf(self, v):
            if np.ndim(v):
                raise TypeError("scalar only")
            return -0.5 * (v - 5.0) ** 2

    m.thaw("all")
    B4, cols4 = m._columns(P)
    lp = m._prior_rows(B4, cols4, {"vsini": ScalarOnly(), "T": st.uniform(6000, 200), "nope": st.norm()})
    assert np.allclose(lp, [-0.5 * 0.0 + st.uniform(6000, 200).logpdf(m["T"]),
                            -0.5 * 2.5 ** 2 + st.uniform(6000, 200).logpdf(m["T"]),
                            st.uniform(6000, 200).logpdf(m["T"])])
    with np.testing.assert_raises(ValueError):
        m._check_transforms({"vsini": np.array([5.0, 0.0]), **{k: v for k, v in cols4.items() if k != "vsini"}})
