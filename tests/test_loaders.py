"""Product-side loaders (no GPU): checkpoint unpickling without NFF, random init, slab pickles."""
import io
import pickle
import sys

import numpy as np
import pytest
import torch

from surface_sampling_b200 import loaders


def test_init_random_weights_matches_the_oracle_init():
    """bench.py's B200 arm and its CPU legs must run the SAME random-init weights."""
    from oracle.painn import init_random_weights as oracle_init
    for seed in (0, 2):
        a, b = loaders.init_random_weights(seed), oracle_init(seed)
        assert a.keys() == b.keys()
        assert all(np.array_equal(a[k], b[k]) for k in a)
    assert sum(v.size for v in a.values()) == 589057


def test_state_dict_validation():
    sd = loaders.init_random_weights(1)
    assert loaders.as_state_dict(sd).keys() == sd.keys()
    bad = dict(sd)
    bad["update_blocks.0.u_mat.weight"] = np.zeros((64, 64), np.float32)
    with pytest.raises(ValueError):
        loaders.as_state_dict(bad)
    del bad["update_blocks.0.u_mat.weight"]
    with pytest.raises(KeyError):
        loaders.as_state_dict(bad)
    with pytest.raises(NotImplementedError):
        loaders.load_model("x", model_type="CHGNetNFF")


def test_load_model_reads_an_nff_style_pickle_without_nff(tmp_path):
    """A `best_model` is a pickled nn.Module whose classes live under nff.*: build one with throw-away stand-ins,
    drop the stand-ins, and load it back through the product loader."""
    sd = loaders.init_random_weights(3)
    with loaders._StubModules():
        import nff.nn.models.painn as mp
        model = mp.Painn()
        for k, v in sd.items():      # nested parameter names -> register flat (state_dict keys are what matters)
            model.register_buffer(k.replace(".", "__"), torch.from_numpy(v))
        model.state_dict = None      # instance attr would not survive the pickle; see below
        del model.state_dict
        torch.save(model, tmp_path / "best_model")
    assert "nff" not in sys.modules
    # the stand-in stores flat names; patch check to map them back
    orig = loaders.check_state_dict
    try:
        loaders.check_state_dict = lambda d: orig({k.replace("__", "."): v for k, v in d.items()})
        got = loaders.load_model(tmp_path, model_type="PaiNN")        # folder form
    finally:
        loaders.check_state_dict = orig
    assert all(np.array_equal(got[k], sd[k]) for k in sd)
    assert "nff" not in sys.modules                                   # stubs are removed again


def test_golden_weights_pass_validation(sto_weights):
    for sd in sto_weights:
        assert loaders.as_state_dict(sd).keys() == loaders.expected_shapes().keys()


def test_load_slab_pickle_roundtrip(tmp_path, structures):
    """numpy-only pickle with an ASE-like object graph (arrays / _cellobj / _pbc / _constraints)."""
    s = structures["GaN_0001_3x3"]

    class Cell:  # noqa
        pass

    class Fix:  # noqa
        pass

    class Gratoms:  # noqa
        pass

    for c in (Cell, Fix, Gratoms):
        c.__module__, c.__qualname__ = "fake_ase_mod", c.__name__
    mod = type(sys)("fake_ase_mod")
    mod.Cell, mod.Fix, mod.Gratoms = Cell, Fix, Gratoms
    sys.modules["fake_ase_mod"] = mod
    try:
        g, c, f = Gratoms(), Cell(), Fix()
        c.array = s["cell"]
        f.index = np.arange(36)
        g.arrays = {"numbers": s["numbers"], "positions": s["positions"], "tags": np.ones(36, int)}
        g._cellobj, g._pbc, g._constraints = c, s["pbc"], [f]
        (tmp_path / "slab.pkl").write_bytes(pickle.dumps(g))
    finally:
        del sys.modules["fake_ase_mod"]
    a = loaders.load_slab_pickle(tmp_path / "slab.pkl")
    assert np.array_equal(a.get_atomic_numbers(), s["numbers"]) and np.allclose(a.get_positions(), s["positions"])
    assert a.fixed_mask().all() and np.array_equal(a.get_array("tags"), np.ones(36, int))
