"""CPU tests: the numpy/scipy restatement of the reference's feature producer (oracle/features_oracle.py) against the
goldens generated from the unmodified reference (tests/golden/features_golden.npz, oracle/make_golden_features.py),
incl. the two 25-vectors stored in the reference's tests/models_tests/FeaturesTests.ipynb."""
import os

import numpy as np
import pytest

from oracle import features_oracle as fo


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "features_golden.npz"))


def test_frame_average(gold):
    assert np.allclose(fo.average_frames(gold["traj"], 10), gold["avg"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("key,dt", [("feat_dt1", 1.0), ("feat_dt01", 0.1)])
def test_features_match_reference(gold, key, dt):
    got = fo.features_batch(gold["avg"], dt)
    assert got.shape == gold[key].shape == (18, 25)
    assert np.allclose(got, gold[key], rtol=1e-12, atol=1e-15, equal_nan=True)


def test_notebook_vectors(gold):
    got = fo.features_batch(gold["long"], 0.1)
    assert np.allclose(got, gold["long_feat_dt01"], rtol=1e-12, atol=1e-15)
    assert np.allclose(got, gold["notebook"], rtol=2e-6, atol=1e-9)     # 9 printed digits in the notebook


def test_reference_import_agrees_when_present(gold):
    from oracle import refshim
    if not refshim.reference_available():
        pytest.skip("reference checkout not present on this machine")
    refshim.import_reference()
    import importlib
    hf = importlib.import_module("helpers.helpersFeatures")
    ref = np.stack([hf.compute_diffusion_features(t, dt=0.1) for t in gold["avg"][:4]])
    assert np.allclose(fo.features_batch(gold["avg"][:4], 0.1), ref, rtol=1e-12, atol=1e-15)
