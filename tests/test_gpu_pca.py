"""-m gpu: PCA projection of the latents against sklearn's own PCA.transform -- the call
/root/reference/run_dim_reduction.py:86 makes -- on the same fitted model and inputs.  Tolerance 1e-4 of max|out|."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fit(n, l, whiten, n_components, seed):
    from sklearn.decomposition import PCA
    rng = np.random.RandomState(seed)
    basis = rng.randn(24, l).astype(np.float32)
    x = (rng.randn(n, 24).astype(np.float32) * np.linspace(3, 0.2, 24, dtype=np.float32)) @ basis
    x += 0.05 * rng.randn(n, l).astype(np.float32) + rng.randn(l).astype(np.float32)
    pca = PCA(n_components, svd_solver='auto', whiten=whiten)
    pca.fit(x)
    return pca, x


@pytest.mark.parametrize("n,l,whiten,n_components", [(300, 4096, False, 0.5), (129, 4096, True, 7), (1000, 512, False, 100),
                                                     (5, 64, False, 3)])
def test_pca_transform_matches_sklearn(n, l, whiten, n_components):
    from dynamorph_b200.run_dim_reduction import pca_transform
    pca, x = _fit(max(n, 200), l, whiten, n_components, seed=n + l)
    x = x[:n]
    ref = pca.transform(x)
    for chunk in (64, 65536):
        got = pca_transform(pca, x, chunk=chunk)
        assert got.shape == ref.shape and got.dtype == pca.components_.dtype == ref.dtype
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        assert err < 1e-4, (chunk, err)
    with pytest.raises(ValueError):
        pca_transform(pca, x[:, :-4])


def test_process_pca_roundtrip(tmp_path):
    """pickles in -> pickle out, file names as run_dim_reduction.py:83-84 builds them."""
    from dynamorph_b200.run_dim_reduction import process_PCA
    pca, x = _fit(220, 4096, False, 0.5, seed=3)
    wdir, idir, odir = tmp_path / "w", tmp_path / "in", tmp_path / "out"
    os.makedirs(wdir); os.makedirs(idir)
    pickle.dump(pca, open(wdir / "pca_model.pkl", "wb"), protocol=4)
    pickle.dump(x, open(idir / "C5_latent_space_after.pkl", "wb"), protocol=4)
    process_PCA(str(idir), str(odir), str(wdir), "C5", suffix="after")
    got = pickle.load(open(odir / "C5_latent_space_after_PCAed.pkl", "rb"))
    ref = pca.transform(x)
    assert float(np.abs(got - ref).max() / np.abs(ref).max()) < 1e-4
    with pytest.raises(ValueError, match="PCA weights"):
        process_PCA(str(idir), str(odir), str(tmp_path / "missing"), "C5", suffix="after")
