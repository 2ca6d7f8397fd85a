"""
Histogrammer / FieldHistogrammer: the known-answer tests of
test/test_histogram.py, on host tensors through the decomposition's test hook (the same
tensor operations run on the GPU; tests/test_gpu_generated.py).
"""
from itertools import product

import numpy as np
import pytest
import torch

import oracle as O
import pystella_b200 as ps
from helpers import HostOps
from pystella_b200.field import var


def host_decomp(h, grid_shape):
    decomp = ps.DomainDecomposition((1, 1, 1), h, grid_shape=grid_shape)
    decomp.host_ops = HostOps()
    return decomp


@pytest.mark.parametrize("num_bins", [123, 1024])
def test_trivial_histogram(num_bins):
    """test/test_histogram.py:35-62: constant bin and weight expressions."""
    grid_shape = (24, 20, 16)
    decomp = host_decomp(1, grid_shape)
    histograms = {"a": (13., 1), "b": (10.3, 2), "c": (100.9, 3)}
    hist = ps.Histogrammer(decomp, histograms, num_bins, "float64", rank_shape=grid_shape)
    result = hist(None)
    for key, (_b, weight) in histograms.items():
        res = result[key]
        b = int(np.floor(_b))
        assert res[b] == weight * np.prod(grid_shape)
        assert np.all(res[res != res[b]] == 0.)


@pytest.mark.parametrize("dtype, max_rtol, avg_rtol", [("float64", 1e-10, 1e-11),
                                                       ("float32", 5e-4, 5e-5)])
@pytest.mark.parametrize("num_bins", [123, 1024, 1493])
def test_histogram(dtype, max_rtol, avg_rtol, num_bins):
    """test/test_histogram.py:65-121 against numpy.histogram."""
    rng = np.random.default_rng(12321)
    grid_shape = (32, 24, 16)
    decomp = host_decomp(1, grid_shape)
    _fx = ps.Field("fx")
    histograms = {"count": (var("abs")(_fx) * num_bins, 1),
                  "squared": (var("abs")(_fx) * num_bins, _fx**2)}
    hist = ps.Histogrammer(decomp, histograms, num_bins, dtype, rank_shape=grid_shape)
    fx_h = rng.random(grid_shape).astype(dtype)
    result = hist(None, fx=torch.from_numpy(fx_h))
    res = result["count"]
    assert res.dtype == np.dtype(dtype)
    assert np.sum(res.astype("int64")) == np.prod(grid_shape)
    # bin edges as the kernel sees them: floor(|f| * num_bins) in the field's precision
    idx = np.floor(np.abs(fx_h) * np.asarray(num_bins, dtype)).astype(int)
    np_res = np.bincount(idx.reshape(-1), minlength=num_bins)[:num_bins]
    assert np.array_equal(res, np_res.astype(dtype))
    bins = np.linspace(0, 1, num_bins + 1).astype(dtype)
    np_hist = np.histogram(fx_h, bins=bins)[0]
    assert np.sum(np.abs(res - np_hist)) <= (4 if dtype == "float64" else 64)  # edge sites
    res = result["squared"]
    np_res = np.bincount(idx.reshape(-1), weights=fx_h.astype(np.float64).reshape(-1)**2,
                         minlength=num_bins)[:num_bins]
    max_err, avg_err = O.get_errs(res, np_res)
    assert max_err < max_rtol and avg_err < avg_rtol


@pytest.mark.parametrize("dtype, max_rtol, avg_rtol", [("float64", 1e-10, 1e-11),
                                                       ("float32", 5e-4, 5e-5)])
def test_field_histogram(dtype, max_rtol, avg_rtol):
    """test/test_histogram.py:124-181."""
    rng = np.random.default_rng(12321)
    grid_shape, h, num_bins = (24, 16, 20), 1, 432
    decomp = host_decomp(h, grid_shape)
    pencil = tuple(n + 2 * h for n in grid_shape)
    hist = ps.FieldHistogrammer(decomp, num_bins, dtype, rank_shape=grid_shape, halo_shape=h)
    fx = (rng.random((2, 2) + pencil) * 4.2 - 1.2).astype(dtype)
    fx_h = fx[..., h:-h, h:-h, h:-h]
    result = hist(torch.from_numpy(fx))
    for slc in product(range(2), range(2)):
        res = result["linear"][slc]
        assert res.sum() == np.prod(grid_shape)
        np_res = np.histogram(fx_h[slc].astype(np.float64), bins=result["linear_bins"][slc])[0]
        # sites within rounding of a bin edge may land in the neighbouring bin
        assert np.sum(np.abs(res - np_res)) <= (4 if dtype == "float64" else 400)
        assert O.get_errs(np.cumsum(res)[8::8], np.cumsum(np_res)[8::8])[0] < 1e-2
        res = result["log"][slc]
        bins = result["log_bins"][slc].copy()
        eps = 1e-14 if dtype == "float64" else 1e-4
        bins[0] *= (1 - eps)
        bins[-1] *= (1 + eps)
        np_res = np.histogram(np.abs(fx_h[slc].astype(np.float64)), bins=bins)[0]
        assert res.sum() == np.prod(grid_shape)
        assert np.sum(np.abs(res - np_res)) <= (4 if dtype == "float64" else 400)

    # bounds passed by the caller (outer-shaped arrays): first/last bins absorb outliers
    b = {"min_f": np.full((2, 2), -1.), "max_f": np.full((2, 2), 2.),
         "min_log_f": np.full((2, 2), -3.), "max_log_f": np.full((2, 2), 1.)}
    clipped = hist(torch.from_numpy(fx), **b)
    assert clipped["linear"].sum() == 4 * np.prod(grid_shape)
    assert np.allclose(clipped["linear_bins"][0, 1], np.linspace(-1, 2, num_bins + 1))
    n_hi = np.sum(fx_h[1, 0] >= 2. - 3. / num_bins)
    assert clipped["linear"][1, 0][-1] == n_hi


def test_histogram_needs_device_arrays():
    decomp = ps.DomainDecomposition((1, 1, 1), 0, grid_shape=(4, 4, 4))
    hist = ps.Histogrammer(decomp, {"c": (ps.Field("fx") * 4, 1)}, 4, "float64")
    with pytest.raises(ps.PsbError, match="no CPU fallback"):
        hist(None, fx=torch.zeros(4, 4, 4, dtype=torch.float64))
