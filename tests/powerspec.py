"""Spherically averaged power spectrum, as the reference's golden files were made (TEST INFRASTRUCTURE).

The reference's ``tests/produce_integration_test_data.py:367-371, 438-447`` calls
``powerbox.tools.get_power(field, boxlength=L, bins_upto_boxlen=True)`` (powerbox is not in this image).  What that
call computes, checked here against the ``k`` arrays stored next to every golden spectrum (agreement 1e-14):

* ``F = fftn(field) * (L / N)^3``, ``P = |F|^2 / L^3`` on the full complex grid, ``k = 2 pi * fftfreq``;
* ``int(N / 2.2)`` linear bins from 0 to the largest ``|k|`` along one axis (``pi N / L``), half-open, the k = 0 mode
  included in the first bin; modes beyond the last edge (the corners of the cube) dropped;
* per bin the plain mean of ``P`` and of ``|k|`` over the modes in it.
"""
import numpy as np


def get_power(field, boxlength):
    n = field.shape[0]
    assert field.shape == (n, n, n)
    ft = np.fft.fftn(np.asarray(field, np.float64)) * (boxlength / n) ** 3
    power = (ft.real**2 + ft.imag**2).ravel() / boxlength**3
    k1 = 2 * np.pi * np.fft.fftfreq(n, d=boxlength / n)
    k = np.sqrt(k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + k1[None, None, :] ** 2).ravel()
    nbins = int(n / 2.2)
    edges = np.linspace(0.0, np.abs(k1).max(), nbins + 1)
    idx = np.digitize(k, edges) - 1
    keep = idx < nbins
    count = np.bincount(idx[keep], minlength=nbins)
    return (np.bincount(idx[keep], weights=power[keep], minlength=nbins) / count,
            np.bincount(idx[keep], weights=k[keep], minlength=nbins) / count)


def step_pdf(data, xmin, xmax, nbins):
    """produce_integration_test_data.py:449-465: density histogram on ``linspace(xmin, xmax, nbins)`` edges with
    every value repeated (the golden stores the step-plot arrays)."""
    y, _ = np.histogram(data, bins=np.linspace(xmin, xmax, nbins), range=[xmin, xmax], density=True)
    return np.array([y, y]).T.flatten()
