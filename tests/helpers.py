"""Test-side helpers (may import the oracle; product code may not)."""
import numpy as np
import torch

import oracle as O


class HostOps:
    """
    Oracle-backed halo operations for *host* tensors, injected into
    DomainDecomposition.host_ops by the gloo tests so that the rank arithmetic and
    message ordering of share_halos can be exercised without a GPU.
    """

    def wrap(self, fx, halo_shape, mask):
        h = [hh if (mask >> ax) & 1 else 0 for ax, hh in enumerate(halo_shape)]
        arr = fx.numpy()
        N = [n - 2 * hh for n, hh in zip(arr.shape, halo_shape)]
        hx, hy, hz = halo_shape
        if h[2]:
            arr[:, :, :hz] = arr[:, :, N[2]:N[2] + hz]
            arr[:, :, N[2] + hz:] = arr[:, :, hz:2 * hz]
        if h[0]:
            arr[:hx] = arr[N[0]:N[0] + hx]
            arr[N[0] + hx:] = arr[hx:2 * hx]
        if h[1]:
            arr[:, :hy] = arr[:, N[1]:N[1] + hy]
            arr[:, N[1] + hy:] = arr[:, hy:2 * hy]

    def pack(self, fx, buf, halo_shape, axis, unpack):
        assert axis == 1
        hy = halo_shape[1]
        Ny = fx.shape[1] - 2 * hy
        if not unpack:
            buf[1].copy_(fx[:, Ny:Ny + hy, :])
            buf[0].copy_(fx[:, hy:2 * hy, :])
        else:
            fx[:, :hy, :] = buf[1]
            fx[:, Ny + hy:, :] = buf[0]


def periodic_global(rng, grid_shape, h, dtype=np.float64):
    """Globally periodic padded array as in test/test_decomp.py:61-76."""
    h = (h,) * 3 if isinstance(h, int) else tuple(h)
    data = rng.random(tuple(N + 2 * hi for N, hi in zip(grid_shape, h))).astype(dtype)
    if h[0] > 0:
        data[:h[0], :, :] = data[-2 * h[0]:-h[0], :, :]
        data[-h[0]:, :, :] = data[h[0]:2 * h[0], :, :]
    if h[1] > 0:
        data[:, :h[1], :] = data[:, -2 * h[1]:-h[1], :]
        data[:, -h[1]:, :] = data[:, h[1]:2 * h[1], :]
    if h[2] > 0:
        data[:, :, :h[2]] = data[:, :, -2 * h[2]:-h[2]]
        data[:, :, -h[2]:] = data[:, :, h[2]:2 * h[2]]
    return data


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scale_err(got, ref):
    """max |got-ref| / max |ref|  (array-scale relative error)."""
    ref = np.asarray(ref)
    return float(np.max(np.abs(np.asarray(got) - ref)) / max(np.max(np.abs(ref)), 1e-300))


__all__ = ["HostOps", "periodic_global", "to_dev", "scale_err", "O"]
