"""Seeded synthetic particle sets: uniform-random and Zel'dovich-displaced.

These are INPUT GENERATORS for benchmarks and tests, not part of the MA -> Pk path; the device
versions use torch's Philox generator and torch.fft only to fabricate inputs."""
import numpy as np
import torch


def uniform_host(n, BoxSize, seed):
    """pos = rng.random((n,3), float32) * BoxSize, W = rng.random(n, float32) (docs/source/construction.rst:113)."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32) * np.float32(BoxSize)
    return pos


def zeldovich_host(n_side, BoxSize, seed, rms_cells=1.5, slope=-1.5, kcut_frac=0.25):
    """NumPy twin of zeldovich_device (same recipe, NumPy's generator): host inputs for the CPU legs and tests."""
    n = n_side
    rng = np.random.default_rng(seed)
    kf = 2.0 * np.pi / BoxSize
    k1 = (np.fft.fftfreq(n, d=1.0 / n) * kf).astype(np.float32)
    kz = (np.fft.rfftfreq(n, d=1.0 / n) * kf).astype(np.float32)
    k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    kc = kcut_frac * (n // 2) * kf
    amp = np.sqrt(k2 ** (slope / 2.0) * np.exp(-k2 / kc ** 2)).astype(np.float32)
    amp[0, 0, 0] = 0.0
    dk = (rng.standard_normal(k2.shape, dtype=np.float32) + 1j * rng.standard_normal(k2.shape, dtype=np.float32)) * amp
    disp = [np.fft.irfftn(1j * kv * dk / k2, s=(n, n, n)).astype(np.float32)
            for kv in (k1[:, None, None], k1[None, :, None], kz[None, None, :])]
    rms = np.sqrt(sum(np.mean(d.astype(np.float64) ** 2) for d in disp) / 3.0)
    scale = np.float32(rms_cells * (BoxSize / n) / max(rms, 1e-30))
    q = ((np.arange(n, dtype=np.float32) + np.float32(0.5)) * np.float32(BoxSize / n))
    pos = np.empty((n, n, n, 3), dtype=np.float32)
    shapes = ((n, 1, 1), (1, n, 1), (1, 1, n))
    for ax in range(3):
        pos[..., ax] = np.remainder(q.reshape(shapes[ax]) + disp[ax] * scale, np.float32(BoxSize))
    pos = pos.reshape(-1, 3)
    np.clip(pos, 0.0, np.nextafter(np.float32(BoxSize), np.float32(0)), out=pos)
    return pos


def slab_x_bounds(x0, x1, dims, BoxSize):
    """Smallest and largest float32 x whose CIC cell floor(fl32(x * fl32(dims/BoxSize))) lies in planes [x0, x1):
    positions generated "on their owner rank" must stay inside after the float32 product of the deposit, whose
    rounding can carry a position one ulp below a slab boundary onto the next plane."""
    inv = np.float32(dims) / np.float32(BoxSize)
    lo = np.float32(x0 * (BoxSize / dims))
    hi = np.float32(x1 * (BoxSize / dims))
    while np.float32(lo * inv) < np.float32(x0):
        lo = np.nextafter(lo, np.float32(np.inf))
    hi = np.nextafter(hi, np.float32(-np.inf))
    while np.float32(hi * inv) >= np.float32(x1):
        hi = np.nextafter(hi, np.float32(-np.inf))
    return float(lo), float(hi)


def uniform_device(n, BoxSize, seed, device, x_range=None):
    """Uniform positions on the device.  x_range=(lo,hi) restricts the first coordinate to the CLOSED interval
    [lo, hi] (per-slab generation for multi-GPU runs: same density everywhere, particles already on their owner;
    take the bounds from slab_x_bounds)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    pos = torch.rand((n, 3), generator=g, device=device, dtype=torch.float32)
    pos.mul_(float(BoxSize))
    if x_range is not None:
        lo, hi = x_range
        pos[:, 0].mul_((hi - lo) / float(BoxSize)).add_(lo)
        pos[:, 0].clamp_(min=lo, max=hi)
    return pos


def weights_device(n, seed, device, kind="uniform"):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) + 7919)
    if kind == "uniform":
        return torch.rand(n, generator=g, device=device, dtype=torch.float32)
    if kind == "heavy":       # HI-like heavy tail: exp(2*N(0,1))
        return torch.exp(2.0 * torch.randn(n, generator=g, device=device, dtype=torch.float32))
    return torch.ones(n, device=device, dtype=torch.float32)


def zeldovich_device(n_side, BoxSize, seed, device, rms_cells=1.5, slope=-1.5, kcut_frac=0.25):
    """n_side^3 particles on a lattice q=(i+0.5)L/n displaced by psi = IFFT(i k/k^2 delta_k), delta_k a
    Gaussian field with P(k) = A k^slope exp(-(k/kc)^2), scaled to an rms displacement of `rms_cells`
    lattice spacings, wrapped into [0,L) (conventions of density_field_library.pyx:113-126)."""
    n = n_side
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    kf = 2.0 * np.pi / BoxSize
    k1 = torch.fft.fftfreq(n, d=1.0 / n, device=device) * kf
    kz = torch.fft.rfftfreq(n, d=1.0 / n, device=device) * kf
    k2 = k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    kc = kcut_frac * (n // 2) * kf
    amp = torch.sqrt(k2 ** (slope / 2.0) * torch.exp(-k2 / kc ** 2))
    amp[0, 0, 0] = 0.0
    dk = torch.complex(torch.randn(k2.shape, generator=g, device=device),
                       torch.randn(k2.shape, generator=g, device=device)) * amp
    del amp
    q = (torch.arange(n, device=device, dtype=torch.float32) + 0.5) * (BoxSize / n)
    pos = torch.empty((n, n, n, 3), device=device, dtype=torch.float32)
    disp = []
    for ax, kv in enumerate((k1[:, None, None], k1[None, :, None], kz[None, None, :])):
        disp.append(torch.fft.irfftn(1j * kv * dk / k2, s=(n, n, n)).float())
    rms = torch.sqrt(sum((d.double() ** 2).mean() for d in disp) / 3.0).item()
    scale = rms_cells * (BoxSize / n) / max(rms, 1e-30)
    shapes = ((n, 1, 1), (1, n, 1), (1, 1, n))
    for ax in range(3):
        pos[..., ax] = torch.remainder(q.view(shapes[ax]) + disp[ax] * scale, float(BoxSize))
    del disp, dk
    pos = pos.view(-1, 3)
    pos.clamp_(min=0.0, max=float(np.nextafter(np.float32(BoxSize), np.float32(0))))
    return pos
