"""Seeded synthetic particle sets for the BASELINE.json configs.

All outputs are float64 C-order, G = 1, total mass 1 unless stated, radii truncated so the
octree depth stays inside the 42-level path key, exact duplicate positions rejected.
"""
from __future__ import annotations

import numpy as np


def _isotropic(rng, r):
    n = r.shape[0]
    cos_t = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sin_t = np.sqrt(1.0 - cos_t * cos_t)
    pos = np.empty((n, 3), dtype=np.float64)
    pos[:, 0] = r * sin_t * np.cos(phi)
    pos[:, 1] = r * sin_t * np.sin(phi)
    pos[:, 2] = r * cos_t
    return pos


def _dedup(pos, rng):
    """Reject exact duplicate rows by nudging them (vanishingly rare for float64 draws)."""
    _, first = np.unique(pos, axis=0, return_index=True)
    if first.shape[0] != pos.shape[0]:
        dup = np.setdiff1d(np.arange(pos.shape[0]), first)
        pos[dup] += rng.normal(0.0, 1e-9, (dup.shape[0], 3))
    return pos


def plummer(n, seed=1, a=1.0, rmax=50.0, dedup=True):
    """Plummer sphere, inverse-CDF radii r = a / sqrt(u^(-2/3) - 1), truncated at rmax*a (config 1)."""
    rng = np.random.default_rng(seed)
    umax = (1.0 + (rmax) ** -2.0) ** -1.5  # M(<rmax)/M
    u = rng.uniform(0.0, umax, n)
    u = np.maximum(u, 1e-300)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    pos = _isotropic(rng, r)
    if dedup:
        pos = _dedup(pos, rng)
    mass = np.full(n, 1.0 / n)
    return np.ascontiguousarray(pos), mass


def hernquist(n, seed=2, a=1.0, rmax=100.0, dedup=True):
    """Hernquist halo, r = a sqrt(u) / (1 - sqrt(u)), truncated at rmax*a (config 2)."""
    rng = np.random.default_rng(seed)
    umax = (rmax / (1.0 + rmax)) ** 2
    u = rng.uniform(0.0, umax, n)
    s = np.sqrt(u)
    r = a * s / (1.0 - s)
    pos = _isotropic(rng, r)
    if dedup:
        pos = _dedup(pos, rng)
    mass = np.full(n, 1.0 / n)
    return np.ascontiguousarray(pos), mass


def _nfw_radii(rng, n, c):
    """NFW radii in units of r_vir (r_s = 1/c), by tabulated inverse CDF."""
    x = np.logspace(-4, np.log10(c), 4096)
    mu = np.log1p(x) - x / (1.0 + x)
    mu /= mu[-1]
    u = rng.uniform(0.0, 1.0, n)
    return np.interp(u, mu, x) / c


def nfw_disc(n, seed=3, c=10.0, disc_frac_n=0.2, disc_frac_m=0.05, rd=0.03, zd_over_rd=0.1, dedup=True):
    """NFW halo (r <= r_vir = 1) + exponential stellar disc (config 3). Returns pos, mass, softening."""
    rng = np.random.default_rng(seed)
    nd = int(round(n * disc_frac_n))
    nh = n - nd
    r = _nfw_radii(rng, nh, c)
    halo = _isotropic(rng, r)
    # exponential disc: R ~ Gamma(2, rd), z ~ sech^2-like via logistic
    R = rng.gamma(2.0, rd, nd)
    R = np.minimum(R, 1.0)
    ph = rng.uniform(0.0, 2.0 * np.pi, nd)
    z = rng.logistic(0.0, 0.5 * zd_over_rd * rd, nd)
    disc = np.stack([R * np.cos(ph), R * np.sin(ph), z], axis=1)
    pos = np.concatenate([halo, disc])
    mass = np.concatenate([np.full(nh, (1.0 - disc_frac_m) / nh), np.full(nd, disc_frac_m / max(nd, 1))])
    soft = np.concatenate([np.full(nh, 1.0e-3), np.full(nd, 5.0e-4)])
    if dedup:
        pos = _dedup(pos, rng)
    order = rng.permutation(n)  # mixed particle order, so contiguous target shards cost about the same
    return np.ascontiguousarray(pos[order]), np.ascontiguousarray(mass[order]), np.ascontiguousarray(soft[order])


def zoom_families(n, seed=4, dedup=False):
    """dm/gas/star zoom-shaped set (config 4): nested high-res centre + low-res shell.
    Fractions 0.5/0.3/0.2, masses 1/0.19/0.05 (normalised to total mass 1), gas h ∝ local spacing."""
    rng = np.random.default_rng(seed)
    ndm, ngas = int(0.5 * n), int(0.3 * n)
    nstar = n - ndm - ngas
    # dm: 70% in a Hernquist core (a=0.05, r<=1) + 30% low-res shell 1<r<8
    n_core = int(0.7 * ndm)
    u = rng.uniform(0.0, (1.0 / 1.05) ** 2, n_core)
    s = np.sqrt(u)
    r_core = 0.05 * s / (1.0 - s)
    r_shell = (1.0 + rng.uniform(0.0, 1.0, ndm - n_core) * (8.0 ** 3 - 1.0)) ** (1.0 / 3.0)
    dm = _isotropic(rng, np.concatenate([r_core, r_shell]))
    # gas: Plummer-like a=0.1 truncated at 1
    ug = rng.uniform(0.0, (1.0 + 0.1 ** 2) ** -1.5, ngas)
    rg = 0.1 / np.sqrt(np.maximum(ug, 1e-300) ** (-2.0 / 3.0) - 1.0)
    gas = _isotropic(rng, rg)
    # stars: exponential disc rd=0.02
    R = np.minimum(rng.gamma(2.0, 0.02, nstar), 1.0)
    ph = rng.uniform(0.0, 2.0 * np.pi, nstar)
    z = rng.logistic(0.0, 0.001, nstar)
    star = np.stack([R * np.cos(ph), R * np.sin(ph), z], axis=1)
    pos = np.concatenate([dm, gas, star])
    mass = np.concatenate([np.full(ndm, 1.0), np.full(ngas, 0.19), np.full(nstar, 0.05)])
    mass /= mass.sum()
    # softenings: dm/star constant, gas ∝ local spacing ~ (r^2+a^2)^(1/2) * n^(-1/3)
    h_gas = 0.5 * np.sqrt(rg * rg + 0.1 ** 2) * max(ngas, 1) ** (-1.0 / 3.0)
    soft = np.concatenate([np.full(ndm, 2.0e-3), h_gas, np.full(nstar, 5.0e-4)])
    if dedup:
        pos = _dedup(pos, rng)
    order = rng.permutation(n)
    return np.ascontiguousarray(pos[order]), np.ascontiguousarray(mass[order]), np.ascontiguousarray(soft[order])


ZOOM_BLOCK = 1 << 18
_ZOOM_FRAC = (0.5, 0.3, 0.2)          # dm / gas / star number fractions
_ZOOM_MASS = (1.0, 0.19, 0.05)        # relative particle masses


def zoom_range(n, lo, hi, seed=4):
    """Particles [lo, hi) of the N-particle zoom set of config 4, generated block by block from
    ``default_rng([seed, block])`` so that ANY rank layout reproduces the same global set (the set depends on
    (n, seed) only, never on the world size). Same shapes as :func:`zoom_families`: dm = Hernquist core
    (a=0.05, r<=1) + low-resolution shell 1<r<8, gas = Plummer-like a=0.1 with h ∝ local spacing, stars =
    exponential disc; families are interleaved at random (each particle draws its family), so contiguous index
    ranges are statistically uniform samples. Masses are normalised by the EXPECTED family counts
    (sum ≈ 1 to ~1e-4), which keeps them independent of the realised counts of other blocks.
    Returns pos (hi-lo,3), mass, softening, family (uint8: 0 dm, 1 gas, 2 star)."""
    lo, hi = int(lo), int(hi)
    assert 0 <= lo <= hi <= n
    norm = n * sum(f * w for f, w in zip(_ZOOM_FRAC, _ZOOM_MASS))
    h_gas_scale = 0.5 * max(_ZOOM_FRAC[1] * n, 1.0) ** (-1.0 / 3.0)
    pos_out, mass_out, soft_out, fam_out = [], [], [], []
    last_block = (hi - 1) // ZOOM_BLOCK if hi > lo else lo // ZOOM_BLOCK - 1
    for b in range(lo // ZOOM_BLOCK, last_block + 1):
        b0 = b * ZOOM_BLOCK
        cnt = min(ZOOM_BLOCK, n - b0)
        rng = np.random.default_rng([seed, b])
        u = rng.random(cnt)
        fam = np.where(u < _ZOOM_FRAC[0], 0, np.where(u < _ZOOM_FRAC[0] + _ZOOM_FRAC[1], 1, 2)).astype(np.uint8)
        pos = np.empty((cnt, 3))
        soft = np.empty(cnt)
        # dm: 70 % Hernquist core, 30 % shell
        k = np.nonzero(fam == 0)[0]
        core = rng.random(k.shape[0]) < 0.7
        s = np.sqrt(rng.uniform(0.0, (1.0 / 1.05) ** 2, k.shape[0]))
        r_core = 0.05 * s / (1.0 - s)
        r_shell = (1.0 + rng.uniform(0.0, 1.0, k.shape[0]) * (8.0 ** 3 - 1.0)) ** (1.0 / 3.0)
        pos[k] = _isotropic(rng, np.where(core, r_core, r_shell))
        soft[k] = 2.0e-3
        # gas
        k = np.nonzero(fam == 1)[0]
        ug = rng.uniform(0.0, (1.0 + 0.1 ** 2) ** -1.5, k.shape[0])
        rg = 0.1 / np.sqrt(np.maximum(ug, 1e-300) ** (-2.0 / 3.0) - 1.0)
        pos[k] = _isotropic(rng, rg)
        soft[k] = h_gas_scale * np.sqrt(rg * rg + 0.1 ** 2)
        # stars
        k = np.nonzero(fam == 2)[0]
        R = np.minimum(rng.gamma(2.0, 0.02, k.shape[0]), 1.0)
        ph = rng.uniform(0.0, 2.0 * np.pi, k.shape[0])
        z = rng.logistic(0.0, 0.001, k.shape[0])
        pos[k] = np.stack([R * np.cos(ph), R * np.sin(ph), z], axis=1)
        soft[k] = 5.0e-4
        mass = np.asarray(_ZOOM_MASS)[fam] / norm
        a, e = max(lo, b0) - b0, min(hi, b0 + cnt) - b0
        pos_out.append(pos[a:e]); mass_out.append(mass[a:e]); soft_out.append(soft[a:e]); fam_out.append(fam[a:e])
    if not pos_out:
        return np.empty((0, 3)), np.empty(0), np.empty(0), np.empty(0, np.uint8)
    return (np.ascontiguousarray(np.concatenate(pos_out)), np.concatenate(mass_out), np.concatenate(soft_out),
            np.concatenate(fam_out))


def zoom_set(n, seed=4):
    """The whole deterministic zoom set (config 4): pos, mass, softening."""
    pos, mass, soft, _ = zoom_range(n, 0, n, seed)
    return pos, mass, soft


def rz_grid_targets(m, seed=5, rmin=1e-3, rmax=5.0):
    """Log-spaced (R, z) grid targets, random azimuth (config 5)."""
    rng = np.random.default_rng(seed)
    k = int(np.ceil(np.sqrt(m)))
    R = np.logspace(np.log10(rmin), np.log10(rmax), k)
    z = np.concatenate([[0.0], np.logspace(np.log10(rmin), np.log10(rmax), k - 1)])
    RR, ZZ = np.meshgrid(R, z, indexing="ij")
    RR, ZZ = RR.ravel()[:m], ZZ.ravel()[:m]
    ph = rng.uniform(0.0, 2.0 * np.pi, RR.shape[0])
    return np.ascontiguousarray(np.stack([RR * np.cos(ph), RR * np.sin(ph), ZZ], axis=1))


def uniform_cube(n, seed, with_masses=True):
    """The reference tests' distribution: uniform [-0.5,0.5)^3, masses 0.5+U(0,1) (gravity_tests.rs:3-20)."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)) - 0.5
    mass = 0.5 + rng.random(n) if with_masses else None
    return pos, mass
