"""Pin the CPU oracle against every golden file of the reference (CPU-only)."""
import numpy as np

import golden_util as G
import lpt_oracle as O


def close(a, b):
    return np.allclose(a, b, rtol=G.RTOL, atol=G.ATOL)


def test_bootstrap_is_clean():
    g, phi, vfac = G.bootstrap()
    # NGenIC leaves the Nyquist planes zero (random_ngenic.cc:117)
    nh = g.nh
    ny = max(np.abs(phi[nh]).max(), np.abs(phi[:, nh]).max(), np.abs(phi[:, :, nh]).max())
    assert ny < 1e-15 * np.abs(phi).max() + 1e-16
    assert abs(vfac - 396.6919511923062) < 1e-6


def test_golden_1lpt_sc_generic():
    g, phi, vfac = G.bootstrap()
    gold = G.load("test_1lpt_sc_generic")
    pots = O.lpt_potentials(g, phi, 1, -1.0)
    f = O.fields(g, pots, 1, vfac)
    for c in ["dx", "dy", "dz", "vx", "vy", "vz", "delta"]:
        assert close(f[c], gold["DM_" + c]), c


def test_golden_2lpt_baryons_generic():
    g, phi, vfac = G.bootstrap()
    gold = G.load("test_2lpt_baryons_generic")
    assert bool(gold["flag__BA_equals_DM_bitwise"])
    pots = O.lpt_potentials(g, phi, 2, -1.0)
    f = O.fields(g, pots, 2, vfac)
    for c in ["dx", "dy", "dz", "vx", "vy", "vz", "delta"]:
        assert close(f[c], gold["DM_" + c]), c
    # mass grids: (1 + C*rho)*Omega*munit with rho == 0 for E&H (ic_generator.cc:673-675)
    assert np.allclose(gold["massgrid__DM_"], 0.3099 - 0.048891054, rtol=1e-12)
    assert np.allclose(gold["massgrid__BA_"], 0.048891054, rtol=1e-12)


def _particles(order, lattice, lattice_index, fmt):
    g, phi, vfac = G.bootstrap()
    lunit, vunit, r64 = G.units(fmt)
    pots = O.lpt_potentials(g, phi, order, -1.0)
    lt = O.LATTICE_TYPES[lattice]
    pos, vel = [], []
    for d in range(3):
        dr = g.bwd(O.displacement_k(g, pots, order, d, lunit / G.L))
        pos.append(O.set_positions(g, dr, d, lt, lattice_index, lunit, r64))
    for d in range(3):
        vr = g.bwd(O.velocity_k(g, pots, order, d, vfac, vunit / G.L))
        vel.append(O.set_velocities(g, vr, lt, lattice_index, r64))
    return np.stack(pos, 1), np.stack(vel, 1)


def test_golden_2lpt_sc_gadget():
    gold = G.load("test_2lpt_sc_gadget")
    pos, vel = _particles(2, "sc", 0, "gadget")
    assert pos.dtype == np.float32
    assert close(pos, gold["PartType1__Coordinates"])
    assert close(vel, gold["PartType1__Velocities"])
    # float32 goldens at rtol 1e-9 effectively demand identical rounding
    assert (pos == gold["PartType1__Coordinates"]).mean() > 0.9999
    ids = O.particle_ids(G.N, 0, 0, False)
    assert ids.dtype == np.uint32 and np.array_equal(ids, gold["PartType1__ParticleIDs"])


def test_golden_2lpt_baryons_vrel_gadget():
    gold = G.load("test_2lpt_baryons_vrel_gadget")
    pos, vel = _particles(2, "sc", 0, "gadget")
    assert close(pos, gold["PartType1__Coordinates"])
    assert close(vel, gold["PartType1__Velocities"])
    posb, velb = _particles(2, "sc", 1, "gadget")  # baryons: shifted secondary lattice
    assert close(posb, gold["PartType0__Coordinates"])
    assert close(velb, gold["PartType0__Velocities"])
    assert np.array_equal(O.particle_ids(G.N, 0, 0), gold["PartType1__ParticleIDs"])
    assert np.array_equal(O.particle_ids(G.N, 0, 1), gold["PartType0__ParticleIDs"])
    # masses: Omega_species * munit / N^3, munit = rhoc*L^3 (output_gadget_hdf5.cc:85)
    g = O.GridSpec(G.N, G.L)
    munit = 27.7519737 * G.L ** 3
    ones = np.ones((G.N,) * 3)
    mdm = O.set_masses(g, ones * (0.3099 - 0.048891054) * munit, 0, 0, 1.0, False)
    mba = O.set_masses(g, ones * 0.048891054 * munit, 0, 1, 1.0, False)
    assert close(mdm, gold["PartType1__Masses"])
    assert close(mba, gold["PartType0__Masses"])


def test_golden_3lpt_bcc_swift():
    gold = G.load("test_3lpt_bcc_swift")
    pos, vel = _particles(3, "bcc", 0, "swift")
    assert pos.shape == (65536, 3)
    assert close(pos, gold["PartType1__Coordinates"])
    assert close(vel, gold["PartType1__Velocities"])


def test_convolve2_is_dealiased_product():
    """U(conv2(a,b)) == fwd_N(a_r*b_r) restricted to |n|<N/2 when a,b are band-limited to N/4
    (then the plain N-grid product is alias-free too)."""
    n = 24
    g = O.GridSpec(n, 1.0)
    rng = np.random.default_rng(1)
    def bl(x):
        X = g.fwd(x)
        m = np.abs(g.nsigned) < n / 4 - 1e-9
        mask = m[:, None, None] & m[None, :, None] & m[None, None, : g.nzc]
        return X * mask
    A, B = bl(rng.standard_normal((n,) * 3)), bl(rng.standard_normal((n,) * 3))
    ref = g.fwd(g.bwd(A) * g.bwd(B))
    got = O.OrszagConvolver(g).convolve2(A, B)
    nh = n // 2
    ref[nh], ref[:, nh], ref[:, :, nh] = 0, 0, 0
    assert np.abs(got - ref).max() < 1e-13 * np.abs(ref).max()


def test_ngenic_noise_is_the_reference_s():
    """pins oracle.Ranlxd1 / ngenic_noise (and with them the device fill mfb_noise_fill_ngenic_ranlxd and the GSL stand-in of
    tests/cpp) against the reference's REAL output: the golden density field of the 1LPT regression run (made with GSL's
    gsl_rng_ranlxd1, seed 12345) is delta_k = T(|k|) W_k with a real positive T - so delta_k / W_k must be real, positive and
    the same for all modes of one |k| shell, mode by mode."""
    g = O.GridSpec(G.N, G.L)
    W = O.ngenic_noise(G.N, 12345)
    dk = g.fwd(G.load("test_1lpt_sc_generic")["DM_delta"])
    m = np.abs(W) > 0
    assert m.sum() == (G.N - 1) ** 2 * (G.N // 2 - 1) + ((G.N - 1) ** 2 - 1)   # every mode without an index N/2, minus DC
    ratio = dk[m] / W[m]
    assert np.abs(ratio.imag).max() < 1e-13 * np.abs(ratio.real).max()
    assert ratio.real.min() > 0
    k2 = np.round(g.k2()[m] / g.k2()[1, 0, 0]).astype(np.int64)
    for shell in np.unique(k2)[:200]:
        r = ratio.real[k2 == shell]
        assert r.max() - r.min() < 1e-12 * r.mean(), shell
    assert not dk[~m].any() or np.abs(dk[~m]).max() < 1e-12 * np.abs(dk).max()
