"""Host driver of the LPT field pipeline - the numerical spine of ic_generator::run
(src/ic_generator.cc:90-982) expressed as calls into the C ABI.

What stays on the host (and keeps the reference's plugin APIs): the white-noise plugin fills a
grid, the transfer-function plugin is evaluated once per distinct |k| into a radial table
(SURVEY.md 0.7), the output plugin receives host arrays.  Everything between
"conditioned white noise in k-space" and "particle arrays / component fields" runs on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from .grid import Grid_FFT, Grid_FFT_complex, OrszagConvolver, NaiveConvolver, Particles, kspace_id, rspace_id, _lib, LATTICE


@dataclass
class lpt_config:
    """the [setup]/[output] keys that reach the hot path"""
    GridRes: int = 32
    BoxLength: float = 100.0
    LPTorder: int = 3
    ParticleLoad: str = "sc"
    DoFixing: bool = False
    DoInversion: bool = False
    DoRemoveCornerModes: bool = False
    Dplus0: float = 1.0          # cosmology::calculator::get_growth_factor(astart)
    vfac: float = 1.0            # cosmology::calculator::get_vfact(astart)
    lunit: float = 1.0           # output_plugin::position_unit()
    vunit: float = 1.0           # output_plugin::velocity_unit()
    real64: bool = True          # output_plugin::has_64bit_reals()
    ids64: bool = False          # output_plugin::has_64bit_ids()
    tides: tuple | None = None   # LSS_aniso_l{x,y,z}
    ConvolverType: str = "orszag"  # the reference's compile-time CONVOLVER_TYPE (CMakeLists.txt): "orszag" | "naive"
    DoBaryons: bool = False      # masked lattice: selects the species mask (include/particle_generator.hh:184-205)
    ParticleMaskType: int = 3
    glass: tuple | None = None   # ParticleLoad "glass": (glass_loader::glass_posr in grid units [np][3], offset)


def radial_table(n: int, boxlen: float, amplitude) -> np.ndarray:
    """table[m] = A(kfac*sqrt(m)), m = |n|^2 = 0..3(N/2)^2: one host call of the transfer-function
    plugin per distinct |k| instead of one per mode (include/cosmology_calculator.hh:405-408)."""
    h = n // 2
    m = np.arange(3 * h * h + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.asarray(amplitude(2.0 * np.pi / boxlen * np.sqrt(m)), dtype=np.float64)
    return np.ascontiguousarray(np.where(np.isfinite(t), t, 0.0))


class LPTPipeline:
    """Device-resident state of one IC run: wnoise, phi, phi2, phi3, A3[3], tmp + the convolver."""

    def __init__(self, cfg: lpt_config):
        self.cfg = cfg
        n, Lb = cfg.GridRes, cfg.BoxLength
        if cfg.ConvolverType not in ("orszag", "naive"):
            raise RuntimeError(f"unknown convolver type '{cfg.ConvolverType}'")
        if cfg.LPTorder > 1 and n % 2 and cfg.ConvolverType == "orszag":
            # src/ic_generator.cc:118-121
            raise RuntimeError("Orszag convolver for LPTorder>1 requires even grid size!")
        self.N, self.Lbox = (n, n, n), (Lb, Lb, Lb)
        self._lib = _lib()
        mk = lambda alloc=True, sp=rspace_id: Grid_FFT(self.N, self.Lbox, alloc, sp)
        order = min(cfg.LPTorder, 3)
        self.order = order
        self.wnoise = mk()
        self.phi = mk()
        self.phi2 = mk(order > 1)
        self.phi3 = mk(order > 2)
        self.A3 = [mk(order > 2) for _ in range(3)]
        self.tmp = mk()
        mkconv = NaiveConvolver if cfg.ConvolverType == "naive" else OrszagConvolver
        self.conv = mkconv(self.N, self.Lbox) if order > 1 else None
        self.particles = None
        self._ids_streamed = False
        self.lg = None   # the gradient operator of src/ic_generator.cc:377-381: None = op::fourier_gradient, or a lattice_gradient

    # -- stages -------------------------------------------------------------------------
    def set_noise_k(self, W: np.ndarray):
        """white noise in k-space [kx][ky][kz] as left by RNG_plugin::Fill_Grid + forward FFT"""
        self.wnoise.set_kspace(W)

    def condition_noise(self):
        c = self.cfg
        volfac = (c.BoxLength / c.GridRes / 2.0 / np.pi) ** 1.5  # src/ic_generator.cc:210
        L.check(self._lib.mfb_noise_condition(self.wnoise._h, volfac, int(c.DoFixing), int(c.DoInversion),
                                              int(c.DoRemoveCornerModes)), self._lib)

    def compute_phi1(self, table: np.ndarray):
        t = np.ascontiguousarray(table, dtype=np.float64)
        L.check(self._lib.mfb_apply_radial_table(self.phi._h, self.wnoise._h, t.ctypes.data_as(C.c_void_p), t.size, 1),
                self._lib)

    def compute_sources(self):
        """phi2, phi3, A3 and the growth-factor scaling (src/ic_generator.cc:475-574)"""
        c = self.cfg
        if self.order == 1:
            self.phi *= -c.Dplus0
            return
        A3 = (C.c_void_p * 3)(*[(g._h if g.is_allocated() else None) for g in self.A3])
        lam = (C.c_double * 3)(*c.tides) if c.tides is not None else None
        L.check(self._lib.mfb_lpt_sources(self.conv._h, self.order, float(c.Dplus0), self.phi._h,
                                          self.phi2._h if self.order > 1 else None,
                                          self.phi3._h if self.order > 2 else None, A3, lam), self._lib)

    def _assemble(self, d, w, wc, scale, vbc=None, velocity=False):
        A3 = (C.c_void_p * 3)(*[(g._h if g.is_allocated() else None) for g in self.A3])
        o = self.order
        tab, ntab, cv, wn = None, 0, 0.0, None
        if vbc is not None:
            t = np.ascontiguousarray(vbc[0], dtype=np.float64)
            tab, ntab, cv, wn = t.ctypes.data_as(C.c_void_p), t.size, float(vbc[1]), self.wnoise._h
        glass = self.cfg.ParticleLoad == "glass"   # particles read out by CIC interpolation: compensation kernel (:861, :924)
        L.check(self._lib.mfb_lpt_assemble_lg(self.tmp._h, d, self.phi._h, self.phi2._h if o > 1 else None,
                                              self.phi3._h if o > 2 else None, A3 if o > 2 else None,
                                              w[0], w[1], w[2], wc, scale, wn, tab, ntab, cv,
                                              self.lg._h if self.lg is not None else None, int(velocity), int(glass)), self._lib)

    def displacement_field(self, d, transform=True):
        """tmp <- Psi_d in real space (src/ic_generator.cc:832-871).  transform=False leaves the assembled field in k-space:
        handed like that to Particles.set_positions, the backward transform and the lattice read-out are one pass for
        single-shift (sc) loads"""
        c = self.cfg
        self._assemble(d, (1.0, 1.0, 1.0), 1.0, c.lunit / c.BoxLength)
        if transform:
            self.tmp.FourierTransformBackward()
        return self.tmp

    def velocity_field(self, d, vbc=None, transform=True):
        """tmp <- v_d in real space (src/ic_generator.cc:888-942)"""
        c = self.cfg
        v = c.vfac
        scale = c.vunit / c.BoxLength
        if c.tides is not None and self.order > 1:
            # external tides: anisotropic velocity growth factor alpha_d^2, alpha_d = 1 - D+ lambda_d
            # (src/ic_generator.cc:238-242, applied at :931-934 to the whole mode incl. the v_bc term)
            scale *= (1.0 - c.Dplus0 * c.tides[d]) ** 2
        self._assemble(d, (v, 2 * v, 3 * v), 3 * v, scale, vbc, velocity=True)
        if transform:
            self.tmp.FourierTransformBackward()
        return self.tmp

    def density_field(self):
        """tmp <- delta = k^2 phi in real space (src/ic_generator.cc:962-973)"""
        L.check(self._lib.mfb_density_from_phi(self.tmp._h, self.phi._h), self._lib)
        self.tmp.FourierTransformBackward()
        return self.tmp

    def eulerian_fields(self, delta_bc_table, C_species, velocity_unit=1.0, sink=None):
        """Species written as output_type::field_eulerian - the semiclassical PPT branch (src/ic_generator.cc:685-803),
        after compute_sources(): rho = sqrt(1 + C delta_bc), psi = rho exp(i(phi[+phi2])/(hbar D)), one drift (c2c FFT
        pair) [+ kick], density |psi|^2-1 and v_d = hbar Im(psi* grad_d psi)/(1+rho) from three more c2c pairs.
        `delta_bc_table`: get_amplitude_delta_bc on the integer |n|^2 lattice (radial_table); `sink(name, grid)` plays
        write_grid_data; without it the fields are returned as host arrays.  Leaves phi (and phi2) in real space like
        the reference does.  Returns (fields, hbar)."""
        c = self.cfg
        lib = self._lib
        o = min(self.order, 2)
        rho = Grid_FFT(self.N, self.Lbox)
        t = np.ascontiguousarray(delta_bc_table, dtype=np.float64)
        L.check(lib.mfb_apply_radial_table(rho._h, self.wnoise._h, t.ctypes.data_as(C.c_void_p), t.size, 0), lib)  # :697-699
        rho.zero_DC_mode()
        rho.FourierTransformBackward()
        L.check(lib.mfb_ppt_sqrt_density(rho._h, float(C_species)), lib)
        self.phi.FourierTransformBackward()
        hbar = self.phi.max_neighbour_diff() / np.pi / c.Dplus0 * 1.01   # :733-734, hbar_safefac = 1.01
        psi = Grid_FFT_complex(self.N, self.Lbox)
        if o >= 2:
            self.phi2.FourierTransformBackward()
        L.check(lib.mfb_ppt_init_psi(psi._h, rho._h, self.phi._h, self.phi2._h if o >= 2 else None, hbar, c.Dplus0), lib)
        psi.FourierTransformForward()
        L.check(lib.mfb_ppt_drift(psi._h, hbar, c.Dplus0), lib)
        psi.FourierTransformBackward()
        if o >= 2:
            L.check(lib.mfb_ppt_kick(psi._h, self.phi2._h, hbar, c.Dplus0), lib)
        L.check(lib.mfb_ppt_density(rho._h, psi._h), lib)
        out = {}

        def emit(name, g):
            if sink is not None:
                sink(name, g)
            else:
                out[name] = g.get_real()
        emit("density", rho)
        grad = Grid_FFT_complex(self.N, self.Lbox)
        vunit = c.Dplus0 * c.vfac / c.BoxLength * velocity_unit   # :779
        for d, nm in enumerate(("vx", "vy", "vz")):
            grad.copy_from(psi)
            grad.FourierTransformForward()
            grad.apply_gradient(d)
            grad.FourierTransformBackward()
            L.check(lib.mfb_ppt_velocity(self.tmp._h, psi._h, grad._h, rho._h, hbar, vunit), lib)
            emit(nm, self.tmp)
        return out, hbar

    def make_particles(self, lattice_index=0, with_masses=False, id_offset=0):
        c = self.cfg
        if self.particles is not None:
            if self.conv is not None and c.ConvolverType == "orszag":
                self.conv.borrow_accumulator(None)  # a previous loan ends with its lender
            self.particles = None                   # release the old container first: two would not fit at 1024^3
        self._lattice_index = lattice_index
        mask = None
        if c.ParticleLoad == "masked":
            from .grid import particle_type_mask

            mask = particle_type_mask(c.DoBaryons, c.ParticleMaskType)
        self.particles = Particles(c.ParticleLoad, lattice_index, c.real64, c.ids64, with_masses, id_offset, self.tmp,
                                   mask=mask, glass=c.glass)
        if self.conv is not None and c.ConvolverType == "orszag":
            # the position/velocity arrays hold nothing while the sources are computed: the convolver may use them as
            # workspace (slots of the shared-transform source stage, or the accumulator of product groups)
            self.conv.borrow_accumulator(self.particles)
        return self.particles

    def fill_particles(self, lattice_index=0, stream_to=None):
        """the species output loop for particle output (src/ic_generator.cc:832-955).  With `stream_to` (dict of pinned
        host arrays "pos", "vel", "ids") every attribute is handed to the host as soon as it is complete, on the copy
        stream, overlapping the remaining component FFTs; the caller ends with particles.stream_wait()."""
        c = self.cfg
        if self.particles is None or getattr(self, "_lattice_index", 0) != lattice_index:
            # a second species on the secondary lattice (baryons, lattice_index 1) needs its own container
            self.make_particles(lattice_index)
        if stream_to is not None and "ids" in stream_to and not self._ids_streamed:
            self.particles.stream_out("ids", stream_to["ids"])
        self._ids_streamed = False
        for d in range(3):
            self.particles.set_positions(d, c.lunit, self.displacement_field(d, transform=False))
        if stream_to is not None and "pos" in stream_to:
            self.particles.stream_out("pos", stream_to["pos"].reshape(-1))
        for d in range(3):
            self.particles.set_velocities(d, self.velocity_field(d, transform=False))
        if stream_to is not None and "vel" in stream_to:
            self.particles.stream_out("vel", stream_to["vel"].reshape(-1))
        return self.particles

    def run_device(self, table: np.ndarray, particles=True, stream_to=None):
        """conditioned noise resident in k-space -> all LPT terms -> Psi, v (-> particle AoS), all on device"""
        if particles and stream_to is not None and "ids" in stream_to and self.particles is not None:
            # the IDs do not depend on the fields: their hand-off runs under the source stage
            self.particles.stream_out("ids", stream_to["ids"])
            self._ids_streamed = True
        self.compute_phi1(table)
        self.compute_sources()
        if particles:
            return self.fill_particles(stream_to=stream_to)
        out = {}
        for d, nm in enumerate("xyz"):
            out["d" + nm] = self.displacement_field(d).get_real()
        for d, nm in enumerate("xyz"):
            out["v" + nm] = self.velocity_field(d).get_real()
        return out
