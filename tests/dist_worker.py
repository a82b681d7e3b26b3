"""Worker of the multi-rank tests: run under torchrun (one process per rank).

  --emu   : CPU thread-emulation build of the kernels + gloo, exchange through all_to_all ("staged" mode);
            validates the slab algorithm and the host logic without a GPU (pytest -m "not gpu").
  default : libmfb200.so on cuda:LOCAL_RANK, exchange fused into the FFT passes over peer memory.
  --staged: GPU build with the exchange done by NCCL all_to_all (the library-collective baseline).

Checks, all against the single-process CPU oracle on the same seeded inputs: 3-D FFT round trip, k-space
operators incl. reductions, every named product, whole 1/2/3LPT pipelines with particles (ids exact), and
equality with the reference's slab bookkeeping (local_0/1_start/size)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emu", action="store_true")
    ap.add_argument("--staged", action="store_true")
    ap.add_argument("--grid", type=int, default=32)
    ap.add_argument("--lattice", default="bcc")
    ap.add_argument("--nslots", type=int, default=0,
                    help="exchange buffers in the peer-mapped block (0: chosen from the free memory; 2: per-product source stage)")
    ap.add_argument("--borrow", action="store_true",
                    help="no own accumulator (as under memory pressure at 2048^3 / 8 GPUs): it is borrowed from the particle arrays")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import lpt_oracle as O
    import monofonic_b200 as M
    import parity_cases as PC
    from monofonic_b200 import dist as MD

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.emu:
        import emu_util

        dist.init_process_group("gloo")
        MD.init(0, args.grid, True, staged=True, lib_path=emu_util.EMU, nslots=args.nslots or None)
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device(f"cuda:{local_rank}"))  # gloo carries the host-side gathers of the checks
        MD.init(local_rank, args.grid, True, staged=args.staged, nslots=args.nslots or None)
    lib = M.grid._lib()
    assert lib.mfb_dist_nranks() == world and lib.mfb_dist_rank() == rank
    if args.borrow:   # the per-product path under memory pressure
        assert lib.mfb_set_option(b"conv_own_accumulator", 0) == 0 and lib.mfb_set_option(b"sources_v2", 0) == 0
    v2_expected = not args.borrow and lib.mfb_dist_nslots() >= 5

    n, L, TOL = args.grid, 100.0, PC.TOL
    N3, L3 = (n,) * 3, (L,) * 3

    # ---- slab bookkeeping == fftw_mpi_local_size_3d_transposed for divisible sizes
    g0 = M.Grid_FFT(N3, L3)
    assert g0.is_distributed()
    assert (g0.local_0_start_, g0.local_0_size_) == (rank * n // world, n // world)
    assert (g0.local_1_start_, g0.local_1_size_) == (rank * n // world, n // world)

    # ---- FFT round trip against numpy
    rng = np.random.default_rng(3)
    x = rng.standard_normal(N3)
    g0.set_real(x)
    g0.FourierTransformForward()
    ref = np.fft.rfftn(x) / np.sqrt(float(n) ** 3)
    assert PC.relerr(MD.gather_kspace(g0), ref) < TOL
    assert g0.size(0) == n // world and g0.size(1) == n
    g0.FourierTransformBackward()
    assert PC.relerr(MD.gather_real(g0), x) < TOL

    # ---- reductions (Allreduce inside), DC removal in both spaces, shift_field
    assert abs(g0.mean() - x.mean()) < 1e-13 and abs(g0.std() - x.std()) < 1e-12
    g0.zero_DC_mode()
    assert abs(MD.gather_real(g0).mean()) < 1e-14
    g0.set_real(x)
    g0.shift_field((0.5, 0.25, 0.5))
    gs = O.GridSpec(n, L)
    assert PC.relerr(MD.gather_real(g0), O.shift_field_r(gs, x, (0.5, 0.25, 0.5))) < TOL

    # ---- named products, Nyquist-carrying input (MPI pad semantics)
    g, W, Wc, phi = PC.noise_and_phi(n, L, 9, False)
    _, _, _, psi = PC.noise_and_phi(n, L, 10, False)
    cv = O.OrszagConvolver(g)
    P, Q = M.Grid_FFT(N3, L3), M.Grid_FFT(N3, L3)
    out = M.Grid_FFT(N3, L3, True, M.kspace_id)
    P.set_kspace(phi)
    Q.set_kspace(psi)
    C = M.OrszagConvolver(N3, L3)
    op = M.op
    ref = cv.convolve_Hessians(phi, (0, 1), psi, (1, 2))
    C.convolve_Hessians(P, (0, 1), Q, (1, 2), op.assign_to(out))
    got = MD.gather_kspace(out)
    assert PC.relerr(got, ref) < TOL
    h = n // 2
    assert not got[h].any() and not got[:, h].any() and not got[:, :, h].any()
    ref = ref + cv.convolve_SumOfHessians(phi, (0, 0), psi, (1, 1), (2, 2))
    C.convolve_SumOfHessians(P, (0, 0), Q, (1, 1), (2, 2), op.add_to(out))
    ref = ref + (-5.0 / 7.0) * cv.convolve_Hessians3(phi, (0, 0), psi, (0, 2), phi, (2, 2))
    C.convolve_Hessians(P, (0, 0), Q, (0, 2), P, (2, 2), op.multiply_add_to(out, -5.0 / 7.0))
    assert PC.relerr(MD.gather_kspace(out), ref) < TOL
    ref = cv.convolve_Gradient_and_Hessian(phi, 1, psi, (0, 2))
    C.convolve_Gradient_and_Hessian(P, (1,), Q, (0, 2), op.assign_to(out))
    assert PC.relerr(MD.gather_kspace(out), ref) < TOL
    # ---- the same on the NaiveConvolver (include/convolution.hh:296-426): no padding, Nyquist planes take part
    nv = O.NaiveConvolver(g)
    Cn = M.NaiveConvolver(N3, L3)
    ref = nv.convolve_Hessians(phi, (0, 1), psi, (1, 2))
    Cn.convolve_Hessians(P, (0, 1), Q, (1, 2), op.assign_to(out))
    assert PC.relerr(MD.gather_kspace(out), ref) < TOL
    ref = ref + nv.convolve_SumOfHessians(phi, (0, 0), psi, (1, 1), (2, 2))
    Cn.convolve_SumOfHessians(P, (0, 0), Q, (1, 1), (2, 2), op.add_to(out))
    ref = ref + (-5.0 / 7.0) * nv.convolve_Hessians3(phi, (0, 0), psi, (0, 2), phi, (2, 2))
    Cn.convolve_Hessians(P, (0, 0), Q, (0, 2), P, (2, 2), op.multiply_add_to(out, -5.0 / 7.0))
    assert PC.relerr(MD.gather_kspace(out), ref) < TOL
    ref = nv.convolve_Gradients(phi, 0, psi, 2)
    Cn.convolve_Gradients(P, (0,), Q, (2,), op.assign_to(out))
    assert PC.relerr(MD.gather_kspace(out), ref) < TOL
    del Cn

    # ---- device noise fill in the NGenIC layout: every rank fills its own k-slab
    P.fill_noise_ngenic_layout(4242)
    assert PC.relerr(MD.gather_kspace(P), O.ngenic_layout_noise(n, 4242)) < 1e-13
    P.fill_noise_ngenic(12345)          # the reference's own generator: one RANLUX stream per column of the local k-slab
    assert np.abs(MD.gather_kspace(P) - O.ngenic_noise(n, 12345)).max() < 1e-14

    # ---- masked sc lattice on slabs (include/particle_generator.hh:171-260): closed-form particle ranks per slab
    mask = O.particle_type_mask(True, 3)
    xr, vr = g.bwd(phi), g.bwd(psi)
    l0, o0 = n // world, rank * (n // world)
    for lidx in (0, 1):
        parts = M.Particles("masked", lidx, True, True, True, lidx, P, mask=mask)
        for d in range(3):
            P.set_kspace(phi)
            P.FourierTransformBackward()
            parts.set_positions(d, L, P)
            Q.set_kspace(psi)
            Q.FourierTransformBackward()
            parts.set_velocities(d, Q)
        P.set_kspace(phi)
        P.FourierTransformBackward()
        parts.set_masses(2.5, P)
        host = parts.download()
        sel = O._mask_select(n, mask, lidx)
        before, here = int(sel[:o0].sum()), int(sel[o0:o0 + l0].sum())
        assert len(host["ids"]) == here
        assert np.array_equal(host["ids"], O.masked_ids(n, mask, lidx, lidx, True)[before:before + here])
        for d in range(3):
            assert np.abs(host["pos"][:, d] - O.masked_positions(g, xr, d, mask, lidx, L)[before:before + here]).max() < 1e-12 * L
            assert PC.relerr(host["vel"][:, d], O.masked_velocities(g, vr, mask, lidx)[before:before + here]) < TOL
        assert PC.relerr(host["mass"], O.masked_masses(g, xr, mask, lidx, 2.5)[before:before + here]) < TOL
        del parts
    del P, Q, out, C

    # ---- complex grids on slabs: c2c transforms (exchange in two z halves) against numpy
    rngc = np.random.default_rng(8)
    xc = rngc.standard_normal(N3) + 1j * rngc.standard_normal(N3)
    Gc = M.Grid_FFT_complex(N3, L3)
    assert Gc.distributed_ and Gc.local_0_size_ == n // world
    Gc.set(xc)
    Gc.FourierTransformForward()
    refc = np.fft.fftn(xc) / np.sqrt(float(n) ** 3)
    loc = Gc.get()                                        # [ky_loc][kx][kz]
    o1, l1 = Gc.local_1_start_, Gc.local_1_size_
    assert PC.relerr(loc, np.transpose(refc[:, o1:o1 + l1], (1, 0, 2))) < TOL
    Gc.apply_gradient(1)
    Gc.FourierTransformBackward()
    kyv = g.k1d[None, :, None]
    refg = np.fft.ifftn(1j * kyv * refc) * np.sqrt(float(n) ** 3)
    o0, l0 = Gc.local_0_start_, Gc.local_0_size_
    assert PC.relerr(Gc.get(), refg[o0:o0 + l0]) < TOL
    del Gc

    # ---- the PPT / Eulerian branch on slabs (src/ic_generator.cc:685-803).  hbar follows the reference's MPI loop: every rank
    # leaves out the x neighbour of its last local plane before the MAX reduction
    Dp, vfac = 0.0257, 396.69
    for order in (1, 2):
        gp, Wp, Wcp, phip = PC.noise_and_phi(n, L, 31)
        amp = O.power_law_amplitude()
        amp_bc = lambda k: 2e-3 * amp(k) * k * k
        Csp, vu = 0.8433, 1.7
        pots = O.lpt_potentials(gp, phip, order, Dp)
        ref, hbar_ref = O.ppt_eulerian(gp, Wcp, pots, order, Dp, vfac, amp_bc, Csp, vu, nranks=world)
        pl = M.LPTPipeline(M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad="sc", Dplus0=Dp, vfac=vfac))
        pl.set_noise_k(Wp)
        pl.condition_noise()
        pl.compute_phi1(O.radial_table(gp, amp))
        pl.compute_sources()
        got = {}
        hbar = pl.eulerian_fields(O.radial_table(gp, amp_bc), Csp, vu, sink=lambda name, grid: got.__setitem__(name, MD.gather_real(grid)))[1]
        assert abs(hbar - hbar_ref) < 1e-12 * hbar_ref, (hbar, hbar_ref)
        for k in ("density", "vx", "vy", "vz"):
            err = np.abs(got[k] - ref[k]).max() / np.abs(ref[k]).max()
            assert err < 1e-9, (k, err)
        del pl

    # ---- whole pipelines vs the oracle
    for order, lattice, real64 in [(1, "sc", True), (3, args.lattice, True), (2, "sc", False)]:
        Dp, vfac = 0.0257, 396.69
        lunit, vunit = L, L / np.sqrt(1 / 51.0)
        g, W, Wc, phi = PC.noise_and_phi(n, L, 21)
        amp = O.power_law_amplitude()
        ref = O.run_pipeline(n, L, Wc, amp, order, Dp, vfac, lattice, lunit, vunit, real64, ids64=True)
        cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad=lattice, Dplus0=Dp, vfac=vfac,
                           lunit=lunit, vunit=vunit, real64=real64, ids64=True)
        pl = M.LPTPipeline(cfg)
        if args.borrow and order > 1:
            assert not pl.conv.grouped()
            pl.make_particles()
            assert pl.conv.grouped()      # the slab's particle arrays now serve as the product-group accumulator
        pl.set_noise_k(W)
        pl.condition_noise()
        parts = pl.run_device(O.radial_table(g, amp), particles=True)
        if order > 1:
            # shared x passes (3 exchanges per potential) + x-plane chunks + multi-product z pass when the block has the slots
            info = pl.conv.sources_info()
            assert info["path"] == (2 if v2_expected or (order == 2 and not args.borrow and lib.mfb_dist_nslots() >= 4) else 1), info
            if info["path"] == 2 and info["slots"] >= 7:
                assert info["x_passes"] == ((10 if info["slots"] >= 8 else 11) if order > 2 else 3), info
            if info["path"] == 2 and order > 2 and info["slots"] >= 8:
                # shared-field loops (z_fields_kernel): the first stage with 8 slots, the last one with 10 (a slot per target)
                assert info["loops"] == (3 if info["slots"] >= 10 else 4), info
        assert PC.relerr(MD.gather_kspace(pl.phi), ref["pots"]["phi"]) < TOL
        if order > 2:
            assert PC.relerr(MD.gather_kspace(pl.phi3), ref["pots"]["phi3"]) < 10 * TOL
            for d in range(3):
                assert PC.relerr(MD.gather_kspace(pl.A3[d]), ref["pots"]["A3"][d]) < 10 * TOL
        host = parts.download()
        # reference ordering: shift iload occupies [iload*N^3, (iload+1)*N^3); this rank holds planes off0..off0+l0
        nsh = 1 << O.LATTICE_TYPES[lattice]
        ncell, l0, o0 = n ** 3, n // world, rank * (n // world)
        sel = np.concatenate([np.arange(s * ncell + o0 * n * n, s * ncell + (o0 + l0) * n * n) for s in range(nsh)])
        assert np.array_equal(host["ids"], ref["ids"][sel])
        maxpsi = max(np.abs(ref["fields"]["d" + c]).max() for c in "xyz")
        maxv = max(np.abs(ref["fields"]["v" + c]).max() for c in "xyz")
        if real64:
            assert np.abs(host["pos"] - ref["pos"][sel]).max() < 1e-11 * maxpsi
            assert np.abs(host["vel"] - ref["vel"][sel]).max() < 1e-11 * maxv
        else:
            assert (host["pos"] == ref["pos"][sel]).mean() > 0.999
            assert np.allclose(host["vel"], ref["vel"][sel], rtol=1e-6, atol=1e-6 * maxv)
        if order > 2 and info["path"] == 2:
            # the z-blocked exchange layout (contiguous peer stores per tile) and the plain one carry the same numbers
            keep = [pl.phi2.get_kspace(), pl.phi3.get_kspace(), pl.A3[1].get_kspace()]
            assert lib.mfb_set_option(b"exchange_blocked", 0) == 0
            try:
                pl.set_noise_k(W)
                pl.condition_noise()
                pl.compute_phi1(O.radial_table(g, amp))
                pl.compute_sources()
            finally:
                lib.mfb_set_option(b"exchange_blocked", 1)
            for a_, b_ in zip(keep, [pl.phi2.get_kspace(), pl.phi3.get_kspace(), pl.A3[1].get_kspace()]):
                assert np.array_equal(a_, b_)
        del pl, parts
    MD.check()
    nex = lib.mfb_dist_exchanges()
    dist.barrier()
    if rank == 0:
        print(f"dist_worker ok: world={world} grid={n} exchanges={nex} slots={lib.mfb_dist_nslots()} mode="
              f"{'emu+gloo staged' if args.emu else ('nccl staged' if args.staged else 'peer-memory fused')}")
    MD.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
