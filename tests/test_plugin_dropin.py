"""CPU-only, build container only (needs the reference sources): "the existing plugins drop in untouched" as a tested fact.

The reference's plugin sources are compiled UNMODIFIED, where they lie under /root/reference/src/plugins, against a shadow
include directory (tests/cpp/shadow) in which grid_fft.hh, convolution.hh and particle_generator.hh resolve to the
GPU-backed drop-ins of include/mfb/.  Everything else (output_plugin.hh, random_plugin.hh, transfer_function_plugin.hh,
config_file.hh, particle_container.hh, HDF_IO.hh ...) is the reference's own header.  FFTW / HDF5 / GSL are not installed:
their headers come from oracle/refshim (inert hdf5.h, fftw3.h) and tests/cpp/gslshim (declarations only).
One plugin is also linked and RUN through the reference's own factory (transfer_eisenstein: no third-party calls).
Round 2c: every other reference source compiles too - src/ic_generator.cc (the numerical driver, host lambdas and all),
src/testing.cc, src/main.cc, the cosmology calculator and the factories; only src/grid_fft.cc is replaced."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
INC = ["-I" + os.path.join(ROOT, "tests", "cpp", "shadow"), "-I" + os.path.join(ROOT, "include"),
       "-I" + os.path.join(ROOT, "oracle", "refshim"), "-I" + os.path.join(ROOT, "tests", "cpp", "gslshim"),
       "-I" + os.path.join(REF, "include")]
FLAGS = ["-std=c++17", "-w", "-DMFB_HAVE_REFERENCE_GENERAL_HH", "-DUSE_HDF5"]

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "plugins")), reason="reference sources not mounted")

PLUGINS = ["output_generic", "output_gadget_hdf5", "output_swift", "output_grafic2", "output_arepo", "output_genericio",
           "transfer_eisenstein", "transfer_CAMB_file", "transfer_CLASS",
           "random_music", "random_music_wnoise_generator", "random_ngenic", "random_panphasia"]
# the host code around the path: the numerical driver itself (ic_generator::run with its host lambdas, the PPT branch on
# Grid_FFT<ccomplex_t>, the glass branch with the reference's own grid_interpolate.hh), the unit tests of src/testing.cc,
# main, the cosmology calculator and the plugin factories - every reference source except src/grid_fft.cc, which the library replaces
HOST_SOURCES = ["ic_generator", "testing", "main", "cosmology_calculator", "cosmology_parameters", "output_plugin", "random_plugin",
                "transfer_function_plugin", "logger"]


@pytest.mark.parametrize("plugin", PLUGINS)
def test_reference_plugin_compiles_against_the_dropin_headers(plugin, tmp_path):
    src = os.path.join(REF, "src", "plugins", plugin + ".cc")
    r = subprocess.run(["g++", "-c", "-o", str(tmp_path / (plugin + ".o"))] + FLAGS + INC + [src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # the object really was built against the drop-in: plugins that touch a grid reference its C ABI or its host mirror
    if plugin in ("output_generic", "output_grafic2", "random_music", "random_ngenic"):
        syms = subprocess.run(["nm", "-C", str(tmp_path / (plugin + ".o"))], capture_output=True, text=True).stdout
        assert "mfb_grid_download" in syms or "mfb_grid_upload" in syms, plugin


@pytest.mark.parametrize("source", HOST_SOURCES)
def test_reference_host_source_compiles_against_the_dropin_headers(source, tmp_path):
    src = os.path.join(REF, "src", source + ".cc")
    obj = str(tmp_path / (source + ".o"))
    r = subprocess.run(["g++", "-c", "-o", obj] + FLAGS + INC + [src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    if source == "ic_generator":
        # the driver's object binds the C ABI: real and complex grids, both convolvers' named products, the particle read-out
        syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
        for need in ("mfb_fft_forward", "mfb_cgrid_fft_forward", "mfb_conv_hessians", "mfb_conv_hessians3", "mfb_conv_sum_hessians",
                     "mfb_conv_diff_hessians", "mfb_grid_inverse_laplacian", "mfb_particles_set_positions", "mfb_particles_set_velocities"):
            assert "U " + need in syms, need
        assert "fftw_" not in syms      # nothing of the FFTW path is left in the numerical driver


def test_reference_transfer_plugin_links_and_runs(tmp_path):
    exe = str(tmp_path / "dropin_transfer")
    srcs = [os.path.join(ROOT, "tests", "cpp", "dropin_transfer.cc")] + [os.path.join(REF, "src", f) for f in (
        "plugins/transfer_eisenstein.cc", "transfer_function_plugin.cc", "cosmology_parameters.cc", "logger.cc")]
    r = subprocess.run(["g++", "-O1", "-o", exe] + FLAGS + INC + srcs, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, os.path.join(REF, "tests", "configs", "test_1lpt_sc_generic.conf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    tab = np.array([[float(x) for x in ln.split()] for ln in r.stdout.splitlines() if ln and ln[0].isdigit()])
    assert len(tab) > 10
    k, T = tab[:, 0], tab[:, 1]
    assert 0.9 < T[0] <= 1.0 + 1e-6            # T -> 1 for k -> 0 (Eisenstein & Hu, un-normalised)
    assert np.all(np.diff(T) < 0) and T[-1] < 1e-3
