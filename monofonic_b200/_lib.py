"""ctypes loader for libmfb200.so (the nvcc sm_100a build).  There is NO fallback: if the
library is missing or no CUDA device is visible, importing users get a loud error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmfb200.so")
_lib = None


class MfbError(RuntimeError):
    pass


# mfb_exchange_fn (include/mfb200.h): void cb(const void* send, void* recv, size_t chunk_bytes, void* user)
EXCHANGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def _declare(lib):
    vp, i, d, sz, ull = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_ulonglong
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    vpp = C.POINTER(C.c_void_p)
    sig = {
        "mfb_init": (i, [i]),
        "mfb_device_count": (i, []),
        "mfb_device_mem_info": (i, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "mfb_finalize": (i, []),
        "mfb_last_error": (C.c_char_p, []),
        "mfb_build_info": (C.c_char_p, []),
        "mfb_synchronize": (i, []),
        "mfb_stream": (vp, []),
        "mfb_copy_stream": (vp, []),
        "mfb_kernel_launches": (C.c_longlong, []),
        "mfb_supported_fft_size": (i, [i]),
        "mfb_set_option": (i, [C.c_char_p, i]),
        "mfb_profile_enable": (i, [i]),
        "mfb_profile_reset": (i, []),
        "mfb_profile_num_roles": (i, []),
        "mfb_profile_role_name": (C.c_char_p, [i]),
        "mfb_profile_get": (i, [i, dp, C.POINTER(C.c_longlong), dp]),
        "mfb_dist_xbuf_bytes": (sz, [ip, i, i]),
        "mfb_dist_init": (i, [i, i, sz, vp]),
        "mfb_dist_init_slots": (i, [i, i, sz, i, vp]),
        "mfb_dist_connect": (i, [vp]),
        "mfb_dist_init_staged_slots": (i, [i, i, sz, i, EXCHANGE_FN, vp]),
        "mfb_dist_nslots": (i, []),
        "mfb_dist_init_staged": (i, [i, i, sz, EXCHANGE_FN, vp]),
        "mfb_dist_shutdown": (i, []),
        "mfb_dist_rank": (i, []),
        "mfb_dist_nranks": (i, []),
        "mfb_dist_exchanges": (C.c_longlong, []),
        "mfb_dist_check": (i, []),
        "mfb_grid_local": (i, [vp, ip]),
        "mfb_grid_create": (i, [vpp, ip, dp, i]),
        "mfb_grid_destroy": (i, [vp]),
        "mfb_grid_space": (i, [vp]),
        "mfb_grid_ntot": (sz, [vp]),
        "mfb_grid_device_ptr": (vp, [vp]),
        "mfb_grid_upload": (i, [vp, vp]),
        "mfb_noise_fill_ngenic": (i, [vp, C.c_ulonglong]),
        "mfb_noise_fill_ngenic_ranlxd": (i, [vp, C.c_ulonglong]),
        "mfb_grid_stream_in": (i, [vp, vp]),
        "mfb_grid_stream_in_wait": (i, [vp]),
        "mfb_grid_download": (i, [vp, vp]),
        "mfb_fft_forward": (i, [vp, i]),
        "mfb_fft_backward": (i, [vp, i]),
        "mfb_grid_zero": (i, [vp]),
        "mfb_grid_copy": (i, [vp, vp]),
        "mfb_grid_scale": (i, [vp, d]),
        "mfb_grid_axpy": (i, [vp, d, vp, d]),
        "mfb_grid_affine": (i, [vp, d, d]),
        "mfb_grid_zero_dc": (i, [vp]),
        "mfb_grid_inverse_laplacian": (i, [vp]),
        "mfb_grid_laplacian": (i, [vp, i]),
        "mfb_grid_shift_field": (i, [vp, dp, i]),
        "mfb_grid_mean": (i, [vp, dp]),
        "mfb_grid_std": (i, [vp, dp]),
        "mfb_grid_2norm": (i, [vp, dp]),
        "mfb_grid_max_neighbour_diff": (i, [vp, dp]),
        "mfb_fourier_resample": (i, [vp, vp]),
        "mfb_noise_condition": (i, [vp, d, i, i, i]),
        "mfb_apply_radial_table": (i, [vp, vp, vp, sz, i]),
        "mfb_add_tides": (i, [vp, vp, dp]),
        "mfb_lpt_assemble": (i, [vp, i, vp, vp, vp, vpp, d, d, d, d, d, vp, vp, sz, d]),
        "mfb_lpt_assemble_lg": (i, [vp, i, vp, vp, vp, vpp, d, d, d, d, d, vp, vp, sz, d, vp, i, i]),
        "mfb_gradient_create_plt": (i, [vpp, i, i, d, vpp]),
        "mfb_gradient_destroy": (i, [vp]),
        "mfb_density_from_phi": (i, [vp, vp]),
        "mfb_cgrid_create": (i, [vpp, ip, dp]),
        "mfb_cgrid_destroy": (i, [vp]),
        "mfb_cgrid_space": (i, [vp]),
        "mfb_cgrid_nelem": (sz, [vp]),
        "mfb_cgrid_local": (i, [vp, ip]),
        "mfb_cgrid_upload": (i, [vp, vp, i]),
        "mfb_cgrid_download": (i, [vp, vp]),
        "mfb_cgrid_copy": (i, [vp, vp]),
        "mfb_cgrid_fft_forward": (i, [vp, i]),
        "mfb_cgrid_fft_backward": (i, [vp, i]),
        "mfb_cgrid_gradient": (i, [vp, i]),
        "mfb_ppt_sqrt_density": (i, [vp, d]),
        "mfb_ppt_init_psi": (i, [vp, vp, vp, vp, d, d]),
        "mfb_ppt_drift": (i, [vp, d, d]),
        "mfb_ppt_kick": (i, [vp, vp, d, d]),
        "mfb_ppt_density": (i, [vp, vp]),
        "mfb_ppt_velocity": (i, [vp, vp, vp, vp, d, d]),
        "mfb_conv_create": (i, [vpp, ip, dp]),
        "mfb_conv_create_naive": (i, [vpp, ip, dp]),
        "mfb_conv_destroy": (i, [vp]),
        "mfb_conv_workspace_bytes": (sz, [vp]),
        "mfb_conv_grouped": (i, [vp]),
        "mfb_conv_borrow_accumulator": (i, [vp, vp]),
        "mfb_conv_sources_info": (i, [vp, C.POINTER(C.c_longlong)]),
        "mfb_conv_hessians": (i, [vp, vp, ip, vp, ip, vp, i, d]),
        "mfb_conv_hessians3": (i, [vp, vp, ip, vp, ip, vp, ip, vp, i, d]),
        "mfb_conv_sum_hessians": (i, [vp, vp, ip, vp, ip, ip, vp, i, d]),
        "mfb_conv_diff_hessians": (i, [vp, vp, ip, vp, ip, ip, vp, i, d]),
        "mfb_conv_gradients": (i, [vp, vp, i, vp, i, vp, i, d]),
        "mfb_conv_gradient_hessian": (i, [vp, vp, i, vp, ip, vp, i, d]),
        "mfb_conv_multiply": (i, [vp, vp, vp, vp, i, d]),
        "mfb_conv_generic": (i, [vp, vp, vp, vp, i, d]),
        "mfb_conv_generic3": (i, [vp, vp, vp, vp, vp, i, d]),
        "mfb_lpt_sources": (i, [vp, i, d, vp, vp, vp, vpp, dp]),
        "mfb_particles_create": (i, [vpp, ip, i, i, i, i]),
        "mfb_particles_create_masked": (i, [vpp, ip, i, ip, i, i, i]),
        "mfb_particles_create_glass": (i, [vpp, ip, vp, sz, C.c_ulonglong, i, i]),
        "mfb_particles_destroy": (i, [vp]),
        "mfb_particles_count": (sz, [vp]),
        "mfb_particles_set_ids": (i, [vp, ull]),
        "mfb_particles_set_positions": (i, [vp, i, i, d, vp]),
        "mfb_particles_set_velocities": (i, [vp, i, i, vp]),
        "mfb_particles_set_masses": (i, [vp, i, d, vp]),
        "mfb_particles_download": (i, [vp, vp, vp, vp, vp]),
        "mfb_particles_stream_out": (i, [vp, i, vp]),
        "mfb_particles_stream_out_range": (i, [vp, i, sz, sz, vp]),
        "mfb_particles_stream_wait": (i, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED = None


def load(path: str | None = None):
    """dlopen libmfb200.so and declare every entry point of include/mfb200.h."""
    global _lib, EXPORTED
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise MfbError(
            f"{p} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  monofonic_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    EXPORTED = _declare(lib)
    if path is None:
        _lib = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or _lib
        raise MfbError(lib.mfb_last_error().decode())
