"""ctypes binding of csrc/liblfengine.so (C ABI declared in include/lf_engine.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Loading fails loudly when the
shared object is missing; creating a context fails loudly when there is no CUDA device.  There is no
CPU fallback anywhere in this package.
"""
import ctypes as C
import os

LF_MAX_FIELDS = 16
LF_MODEL_FREE, LF_MODEL_FIXED, LF_MODEL_Z = 0, 1, 2
LF_PREC_F64, LF_PREC_F32 = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LF_ENGINE_LIB', os.path.join(_HERE, 'csrc', 'liblfengine.so'))

#: every symbol include/lf_engine.h declares (checked by tests/test_abi.py)
EXPORTS = ['lf_ndim', 'lf_create', 'lf_destroy', 'lf_set_sources', 'lf_set_grid', 'lf_set_quadrature_share', 'lf_set_prior_gate', 'lf_set_compressed_sources',
           'lf_lnprob_batch', 'lf_lnprob_batch_device', 'lf_last_call_info', 'lf_veff_bin', 'lf_bin_weights', 'lf_boot_bin', 'lf_boot_bin_device',
           'lf_fp64_peak', 'lf_mufu_peak', 'lf_last_kernel_ms', 'lf_sampler_run', 'lf_sampler_last_ms', 'lf_cosmo_distances', 'lf_interp_linear', 'lf_device_count', 'lf_peer_buffer_create', 'lf_peer_buffer_connect',
           'lf_allreduce_device', 'lf_peer_status', 'lf_peer_reset', 'lf_peer_set_timeout', 'lf_veff_set_sample', 'lf_veff_bin_resident',
           'lf_veff_get_phi', 'lf_veff_set_volume_table', 'lf_veff_volumes', 'lf_omega_sources', 'lf_set_walker_sharding', 'lf_boot_mt_set_state', 'lf_boot_mt_get_state', 'lf_boot_bin_mt', 'lf_last_error', 'lf_version']


class LFCosmology(C.Structure):
    """Mirror of ``lf_cosmology`` (include/lf_engine.h)."""
    _fields_ = [(n, C.c_double) for n in ('H0', 'Om0', 'Ode0', 'Or0', 'Ok0', 'panel')] + \
               [('gl_x', C.c_double * 8), ('gl_w', C.c_double * 8)]


class LFConfig(C.Structure):
    _fields_ = [('model', C.c_int32), ('precision', C.c_int32), ('device', C.c_int32), ('nfields', C.c_int32),
                ('size_ln', C.c_int32), ('fix_sch_al', C.c_int32), ('fixed_prior_ok', C.c_int32),
                ('force_literal', C.c_int32), ('fcmin', C.c_double), ('sch_al', C.c_double),
                ('Lstar_lims', C.c_double * 2), ('phistar_lims', C.c_double * 2), ('sch_al_lims', C.c_double * 2),
                ('Flim_lims', C.c_double * 2), ('alpha_lims', C.c_double * 2), ('z_pivots', C.c_double * 3)]


class EngineError(RuntimeError):
    pass


_lib = None


def load():
    """Load liblfengine.so once and declare the argument types of its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EngineError("CUDA engine library not built: %s is missing -- run `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a).  lumfuncmcmc_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, dp, ip, i64 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int64
    lib.lf_last_error.restype = C.c_char_p
    lib.lf_version.restype = C.c_char_p
    lib.lf_ndim.argtypes = [vp]
    lib.lf_create.argtypes = [C.POINTER(vp), C.POINTER(LFConfig)]
    lib.lf_destroy.argtypes = [vp]
    lib.lf_destroy.restype = None
    lib.lf_set_sources.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    lib.lf_set_grid.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.lf_set_quadrature_share.argtypes = [vp, C.c_int32, C.c_int32]
    lib.lf_set_prior_gate.argtypes = [vp, C.c_int32]
    lib.lf_lnprob_batch.argtypes = [vp, vp, i64, vp]
    lib.lf_lnprob_batch_device.argtypes = [vp, vp, i64, vp, vp]
    lib.lf_last_call_info.argtypes = [vp, ip, ip]
    lib.lf_veff_bin.argtypes = [vp, i64, vp, vp, vp, C.c_int32, vp, C.c_double, C.c_double, C.c_double, C.c_double,
                                vp, vp, vp, C.c_int32, vp, vp, vp]
    lib.lf_bin_weights.argtypes = [vp, i64, vp, vp, vp, C.c_int32, vp, vp]
    lib.lf_boot_bin.argtypes = [vp, vp, vp, vp]
    lib.lf_boot_bin_device.argtypes = [vp, C.c_uint64, i64, vp, vp]
    lib.lf_fp64_peak.argtypes = [vp, C.c_int32, dp, dp]
    lib.lf_mufu_peak.argtypes = [vp, C.c_int32, dp, dp]
    lib.lf_last_kernel_ms.argtypes = [vp, dp]
    lib.lf_set_compressed_sources.argtypes = [vp, i64, vp, vp, vp, C.c_double]
    lib.lf_sampler_run.argtypes = [vp, vp, i64, i64, C.c_uint64, C.c_double, i64, vp, vp, vp, vp, vp]
    lib.lf_sampler_last_ms.argtypes = [vp, dp]
    lib.lf_set_walker_sharding.argtypes = [vp, C.c_int32]
    lib.lf_cosmo_distances.argtypes = [C.c_int32, vp, vp, i64, i64, vp, vp, vp]
    lib.lf_interp_linear.argtypes = [C.c_int32, i64, vp, vp, i64, vp, vp]
    lib.lf_omega_sources.argtypes = [C.c_int32, i64, vp, vp, vp, C.c_int32, vp, vp, C.c_double, C.c_double, i64, vp, vp, vp]
    lib.lf_peer_buffer_create.argtypes = [vp, C.c_int32, C.c_int32, i64, vp]
    lib.lf_peer_buffer_connect.argtypes = [vp, vp]
    lib.lf_allreduce_device.argtypes = [vp, vp, i64, vp]
    lib.lf_peer_status.argtypes = [vp, vp]
    lib.lf_peer_reset.argtypes = [vp]
    lib.lf_peer_set_timeout.argtypes = [vp, C.c_double]
    lib.lf_veff_set_sample.argtypes = [vp, i64, vp, vp, vp, C.c_int32]
    lib.lf_veff_bin_resident.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, vp, C.c_int32,
                                         vp, vp, vp]
    lib.lf_veff_get_phi.argtypes = [vp, vp]
    lib.lf_boot_mt_set_state.argtypes = [vp, vp, C.c_int32]
    lib.lf_boot_mt_get_state.argtypes = [vp, vp, vp]
    lib.lf_boot_bin_mt.argtypes = [vp, vp, vp]
    lib.lf_veff_set_volume_table.argtypes = [vp, vp, vp, i64, i64, vp, vp]
    lib.lf_veff_volumes.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp]
    _lib = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        raise EngineError(lib.lf_last_error().decode('utf-8', 'replace'))
