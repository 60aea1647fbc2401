"""ctypes binding of libffthom_b200.so (the C ABI declared in include/ffthom_b200.h).

There is no CPU fallback: if the shared object is missing or a call fails, an
exception is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libffthom_b200.so')

c_i64 = C.c_int64
c_dbl = C.c_double
c_int = C.c_int
c_vp = C.c_void_p
p_i64 = C.POINTER(C.c_int64)
p_dbl = C.POINTER(C.c_double)
p_int = C.POINTER(C.c_int)


class FhError(RuntimeError):
    pass


class fh_green(C.Structure):
    """Mirror of `struct fh_green` (include/ffthom_b200.h)."""
    _fields_ = [('kind', C.c_int32), ('dim', C.c_int32),
                ('N', C.c_int64*3), ('band', C.c_int64*3), ('Y', C.c_double*3),
                ('c0', c_dbl), ('cI', c_dbl), ('cS', c_dbl), ('cH', c_dbl), ('cL', c_dbl), ('cW', c_dbl),
                ('scale', c_dbl)]


# name -> (restype, argtypes); device pointers travel as void* (integers from tensor.data_ptr())
_SIGNATURES = {
    'fh_init': (c_int, [c_int]),
    'fh_set_stream': (c_int, [c_vp]),
    'fh_sync': (c_int, []),
    'fh_last_error': (C.c_char_p, []),
    'fh_version': (c_int, []),
    'fh_device_info': (c_int, [p_int, p_int, p_int]),
    'fh_download': (c_int, [c_vp, c_vp, c_i64]),
    'fh_plan_create': (c_int, [C.POINTER(c_vp), c_int, p_i64]),
    'fh_plan_destroy': (c_int, [c_vp]),
    'fh_plan_factors': (c_int, [c_vp, c_int, p_int, p_int]),
    'fh_rfftn': (c_int, [c_vp, c_vp, c_vp, c_i64]),
    'fh_irfftn': (c_int, [c_vp, c_vp, c_vp, c_i64, c_dbl, c_vp]),
    'fh_axpby': (c_int, [c_i64, c_dbl, c_vp, c_dbl, c_vp, c_vp]),
    'fh_add_scalar': (c_int, [c_i64, c_vp, c_dbl, c_int, c_vp]),
    'fh_add_comp': (c_int, [c_int, c_i64, c_vp, p_dbl]),
    'fh_dot': (c_int, [c_i64, c_vp, c_vp, p_dbl]),
    'fh_dot_rspec': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, p_dbl]),
    'fh_convert': (c_int, [c_i64, c_vp, c_int, c_vp, c_int]),
    'fh_asum': (c_int, [c_i64, c_vp, c_int, p_dbl]),
    'fh_amax': (c_int, [c_i64, c_vp, c_int, p_dbl]),
    'fh_sum_comp': (c_int, [c_int, c_i64, c_vp, p_dbl]),
    'fh_poke': (c_int, [c_vp, c_i64, p_dbl, c_i64]),
    'fh_peek': (c_int, [c_vp, c_i64, p_dbl, c_i64]),
    'fh_memset0': (c_int, [c_vp, c_i64]),
    'fh_copy': (c_int, [c_vp, c_vp, c_i64]),
    'fh_gather_comps': (c_int, [c_i64, c_int, p_int, c_vp, c_vp]),
    'fh_mul21': (c_int, [c_int, c_i64, c_int, c_vp, c_int, c_vp, c_int, c_vp]),
    'fh_hadamard': (c_int, [c_i64, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_int, c_vp]),
    'fh_contract_first': (c_int, [c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    'fh_inv_dxd': (c_int, [c_int, c_i64, c_vp, c_vp]),
    'fh_assemble_AH': (c_int, [c_int, c_int, c_i64, c_vp, C.POINTER(c_vp), p_dbl]),
    'fh_topologies': (c_int, [c_int, p_i64, c_vp, p_dbl, c_int, p_int, p_dbl, p_dbl, c_vp, p_int]),
    'fh_combine_phases': (c_int, [c_int, c_int, c_i64, p_dbl, c_vp, c_vp]),
    'fh_sep_product': (c_int, [c_int, p_i64, c_vp, c_int, c_vp, c_vp]),
    'fh_gather_periodic': (c_int, [c_int, p_i64, p_i64, p_i64, c_int, c_vp, c_vp]),
    'fh_spec_remap': (c_int, [c_int, p_i64, c_int, p_i64, c_int, c_i64, c_dbl, c_int, c_vp, c_vp]),
    'fh_roll': (c_int, [c_int, p_i64, p_i64, c_int, c_i64, c_vp, c_vp]),
    'fh_grad': (c_int, [c_int, p_i64, p_dbl, c_int, c_int, c_vp, c_vp]),
    'fh_div': (c_int, [c_int, p_i64, p_dbl, c_int, c_int, c_vp, c_vp]),
    'fh_potential': (c_int, [c_int, p_i64, p_dbl, c_int, c_int, c_vp, c_vp]),
    'fh_green_apply': (c_int, [C.POINTER(fh_green), c_int, c_int, c_vp, c_vp]),
    'fh_green_materialize': (c_int, [C.POINTER(fh_green), c_int, c_vp]),
    'fh_green4_materialize': (c_int, [c_int, c_int, p_i64, p_dbl, c_int, c_vp]),
    'fh_ga_work_doubles': (c_i64, [c_vp, c_int]),
    'fh_ga_create': (c_int, [C.POINTER(c_vp), c_vp, c_int, c_vp, c_int, C.POINTER(fh_green), c_vp]),
    'fh_ga_slab_work_doubles': (c_i64, [c_vp, c_int, c_int, c_int]),
    'fh_ga_create_slab': (c_int, [C.POINTER(c_vp), c_vp, c_int, c_vp, c_int, C.POINTER(fh_green), c_vp, c_int, c_int,
                                  c_int]),
    'fh_ga_buffers': (c_int, [c_vp, C.POINTER(c_vp), C.POINTER(c_vp), p_int]),
    'fh_ga_last_dot': (c_int, [c_vp, p_dbl]),
    'fh_ga_slab_direct': (c_int, [c_vp, c_int, c_int, c_vp, c_vp]),
    'fh_ga_slab_peer': (c_int, [c_vp, c_int, c_int, C.POINTER(c_vp)]),
    'fh_ga_slab_push': (c_int, [c_vp, c_int, c_int, C.POINTER(c_vp), C.POINTER(c_vp)]),
    'fh_ga_slab_push_stage': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    'fh_ga_slab_kblock': (c_int, [c_vp, c_int, c_int, c_vp, c_vp]),
    'fh_ga_slab_kblock_info': (c_int, [c_vp, c_int, p_i64, p_i64, p_int, p_int]),
    'fh_ga_slab_kblock_stage': (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    'fh_ga_slab_stage': (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    'fh_cgd_init': (c_int, [c_vp, c_vp, c_vp]),
    'fh_cgd_update': (c_int, [c_vp, c_vp, c_vp]),
    'fh_ga_set_xacc': (c_int, [c_vp, c_vp]),
    'fh_ga_can_defer_x': (c_int, [c_vp]),
    'fh_cgd_update_r': (c_int, [c_vp, c_vp]),
    'fh_cgd_xflush': (c_int, [c_vp, c_vp, c_vp]),
    'fh_cgd_local_sum': (c_int, [c_vp, c_vp]),
    'fh_cgd_scal': (c_int, [c_vp, c_vp, c_int, p_dbl]),
    'fh_cg_xr_update': (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_dbl, p_dbl]),
    'fh_cg_p_update': (c_int, [c_i64, c_vp, c_vp, c_dbl]),
    'fh_ga_destroy': (c_int, [c_vp]),
    'fh_ga_apply': (c_int, [c_vp, c_vp, c_vp]),
    'fh_ga_config': (c_int, [c_vp, p_int, p_int, p_int]),
    'fh_ga_stage': (c_int, [c_vp, c_int, c_vp, c_vp]),
    'fh_cg_begin': (c_int, [c_vp, c_vp, c_vp, c_vp, p_dbl]),
    'fh_cg_steps': (c_int, [c_vp, c_vp, c_vp, c_dbl, c_i64, p_i64, p_dbl, p_dbl]),
    'fh_cg': (c_int, [c_vp, c_vp, c_vp, c_dbl, c_i64, c_vp, p_i64, p_dbl, p_dbl, c_i64]),
    'fh_richardson': (c_int, [c_vp, c_vp, c_vp, c_dbl, c_dbl, c_i64, c_vp, p_i64, p_dbl]),
    'fh_launch_count': (c_i64, []),
}

EXPORTS = tuple(sorted(_SIGNATURES))

_lib = None


def load():
    """Load the shared object (no CUDA call is made; safe without a GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FhError('%s not found: build it with `python -m ffthompy_b200.csrc.build` '
                          '(there is no CPU fallback)' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise FhError('libffthom_b200 error %d: %s' % (rc, load().fh_last_error().decode()))


def i64arr(vals):
    return (C.c_int64*len(vals))(*[int(v) for v in vals])


def dblarr(vals):
    return (C.c_double*len(vals))(*[float(v) for v in vals])


def intarr(vals):
    return (C.c_int*len(vals))(*[int(v) for v in vals])
