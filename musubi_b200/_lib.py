"""ctypes binding of libmusb200.so (the C ABI of include/musb200.h) and of the
host-side synthetic mesh generator libmusb200_mesh.so.

There is NO fallback: if the CUDA library is missing the import fails loudly,
and every compute entry point fails when no CUDA device is present.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# MUSB200_LIB: path of an alternative build of the same library (kernel experiments only)
_SO = os.environ.get("MUSB200_LIB") or os.path.join(_PKG, "libmusb200.so")
_MESH_SO = os.path.join(_PKG, "libmusb200_mesh.so")

c_int = ctypes.c_int
c_double = ctypes.c_double
c_void_p = ctypes.c_void_p
c_char_p = ctypes.c_char_p
P_INT = ctypes.POINTER(ctypes.c_int)
P_I32 = ctypes.POINTER(ctypes.c_int32)
P_I64 = ctypes.POINTER(ctypes.c_int64)
P_DBL = ctypes.POINTER(ctypes.c_double)
P_LL = ctypes.POINTER(ctypes.c_longlong)

# every symbol include/musb200.h declares: name -> (argtypes)
SIGNATURES = {
    "musb200_init": [c_int, c_int, c_int, c_void_p],
    "musb200_get_unique_id": [c_void_p],
    "musb200_finalize": [],
    "musb200_last_error": [c_char_p, c_int],
    "musb200_device_count": [P_INT],
    "musb200_scheme_select": [c_char_p, c_char_p, c_char_p, c_char_p, P_INT, P_INT, P_INT],
    "musb200_level_create": [c_int] * 9 + [P_I32, P_I64, P_I64],
    "musb200_level_create_cube": [c_int, c_int, c_int, c_int],
    "musb200_state_init_equilibrium": [c_int],
    "musb200_level_destroy": [c_int],
    "musb200_neigh_download": [c_int, P_I32],
    "musb200_state_upload": [c_int, c_int, c_void_p],
    "musb200_state_download": [c_int, c_int, c_void_p],
    "musb200_set_now_next": [c_int, c_int, c_int],
    "musb200_get_now_next": [c_int, P_INT, P_INT],
    "musb200_aux_upload": [c_int, c_void_p],
    "musb200_aux_download": [c_int, c_void_p],
    "musb200_aux_probe": [c_int, c_int, P_DBL],
    "musb200_state_copy_next_to_now": [c_int],
    "musb200_set_relaxation": [c_int, c_int, c_int, P_DBL, c_double, c_double, c_double],
    "musb200_set_viscosity": [c_int, P_DBL, c_double],
    "musb200_pdf_serialize": [c_int, P_I64, P_I32, P_DBL],
    "musb200_pdf_unserialize": [c_int, P_I64, P_I32, P_DBL],
    "musb200_source_force": [c_int, c_int, c_int, P_I32, P_DBL, c_int],
    "musb200_set_species": [c_int, c_int, c_int, c_double, c_double],
    "musb200_set_transport_velocity": [c_int, c_int, P_DBL, c_int],
    "musb200_scheme_bind": [c_int],
    "musb200_couple_transport_velocity": [c_int, c_int, c_int],
    "musb200_bc_elembuffer": [c_int, c_int, P_I32],
    "musb200_bc_register": [c_int, c_int, c_int, c_int, P_I32, P_I32, P_I32, P_I32],
    "musb200_bc_set_values": [c_int, c_int, c_int, c_void_p],
    "musb200_bc_register_elems": [c_int, c_int, c_int, P_I32, P_I32, P_I32, c_int, P_I32, P_I32],
    "musb200_comm_register": [c_int, c_int, c_int, c_int, P_I32, P_I32, P_I32],
    "musb200_intp_register": [c_int, c_int, c_int, c_int, P_I32, P_I32, P_I32, P_DBL, P_I32, c_int,
                              P_I32, P_DBL, P_DBL],
    "musb200_step": [c_int, c_int, c_int],
    "musb200_step_schemes": [c_int, P_INT, c_int, c_int, c_int],
    "musb200_set_aux_every_step": [c_int],
    "musb200_fill_helper_elements": [c_int, c_int],
    "musb200_synchronize": [],
    "musb200_reduce": [c_int, P_DBL, P_DBL, P_INT],
    "musb200_compute_host": [c_int, c_int, c_int, P_DBL, P_DBL, P_DBL, P_I32, c_int, c_int, P_DBL,
                             c_double, c_double],
    "musb200_timers": [P_DBL, P_DBL, P_DBL, P_DBL],
    "musb200_timers_reset": [],
    "musb200_launch_count": [P_LL],
    "musb200_set_overlap": [c_int],
    "musb200_set_sweep_wait": [c_int],
    "musb200_set_exchange_timeout": [c_double],
    "musb200_set_graphs": [c_int],
    "musb200_set_fused_bc": [c_int],
    "musb200_set_intp_tiled": [c_int],
    "musb200_p2p_export": [c_int, c_void_p],
    "musb200_p2p_connect": [c_int, c_int, P_I32, c_void_p, P_I32, P_I32],
    "musb200_p2p_enable": [c_int, c_int],
    "musb200_set_fused_push": [c_int],
    "musb200_event_mark": [c_int],
    "musb200_event_elapsed": [P_DBL],
    "musb200_set_profiling": [c_int],
    "musb200_host_alloc": [ctypes.c_size_t, ctypes.POINTER(c_void_p)],
    "musb200_host_free": [c_void_p],
}


class Musb200Error(RuntimeError):
    """non-zero return of a libmusb200 entry point (the Fortran shim calls tem_abort)."""

    def __init__(self, code, msg):
        super().__init__("libmusb200 error %d: %s" % (code, msg))
        self.code = code


def _load(path):
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). musubi_b200 has no CPU fallback." % path)
    return ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)


lib = _load(_SO)
for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError = header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = c_int

mesh = _load(_MESH_SO)
mesh.musb200_mesh_box_create.restype = c_void_p
mesh.musb200_mesh_box_create.argtypes = [c_int] * 7
mesh.musb200_mesh_destroy.argtypes = [c_void_p]
mesh.musb200_mesh_destroy.restype = None
mesh.musb200_mesh_info.argtypes = [c_void_p, P_I64]
for _n in ("total", "property"):
    getattr(mesh, "musb200_mesh_" + _n).argtypes = [c_void_p, P_I64]
for _n in ("nghelems", "neigh", "bc_elembuffer"):
    getattr(mesh, "musb200_mesh_" + _n).argtypes = [c_void_p, P_I32]
mesh.musb200_mesh_comm.argtypes = [c_void_p, c_int, P_I32, P_I32, P_I32, P_I32, P_I32]
mesh.musb200_mesh_bc_info.argtypes = [c_void_p, c_int, P_I32]
mesh.musb200_mesh_bc_lists.argtypes = [c_void_p, c_int, P_I32, P_I32, P_I32, P_I32, P_I32]
mesh.musb200_mesh_bc_elem_lists.argtypes = [c_void_p, c_int, P_I32, P_I32, P_I32, P_I32, P_I32]
mesh.musb200_mesh_bary.argtypes = [c_void_p, c_double, c_double, c_double, c_double, P_DBL]


def last_error():
    buf = ctypes.create_string_buffer(1024)
    lib.musb200_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise Musb200Error(rc, last_error())


def ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None
