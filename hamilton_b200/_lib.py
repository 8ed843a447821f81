"""Loads the C-ABI shared library (include/hamilton_b200.h).  Fails loudly if it is missing:
there is no Python or CPU fallback for any compute entry point."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HB_LIB_PATH: development knob (A/B of a saved build of this same library on one GPU box, profiles/exp/exp_r2_ab.py)
LIB_PATH = os.environ.get("HB_LIB_PATH") or os.path.join(_HERE, "lib", "libhamilton_b200.so")

# hb_status
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_COMPILE, ERR_TAPE, ERR_NUMERIC, ERR_UNSUPPORTED = range(8)
# layouts / memspaces / integrators
AOS, SOA = 0, 1
HOST, DEVICE = 0, 1
RK4, RKF45_GSL = 0, 1
FLAG_NOT_SPD, FLAG_NONFINITE, FLAG_STEP_FAILED = 1, 2, 4


class HbOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("_pad", C.c_int32), ("c", C.c_double)]


class HbTape(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_ops", C.c_int32), ("ops", C.POINTER(HbOp)),
                ("n_out", C.c_int32), ("outs", C.POINTER(C.c_int32))]


class HamiltonError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("hamilton_b200 status %d: %s" % (status, msg))
        self.status = status


class NoDeviceError(HamiltonError):
    pass


class NumericError(HamiltonError, ArithmeticError):
    """The analogue of the reference's `error` / hmatrix exception for a single Phase."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "hamilton_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C hamilton_b200/csrc -j`).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double
    L.hb_abi_version.restype = i32
    L.hb_last_error.restype = C.c_char_p
    L.hb_device_count.argtypes = [ip]
    L.hb_set_device.argtypes = [i32]
    L.hb_system_builtin.argtypes = [i32, dp, i32, C.POINTER(vp)]
    L.hb_system_from_tape.argtypes = [i32, i32, dp, C.POINTER(HbTape), C.POINTER(HbTape), i32, dp, i32, C.POINTER(vp)]
    L.hb_system_free.argtypes = [vp]
    L.hb_system_free.restype = None
    L.hb_system_dims.argtypes = [vp, ip, ip]
    L.hb_system_source.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.hb_system_source.restype = C.c_size_t
    L.hb_system_params.argtypes = [vp, dp, i32]
    L.hb_system_params.restype = i32
    L.hb_batch_ham_eqs.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp]
    L.hb_batch_step.argtypes = [vp, i32, dbl, i32, i64, i32, i32, vp, vp, vp, vp]
    L.hb_batch_evolve.argtypes = [vp, i32, i32, i64, i32, i32, vp, dp, i32, vp, vp, vp]
    L.hb_batch_to_phase.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    L.hb_batch_from_phase.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp]
    L.hb_batch_energies.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp]
    L.hb_batch_underlying_pos.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    L.hb_batch_init_random.argtypes = [vp, C.c_uint64, i64, i64, i32, dp, dp, vp, vp]
    L.hb_underlying_pos.argtypes = [vp, dp, dp]
    L.hb_pe.argtypes = [vp, dp, dp]
    for nm in ("hb_momenta", "hb_velocities"):
        getattr(L, nm).argtypes = [vp, dp, dp, dp]
    for nm in ("hb_ke_c", "hb_ke_p", "hb_lagrangian", "hb_hamiltonian"):
        getattr(L, nm).argtypes = [vp, dp, dp, dp]
    L.hb_ham_eqs.argtypes = [vp, dp, dp, dp, dp]
    L.hb_step_ham.argtypes = [vp, dbl, dp, dp, dp, dp]
    L.hb_evolve_ham.argtypes = [vp, dp, dp, dp, i32, dp]
    L.hb_step_ham_c.argtypes = [vp, dbl, dp, dp, dp, dp]
    L.hb_evolve_ham_c.argtypes = [vp, dp, dp, dp, i32, dp]
    if not hasattr(L, "hb_ensemble_create"):      # an older saved build loaded through HB_LIB_PATH (A/B runs only)
        _lib = L
        return L
    L.hb_ensemble_create.argtypes = [vp, i32, ip, i64, C.POINTER(vp)]
    L.hb_ensemble_free.argtypes = [vp]
    L.hb_ensemble_free.restype = None
    L.hb_ensemble_dims.argtypes = [vp, ip, C.POINTER(i64), C.POINTER(i64)]
    L.hb_ensemble_init_random.argtypes = [vp, C.c_uint64, dp, dp]
    L.hb_ensemble_upload.argtypes = [vp, vp]
    L.hb_ensemble_step.argtypes = [vp, i32, dbl, i32, i32, dp]
    L.hb_ensemble_gather.argtypes = [vp, vp, dp]
    L.hb_ensemble_shard.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64)]
    L.hb_ensemble_gathered.argtypes = [vp, i32, C.POINTER(vp)]
    L.hb_ensemble_flags.argtypes = [vp, vp]
    _lib = L
    return L


def check(status):
    if status == OK:
        return
    msg = lib().hb_last_error().decode("utf-8", "replace")
    if status == ERR_NO_DEVICE:
        raise NoDeviceError(status, msg)
    if status == ERR_NUMERIC:
        raise NumericError(status, msg)
    raise HamiltonError(status, msg)


# every symbol include/hamilton_b200.h declares (checked by the CPU test-suite)
ABI_SYMBOLS = [
    "hb_abi_version", "hb_last_error", "hb_device_count", "hb_set_device", "hb_system_builtin", "hb_system_from_tape",
    "hb_system_free", "hb_system_dims", "hb_system_source", "hb_system_params", "hb_batch_ham_eqs", "hb_batch_step", "hb_batch_evolve",
    "hb_batch_to_phase", "hb_batch_from_phase", "hb_batch_energies", "hb_batch_underlying_pos", "hb_batch_init_random",
    "hb_underlying_pos", "hb_pe", "hb_momenta", "hb_velocities", "hb_ke_c", "hb_ke_p", "hb_lagrangian", "hb_hamiltonian",
    "hb_ham_eqs", "hb_step_ham", "hb_evolve_ham", "hb_step_ham_c", "hb_evolve_ham_c",
    "hb_ensemble_create", "hb_ensemble_free", "hb_ensemble_dims", "hb_ensemble_init_random", "hb_ensemble_upload", "hb_ensemble_step",
    "hb_ensemble_gather", "hb_ensemble_shard", "hb_ensemble_gathered", "hb_ensemble_flags",
]
