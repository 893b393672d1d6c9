"""ctypes binding of libkmcb200.so (include/kmc_b200.h).  No CPU fallback: a missing library or a
missing CUDA device raises, loudly, at first use."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KMC_LIB_PATH") or os.path.join(HERE, "libkmcb200.so")      # override: development builds only

KMC_HOST, KMC_DEVICE = 0, 1
KMC_F_AA, KMC_F_AS = 1, 2
KMC_E_CAPACITY, KMC_E_NOCANDIDATE, KMC_E_ITERCAP = -4, -5, -6


class KmcError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libkmcb200 error %d: %s" % (code, msg))
        self.code = code


class KmcParams(C.Structure):
    _fields_ = [
        ("L", C.c_double), ("bin_size", C.c_double), ("Dmin", C.c_double), ("Dmax", C.c_double),
        ("E_a", C.c_double), ("r_a", C.c_double), ("E_as", C.c_double), ("r_as", C.c_double),
        ("beta", C.c_double), ("omega", C.c_double), ("Rd", C.c_double),
        ("nbins_x", C.c_int32), ("nbins_y", C.c_int32), ("aa_mode", C.c_int32), ("precision", C.c_int32),
        ("hop_index_mode", C.c_int32), ("deposit_mode", C.c_int32),
    ]

    def key(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


class KmcRelaxStats(C.Structure):
    _fields_ = [("line_searches", C.c_int64), ("restarts", C.c_int64), ("pair_evals", C.c_int64),
                ("maxF", C.c_double), ("final_threshold", C.c_double), ("final_scale", C.c_int32),
                ("status", C.c_int32)]


_VP, _I64, _I32, _D = C.c_void_p, C.c_int64, C.c_int32, C.c_double
_PP = C.POINTER(KmcParams)

# every symbol include/kmc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "kmc_last_error": (C.c_char_p, []),
    "kmc_version": (C.c_int, []),
    "kmc_create": (C.c_int, [_PP, C.c_int, C.POINTER(_VP)]),
    "kmc_destroy": (C.c_int, [_VP]),
    "kmc_set_params": (C.c_int, [_VP, _PP]),
    "kmc_get_params": (C.c_int, [_VP, _PP]),
    "kmc_set_stream": (C.c_int, [_VP, _VP]),
    "kmc_synchronize": (C.c_int, [_VP]),
    "kmc_launch_count": (_I64, [_VP]),
    "kmc_timing_enable": (C.c_int, [_VP, C.c_int]),
    "kmc_timing_read": (C.c_int, [_VP, C.c_char_p, C.POINTER(_D), C.POINTER(_I64)]),
    "kmc_timing_reset": (C.c_int, [_VP]),
    "kmc_put_all_in_box": (C.c_int, [_VP, _VP, _I64, C.c_int]),
    "kmc_distances": (C.c_int, [_VP, _VP, _I64, _VP, _I64, _VP, C.c_int]),
    "kmc_displacement": (C.c_int, [_VP, _VP, _VP, _I64, _VP, C.c_int]),
    "kmc_bin_index": (C.c_int, [_VP, _VP, _I64, _VP, C.c_int]),
    "kmc_bins_create": (C.c_int, [_VP, _VP, _I64, C.c_int, C.POINTER(_VP)]),
    "kmc_bins_destroy": (C.c_int, [_VP, _VP]),
    "kmc_bins_count": (_I64, [_VP]),
    "kmc_bins_export": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int]),
    "kmc_nearby_atoms": (C.c_int, [_VP, _VP, _VP, _I64, _VP, _I64, _VP, C.c_int]),
    "kmc_neighbors": (C.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _I64, _VP, C.c_int]),
    "kmc_forces": (C.c_int, [_VP, _VP, _I64, _VP, C.c_int, _VP, C.c_int]),
    "kmc_relax_energy": (C.c_int, [_VP, _VP, _I64, _I64, _VP, _VP, C.c_int]),
    "kmc_line_energies": (C.c_int, [_VP, _VP, _VP, _I64, _I32, _I32, _VP, _VP, C.c_int]),
    "kmc_total_energy": (C.c_int, [_VP, _VP, _I64, _VP, _VP, C.c_int]),
    "kmc_probe": (C.c_int, [_VP, _VP, _I64, _VP, _I64, _VP, _VP, _VP, _VP, _VP, C.c_int]),
    "kmc_rates": (C.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int]),
    "kmc_select": (C.c_int, [_VP, _VP, _VP, _I64, _D, _VP, _VP, _VP, C.c_int]),
    "kmc_state_set": (C.c_int, [_VP, _VP, _I64, C.c_int]),
    "kmc_state_move": (C.c_int, [_VP, _VP, _I64, C.c_int]),
    "kmc_state_get": (C.c_int, [_VP, _VP, _I64, C.POINTER(_I64), C.c_int]),
    "kmc_state_count": (_I64, [_VP]),
    "kmc_relax": (C.c_int, [_VP, _VP, _I32, _D, _I64, C.POINTER(KmcRelaxStats)]),
    "kmc_set_relax_mode": (C.c_int, [_VP, C.c_int]),
    "kmc_set_select_mode": (C.c_int, [_VP, C.c_int]),
    "kmc_set_rate_mode": (C.c_int, [_VP, C.c_int, _D]),
    "kmc_rates_update": (C.c_int, [_VP, _VP, _D, _VP, _VP, _VP, C.POINTER(_I64), C.c_int]),
    "kmc_deposit_place": (C.c_int, [_VP, _VP, _D]),
    "kmc_hop_place": (C.c_int, [_VP, _VP, _I64, _D]),
    "kmc_event": (C.c_int, [_VP, _VP, _D, _D, _I64, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_D),
                            C.POINTER(KmcRelaxStats)]),
    "kmc_set_event_mode": (C.c_int, [_VP, C.c_int]),
    "kmc_event_enqueue": (C.c_int, [_VP, _VP, _D, _D, _I64]),
    "kmc_event_finish": (C.c_int, [_VP, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_D), C.POINTER(KmcRelaxStats)]),
    "kmc_event_place": (C.c_int, [_VP, _VP, _D, _D, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_D)]),
    "kmc_event_ready": (C.c_int, [_VP]),
    "kmc_events_batch": (C.c_int, [C.POINTER(_VP), C.POINTER(_VP), _I32, C.POINTER(_D), C.POINTER(_D), _I64, C.POINTER(_I32),
                                   C.POINTER(_I64), C.POINTER(_D), C.POINTER(KmcRelaxStats), C.POINTER(_I32)]),
    "kmc_domain_sweep": (C.c_int, [_VP, _VP, _I64, _I64, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int]),
    "kmc_domain_line_energies": (C.c_int, [_VP, _VP, _VP, _I64, _I64, _I32, _I32, _VP, _VP, C.c_int]),
    "kmc_fp64_peak": (C.c_int, [_VP, _I32, C.POINTER(_D), C.POINTER(_D)]),
    "kmc_relax_diag": (C.c_int, [_VP, C.POINTER(_I64)]),
    "kmc_barrier_bench": (C.c_int, [_VP, _I32, _I32, _I32, C.POINTER(_D), C.POINTER(_I32)]),
    "kmc_cells_stats": (C.c_int, [_VP, C.POINTER(_I64)]),
    "kmc_cells_check": (C.c_int, [_VP, C.POINTER(_I64)]),
}

_LIB = None


def load():
    """dlopen libkmcb200.so and attach the signatures (no CUDA call is made here)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "kmcislandgrowth_b200: %s is missing -- build it with `python -m kmcislandgrowth_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(lib, name)        # AttributeError here == header / library mismatch
        f.restype, f.argtypes = res, args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise KmcError(rc, load().kmc_last_error().decode("utf-8", "replace"))
