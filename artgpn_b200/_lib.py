"""ctypes binding of ``libartgpn_b200.so`` (C ABI declared in ``include/artgpn_b200.h``).

There is no CPU fallback: if the library is missing, or no CUDA device can be used, every
compute entry point raises.  ``symbols()`` lists what ``include/artgpn_b200.h`` declares so the
CPU test-suite can check that the built library exports all of it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGPN_LIB") or os.path.join(_HERE, "libartgpn_b200.so")   # AGPN_LIB: build-variant A/B tests

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)
c_void_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/artgpn_b200.h
_SIGNATURES = {
    "agpn_version": (ctypes.c_int, []),
    "agpn_last_error": (ctypes.c_char_p, []),
    "agpn_kernel_npars": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "agpn_mean_npars": (ctypes.c_int, [ctypes.c_int]),
    "agpn_create": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, c_void_p, c_void_p, c_void_p]),
    "agpn_destroy": (ctypes.c_int, [c_void_p]),
    "agpn_set_structure": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int]),
    "agpn_theta_dim": (ctypes.c_int, [c_void_p]),
    "agpn_set_exact": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "agpn_loglike_batch": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, ctypes.c_int, c_void_p,
                                          ctypes.c_int, c_void_p, c_void_p]),
    "agpn_comm_unique_id": (ctypes.c_int, [c_void_p]),
    "agpn_comm_init": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int]),
    "agpn_comm_rank": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "agpn_loglike_batch_sharded": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, ctypes.c_int, c_void_p,
                                                  ctypes.c_int, c_void_p, c_void_p]),
    "agpn_loglike_grad": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "agpn_covariance_block": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, c_void_p, ctypes.c_int,
                                             c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p,
                                             ctypes.c_int, c_void_p]),
    "agpn_mean_table": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, c_void_p]),
    "agpn_mean_eval": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int,
                                      c_void_p]),
    "agpn_predict": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "agpn_predict_slice": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_double, ctypes.c_double, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "agpn_predict_cov": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_double,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "agpn_kernel_eval": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p]),
    "agpn_potrf_batched": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p]),
    "agpn_mvn_sample": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_int, c_void_p, c_void_p, c_void_p]),
    "agpn_dmma_peak": (ctypes.c_int, [ctypes.c_int, c_void_p]),
    "agpn_i8_peak": (ctypes.c_int, [ctypes.c_int, c_void_p]),
    "agpn_launch_count": (ctypes.c_longlong, []),
    "agpn_set_profiling": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "agpn_phase_ms": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "agpn_update_path": (ctypes.c_int, [c_void_p]),
    "agpn_predict_path": (ctypes.c_int, [c_void_p]),
    "agpn_timeline_dump": (ctypes.c_int, [c_void_p, ctypes.c_char_p]),
    "agpn_skip_stats": (ctypes.c_int, [c_void_p, c_void_p]),
    "agpn_mega_prof": (ctypes.c_int, [c_void_p, c_void_p]),
    "agpn_slice_ms": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
}

_lib = None


class AgpnError(RuntimeError):
    """A CUDA / usage error reported by libartgpn_b200.so."""


def symbols():
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AgpnError(
            "libartgpn_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C artgpn_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().agpn_last_error()
        raise AgpnError(msg.decode("utf-8", "replace") if msg else "libartgpn_b200 error %d" % rc)


def ptr(a):
    """Raw address of a C-contiguous float64 / int32 numpy array (or None)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)
