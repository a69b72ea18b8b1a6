"""ctypes binding of libcntt_b200.so -- the C ABI declared in include/cntt_b200.h.

The shared library is built in-tree (``__graft_entry__.build()`` / ``make -C concrete-ntt_b200/csrc``).
There is no fallback of any kind: if the library is missing, or no CUDA device is usable, calls fail.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CNTT_B200_LIB", os.path.join(_HERE, "libcntt_b200.so"))  # override: experiments only

OK, INVALID_SIZE, INVALID_MODULUS, NO_ROOT, LENGTH_MISMATCH, CUDA_ERROR, NULL_POINTER, UNSUPPORTED, PANIC_MODULUS, MISALIGNED = range(10)

_vp, _sz, _int = C.c_void_p, C.c_size_t, C.c_int
_u32, _u64 = C.c_uint32, C.c_uint64

# name -> (restype, argtypes).  Keep in sync with include/cntt_b200.h (tests/test_abi.py checks both ways).
SIGNATURES = {
    "cntt_status_string": (C.c_char_p, [_int]),
    "cntt_last_cuda_error": (C.c_char_p, []),
    "cntt_version": (C.c_char_p, []),
    "cntt_is_prime64": (_int, [_u64]),
    "cntt_largest_prime_in_arithmetic_progression64": (_int, [_u64, _u64, _u64, _u64, C.POINTER(_u64)]),
    "cntt_find_primitive_root64": (_int, [_u64, _u64, C.POINTER(_u64)]),
    "cntt_host_alloc": (_int, [C.POINTER(_vp), _sz]),
    "cntt_host_free": (_int, [_vp]),
}
for _b, _w in (("32", _u32), ("64", _u64)):
    _p = "cntt_prime" + _b
    SIGNATURES.update({
        _p + "_plan_new": (_int, [_sz, _w, _int, C.POINTER(_vp)]),
        _p + "_plan_free": (None, [_vp]),
        _p + "_ntt_size": (_sz, [_vp]),
        _p + "_modulus": (_w, [_vp]),
        _p + "_fwd": (_int, [_vp, _vp, _sz, _vp]),
        _p + "_inv": (_int, [_vp, _vp, _sz, _vp]),
        _p + "_mul_assign_normalize": (_int, [_vp, _vp, _vp, _sz, _vp]),
        _p + "_normalize": (_int, [_vp, _vp, _sz, _vp]),
        _p + "_mul_accumulate": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
        _p + "_fwd_host": (_int, [_vp, _vp, _sz, _sz]),
        _p + "_inv_host": (_int, [_vp, _vp, _sz, _sz]),
        _p + "_fwd_inv_host": (_int, [_vp, _vp, _sz, _sz]),
        _p + "_fwd_host_multi": (_int, [C.POINTER(_vp), _int, _vp, _sz, _sz]),
        _p + "_inv_host_multi": (_int, [C.POINTER(_vp), _int, _vp, _sz, _sz]),
        _p + "_fwd_inv_host_multi": (_int, [C.POINTER(_vp), _int, _vp, _sz, _sz]),
        _p + "_mul_assign_normalize_host": (_int, [_vp, _vp, _vp, _sz]),
        _p + "_normalize_host": (_int, [_vp, _vp, _sz]),
        _p + "_mul_accumulate_host": (_int, [_vp, _vp, _vp, _vp, _sz]),
    })
SIGNATURES.update({
    "cntt_native_plan_new": (_int, [_sz, _int, _int, _int, C.POINTER(_vp)]),
    "cntt_native_plan_new_ext": (_int, [_sz, _int, _int, _int, C.POINTER(_vp)]),
    "cntt_native_plan_free": (None, [_vp]),
    "cntt_native_ntt_size": (_sz, [_vp]),
    "cntt_native_num_primes": (_int, [_vp]),
    "cntt_native_prime": (_u32, [_vp, _int]),
    "cntt_native_fwd": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_native_fwd_binary": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_native_inv": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_native_fwd_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_native_fwd_binary_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_native_inv_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_native_polymul": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "cntt_native_polymul_ntt_rhs": (_int, [_vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    "cntt_native_polymul_host": (_int, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "cntt_native_polymul_host_multi": (_int, [C.POINTER(_vp), _int, _vp, _vp, _vp, _sz, _sz]),
    "cntt_native52_plan_new": (_int, [_sz, _int, _int, _int, C.POINTER(_vp)]),
    "cntt_native52_plan_free": (None, [_vp]),
    "cntt_native52_ntt_size": (_sz, [_vp]),
    "cntt_native52_num_primes": (_int, [_vp]),
    "cntt_native52_prime": (_u64, [_vp, _int]),
    "cntt_native52_fwd": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_native52_fwd_binary": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_native52_inv": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_native52_fwd_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_native52_fwd_binary_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_native52_inv_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_native52_polymul": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "cntt_native52_polymul_host": (_int, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "cntt_product_plan_new": (_int, [_sz, _u64, C.POINTER(_u64), _sz, _int, C.POINTER(_vp)]),
    "cntt_product_plan_free": (None, [_vp]),
    "cntt_product_ntt_size": (_sz, [_vp]),
    "cntt_product_modulus": (_u64, [_vp]),
    "cntt_product_ntt_domain_len": (_sz, [_vp]),
    "cntt_product_num_primes": (_int, [_vp, C.POINTER(_int), C.POINTER(_int)]),
    "cntt_product_prime": (_u64, [_vp, _int]),
    "cntt_product_fwd": (_int, [_vp, _vp, _vp, _int, _u64, _sz, _vp]),
    "cntt_product_inv": (_int, [_vp, _vp, _vp, _int, _sz, _vp]),
    "cntt_product_mul_assign_normalize": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cntt_product_normalize": (_int, [_vp, _vp, _sz, _vp]),
    "cntt_product_mul_accumulate": (_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "cntt_product_fwd_host": (_int, [_vp, _vp, _vp, _sz, _sz, _int, _u64, _sz]),
    "cntt_product_inv_host": (_int, [_vp, _vp, _vp, _sz, _sz, _int, _sz]),
    "cntt_product_mul_assign_normalize_host": (_int, [_vp, _vp, _vp, _sz, _sz]),
    "cntt_product_normalize_host": (_int, [_vp, _vp, _sz, _sz]),
    "cntt_product_mul_accumulate_host": (_int, [_vp, _vp, _vp, _vp, _sz, _sz]),
})

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """Load the product library.  Raises LibraryMissing (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class CnttError(RuntimeError):
    def __init__(self, status, where=""):
        l = lib()
        msg = l.cntt_status_string(status).decode()
        if status == CUDA_ERROR:
            msg += ": " + l.cntt_last_cuda_error().decode()
        super().__init__("%s%s" % (where + ": " if where else "", msg))
        self.status = status


class ReferencePanic(AssertionError):
    """Raised where the reference crate panics (assert_eq! on lengths, Div::new on p <= 1)."""


def check(status, where=""):
    if status == OK:
        return
    if status in (LENGTH_MISMATCH, PANIC_MODULUS):
        raise ReferencePanic(lib().cntt_status_string(status).decode())
    raise CnttError(status, where)
