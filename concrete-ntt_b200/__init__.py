"""concrete-ntt_b200 -- B200-native batched negacyclic NTT, drop-in for the hot path of concrete-ntt 0.2.0.

Module layout mirrors the reference crate (src/lib.rs:83-111):

    prime32.Plan, prime64.Plan, prime64.Solinas
    native32.Plan32, native64.Plan32, native128.Plan32
    native_binary32.Plan32, native_binary64.Plan32, native_binary128.Plan32
    product.Plan, product.FwdMode, product.InvMode
    prime.is_prime64, prime.largest_prime_in_arithmetic_progression64

Import with ``importlib.import_module("concrete-ntt_b200")`` (the directory name is not an identifier).
"""
import ctypes as _C
import types as _types

from . import _lib
from ._lib import CnttError, LibraryMissing, ReferencePanic  # noqa: F401
from . import plans as _plans
from . import shard  # noqa: F401  (batch sharding across GPUs; no collective on the data path)


def _module(name, **attrs):
    m = _types.ModuleType(__name__ + "." + name)
    m.__dict__.update(attrs)
    return m


class _Solinas:
    """prime64::Solinas (src/prime64/generic_solinas.rs:36-40)"""
    P = 0xFFFFFFFF00000001


def _native(bits, binary):
    return type("Plan32", (_plans._NativePlan,), {"_bits": bits, "_binary": binary,
                "__doc__": "native%s%d::Plan32" % ("_binary" if binary else "", bits)})


def _is_prime64(n):
    return bool(_lib.lib().cntt_is_prime64(n))


def _largest_prime(factor, offset, lo, hi):
    out = _C.c_uint64()
    ok = _lib.lib().cntt_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, _C.byref(out))
    return out.value if ok else None


def _find_primitive_root64(p, degree):
    out = _C.c_uint64()
    ok = _lib.lib().cntt_find_primitive_root64(p, degree, _C.byref(out))
    return out.value if ok else None


prime32 = _module("prime32", Plan=type("Plan", (_plans.Plan32Prime,), {"__doc__": "prime32::Plan"}))
prime64 = _module("prime64", Plan=type("Plan", (_plans.Plan64Prime,), {"__doc__": "prime64::Plan"}), Solinas=_Solinas)
native32 = _module("native32", Plan32=_native(32, False))
native64 = _module("native64", Plan32=_native(64, False))
native128 = _module("native128", Plan32=_native(128, False))
native_binary32 = _module("native_binary32", Plan32=_native(32, True))
native_binary64 = _module("native_binary64", Plan32=_native(64, True))
native_binary128 = _module("native_binary128", Plan32=_native(128, True))
product = _module("product", Plan=type("Plan", (_plans.ProductPlan,), {"__doc__": "product::Plan"}),
                  FwdMode=_plans.FwdMode, InvMode=_plans.InvMode)
prime = _module("prime", is_prime64=_is_prime64, largest_prime_in_arithmetic_progression64=_largest_prime)
roots = _module("roots", find_primitive_root64=_find_primitive_root64)


def version():
    return _lib.lib().cntt_version().decode()
