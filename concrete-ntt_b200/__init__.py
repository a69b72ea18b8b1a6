"""concrete-ntt_b200 -- B200-native batched negacyclic NTT, drop-in for the hot path of concrete-ntt 0.2.0.

Module layout mirrors the reference crate (src/lib.rs:83-111):

    prime32.Plan, prime64.Plan, prime64.Solinas
    native32.Plan32, native64.Plan32, native128.Plan32
    native_binary32.Plan32, native_binary64.Plan32, native_binary128.Plan32
    product.Plan, product.FwdMode, product.InvMode
    prime.is_prime64, prime.largest_prime_in_arithmetic_progression64, prime.{mul,exp}_mod{32,64}
    fastdiv.Div32, fastdiv.Div64, roots.find_primitive_root64

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


def _native52(bits, binary):
    name = "%s%d::Plan52" % ("native_binary" if binary else "native", bits)
    return type("Plan52", (_plans._Native52Plan,), {"_bits": bits, "_binary": binary, "__doc__": name})


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


class _Div:
    """fastdiv::Div32 / Div64 (src/fastdiv.rs:28-150): plan-time helper, host only.  The reference precomputes a
    double-word reciprocal; the quotients and remainders it returns are the exact ones, which is all that is
    observable, so the mirror keeps the divisor and uses exact integer division.  `new` panics for 0 and 1
    like the reference's `assert!(divisor > 1)` (src/fastdiv.rs:49,99)."""
    _bits = 32

    def __init__(self, divisor):
        divisor = int(divisor)
        if not 1 < divisor < (1 << self._bits):
            raise ReferencePanic("assert!(divisor > 1)")
        self._d = divisor

    @classmethod
    def new(cls, divisor):
        return cls(divisor)

    def divisor(self):
        return self._d

    @staticmethod
    def div(n, d):
        return int(n) // d._d

    @staticmethod
    def rem(n, d):
        return int(n) % d._d


class _Div32(_Div):
    _bits = 32
    div_u64 = staticmethod(lambda n, d: int(n) // d._d)
    rem_u64 = staticmethod(lambda n, d: int(n) % d._d)


class _Div64(_Div):
    _bits = 64
    div_u128 = staticmethod(lambda n, d: int(n) // d._d)
    rem_u128 = staticmethod(lambda n, d: int(n) % d._d)


def _div_of(n, cls):
    return n if isinstance(n, _Div) else cls(n)


def _mul_mod32(n, x, y):
    """prime::mul_mod32 (src/prime.rs:4-6)"""
    return int(x) * int(y) % _div_of(n, _Div32)._d


def _mul_mod64(n, x, y):
    """prime::mul_mod64 (src/prime.rs:8-10)"""
    return int(x) * int(y) % _div_of(n, _Div64)._d


def _exp_mod32(n, base, power):
    """prime::exp_mod32 (src/prime.rs:12-29)"""
    return pow(int(base), int(power), _div_of(n, _Div32)._d)


def _exp_mod64(n, base, power):
    """prime::exp_mod64 (src/prime.rs:31-48)"""
    return pow(int(base), int(power), _div_of(n, _Div64)._d)


prime32 = _module("prime32", Plan=type("Plan", (_plans.Plan32Prime,), {"__doc__": "prime32::Plan"}))
prime64 = _module("prime64", Plan=type("Plan", (_plans.Plan64Prime,), {"__doc__": "prime64::Plan"}), Solinas=_Solinas)
native32 = _module("native32", Plan32=_native(32, False), Plan52=_native52(32, False))
native64 = _module("native64", Plan32=_native(64, False), Plan52=_native52(64, False))
native128 = _module("native128", Plan32=_native(128, False))
native_binary32 = _module("native_binary32", Plan32=_native(32, True), Plan52=_native52(32, True))
native_binary64 = _module("native_binary64", Plan32=_native(64, True), Plan52=_native52(64, True))
native_binary128 = _module("native_binary128", Plan32=_native(128, True))
product = _module("product", Plan=type("Plan", (_plans.ProductPlan,), {"__doc__": "product::Plan"}),
                  FwdMode=_plans.FwdMode, InvMode=_plans.InvMode)
prime = _module("prime", is_prime64=_is_prime64, largest_prime_in_arithmetic_progression64=_largest_prime,
                mul_mod32=_mul_mod32, mul_mod64=_mul_mod64, exp_mod32=_exp_mod32, exp_mod64=_exp_mod64)
fastdiv = _module("fastdiv", Div32=_Div32, Div64=_Div64)
roots = _module("roots", find_primitive_root64=_find_primitive_root64)


HostMulti = _plans.HostMulti  # extension: one host batch over several GPUs in one process


def version():
    return _lib.lib().cntt_version().decode()
