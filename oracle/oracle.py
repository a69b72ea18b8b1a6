"""ctypes binding of the CPU oracle (oracle/cntt_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (concrete-ntt_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(native=False):
    """Compile the oracle if the shared object is missing (gcc is in the image)."""
    name = "libcntt_oracle_native.so" if native else "libcntt_oracle.so"
    path = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in ("cntt_oracle.c", "cntt_simd.c", "cntt_simd32.c")]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "native" if native else "all"],
                              stdout=subprocess.DEVNULL)
    return path


def _load(native=False):
    lib = C.CDLL(build(native))
    vp = C.c_void_p
    sz = C.c_size_t
    sig = {
        "o_is_prime64": (C.c_int, [C.c_uint64]),
        "o_largest_prime_in_arithmetic_progression64": (C.c_int, [C.c_uint64] * 4 + [_u64p]),
        "o_find_primitive_root64": (C.c_int, [C.c_uint64, C.c_uint64, _u64p]),
        "o_exp_mod64": (C.c_uint64, [C.c_uint64] * 3),
        "o_mul_mod64": (C.c_uint64, [C.c_uint64] * 3),
        "o_plan32_new": (C.c_int, [sz, C.c_uint32, C.POINTER(vp)]),
        "o_plan32_free": (None, [vp]),
        "o_plan32_psi": (C.c_uint32, [vp]),
        "o_plan32_twid": (_u32p, [vp]),
        "o_plan32_inv_twid": (_u32p, [vp]),
        "o_plan32_fwd": (None, [vp, vp]),
        "o_plan32_inv": (None, [vp, vp]),
        "o_plan32_mul_assign_normalize": (None, [vp, vp, vp, sz]),
        "o_plan32_normalize": (None, [vp, vp, sz]),
        "o_plan32_mul_accumulate": (None, [vp, vp, vp, vp, sz]),
        "o_plan64_new": (C.c_int, [sz, C.c_uint64, C.POINTER(vp)]),
        "o_plan64_free": (None, [vp]),
        "o_plan64_psi": (C.c_uint64, [vp]),
        "o_plan64_twid": (_u64p, [vp]),
        "o_plan64_inv_twid": (_u64p, [vp]),
        "o_plan64_fwd": (None, [vp, vp]),
        "o_plan64_inv": (None, [vp, vp]),
        "o_plan64_mul_assign_normalize": (None, [vp, vp, vp, sz]),
        "o_plan64_normalize": (None, [vp, vp, sz]),
        "o_plan64_mul_accumulate": (None, [vp, vp, vp, vp, sz]),
        "o_primes32": (C.c_uint32, [C.c_int]),
        "o_reconstruct_32bit_01": (C.c_uint32, [C.c_uint32] * 2),
        "o_reconstruct_32bit_012_u32": (C.c_uint32, [C.c_uint32] * 3),
        "o_reconstruct_32bit_012_u64": (C.c_uint64, [C.c_uint32] * 3),
        "o_reconstruct_32bit_01234_u64": (C.c_uint64, [C.c_uint32] * 5),
        "o_reconstruct_32bit_01234_u128": (None, [C.c_uint32] * 5 + [_u64p]),
        "o_reconstruct_32bit_0123456789_u128": (None, [_u32p, _u64p]),
        "o_native_new": (C.c_int, [sz, C.c_int, C.c_int, C.POINTER(vp)]),
        "o_native_free": (None, [vp]),
        "o_native_nprimes": (C.c_int, [vp]),
        "o_native_fwd": (None, [vp, vp, vp]),
        "o_native_fwd_binary": (None, [vp, vp, vp]),
        "o_native_inv": (None, [vp, vp, vp]),
        "o_native_polymul": (None, [vp, vp, vp, vp]),
        "o_schoolbook32": (None, [sz, C.c_uint32, vp, vp, vp]),
        "o_schoolbook64": (None, [sz, C.c_uint64, vp, vp, vp]),
        "o_schoolbook128": (None, [sz, vp, vp, vp]),
        "o_direct_fwd64": (None, [sz, C.c_uint64, C.c_uint64, vp, vp]),
        "o_negacyclic_wrapping": (None, [sz, C.c_int, vp, vp, vp, C.c_int]),
        "o_max_threads": (C.c_int, []),
        "o_plan32_fwd_batch": (None, [vp, vp, sz, C.c_int]),
        "o_plan32_inv_batch": (None, [vp, vp, sz, C.c_int]),
        "o_plan64_fwd_batch": (None, [vp, vp, sz, C.c_int]),
        "o_plan64_inv_batch": (None, [vp, vp, sz, C.c_int]),
        "o_simd_isa": (C.c_int, []),
        "o_set_batch_isa": (None, [C.c_int]),
        "o_plan64_fwd_simd": (None, [vp, vp, C.c_int]),
        "o_plan64_inv_simd": (None, [vp, vp, C.c_int]),
        "o_plan32_fwd_simd": (None, [vp, vp, C.c_int]),
        "o_plan32_inv_simd": (None, [vp, vp, C.c_int]),
        "o_native_polymul_batch": (None, [vp, vp, vp, vp, sz, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib(native=False):
    global _LIB
    if native:
        return _load(True)
    if _LIB is None:
        _LIB = _load(False)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class ReferencePanic(Exception):
    """The reference would panic (Div32::new / Div64::new assert, src/fastdiv.rs:49,99)."""


def is_prime64(n):
    return bool(lib().o_is_prime64(n))


def largest_prime_in_arithmetic_progression64(factor, offset, lo, hi):
    out = C.c_uint64()
    ok = lib().o_largest_prime_in_arithmetic_progression64(factor, offset, lo, hi, C.byref(out))
    return out.value if ok else None


def find_primitive_root64(p, degree):
    out = C.c_uint64()
    ok = lib().o_find_primitive_root64(p, degree, C.byref(out))
    return out.value if ok else None


def exp_mod64(p, b, e):
    return lib().o_exp_mod64(p, b, e)


def primes32(i):
    return lib().o_primes32(i)


class _PrimePlan:
    _pre = None
    _dt = None

    def __init__(self, handle, n, p, L):
        self._h, self.n, self.p, self._L = handle, n, p, L

    @classmethod
    def try_new(cls, n, p, native=False):
        L = lib(native)
        h = C.c_void_p()
        st = getattr(L, cls._pre + "_new")(n, p, C.byref(h))
        if st == 2:
            raise ReferencePanic("divisor > 1")
        if st != 0:
            return None
        return cls(h, n, p, L)

    def __del__(self):
        try:
            getattr(self._L, self._pre + "_free")(self._h)
        except Exception:
            pass

    def ntt_size(self):
        return self.n

    def modulus(self):
        return self.p

    def psi(self):
        return getattr(self._L, self._pre + "_psi")(self._h)

    def twid(self):
        ptr = getattr(self._L, self._pre + "_twid")(self._h)
        return np.ctypeslib.as_array(ptr, shape=(self.n,)).copy()

    def inv_twid(self):
        ptr = getattr(self._L, self._pre + "_inv_twid")(self._h)
        return np.ctypeslib.as_array(ptr, shape=(self.n,)).copy()

    def _chk(self, *arrs):
        for a in arrs:
            assert a.dtype == self._dt and a.flags.c_contiguous

    def _apply(self, name, buf):
        """buf: (..., n) array transformed in place, one polynomial per row."""
        self._chk(buf)
        assert buf.shape[-1] == self.n, "assert_eq!(buf.len(), self.ntt_size())"
        flat = buf.reshape(-1, self.n)
        fn = getattr(self._L, self._pre + "_" + name)
        for row in flat:
            fn(self._h, _ptr(row))
        return buf

    def fwd(self, buf):
        return self._apply("fwd", buf)

    def inv(self, buf):
        return self._apply("inv", buf)

    def fwd_batch(self, buf, nthreads):
        self._chk(buf)
        getattr(self._L, self._pre + "_fwd_batch")(self._h, _ptr(buf), buf.size // self.n, nthreads)

    def inv_batch(self, buf, nthreads):
        self._chk(buf)
        getattr(self._L, self._pre + "_inv_batch")(self._h, _ptr(buf), buf.size // self.n, nthreads)

    def _simd(self, name, buf, isa):
        self._chk(buf)
        assert buf.shape[-1] == self.n
        fn = getattr(self._L, "%s_%s_simd" % (self._pre, name))
        for row in buf.reshape(-1, self.n):
            fn(self._h, _ptr(row), ISA_CODES[isa])
        return buf

    def fwd_simd(self, buf, isa="best"):
        """Plan::fwd through the AVX-512 / AVX2 port of the reference's vector path (cntt_simd.c: Solinas; cntt_simd32.c: 32-bit
        primes below 2^31); every other class runs the scalar path"""
        return self._simd("fwd", buf, isa)

    def inv_simd(self, buf, isa="best"):
        return self._simd("inv", buf, isa)

    def mul_assign_normalize(self, lhs, rhs):
        self._chk(lhs, rhs)
        getattr(self._L, self._pre + "_mul_assign_normalize")(self._h, _ptr(lhs), _ptr(rhs), min(lhs.size, rhs.size))
        return lhs

    def normalize(self, values):
        self._chk(values)
        getattr(self._L, self._pre + "_normalize")(self._h, _ptr(values), values.size)
        return values

    def mul_accumulate(self, acc, lhs, rhs):
        self._chk(acc, lhs, rhs)
        getattr(self._L, self._pre + "_mul_accumulate")(self._h, _ptr(acc), _ptr(lhs), _ptr(rhs),
                                                       min(acc.size, lhs.size, rhs.size))
        return acc


class Plan32(_PrimePlan):
    """prime32::Plan (src/prime32.rs:602-928)"""
    _pre = "o_plan32"
    _dt = np.dtype(np.uint32)


class Plan64(_PrimePlan):
    """prime64::Plan (src/prime64.rs:222-1129)"""
    _pre = "o_plan64"
    _dt = np.dtype(np.uint64)



ISA_CODES = {"best": -1, "scalar": 0, "avx2": 2, "avx512": 3}


def simd_isa(native=False):
    """widest ISA of cntt_simd.c this host can run: 'avx512', 'avx2' or 'scalar'"""
    return {3: "avx512", 2: "avx2"}.get(lib(native).o_simd_isa(), "scalar")


def set_batch_isa(isa, native=False):
    """ISA used by the *_batch entry points (the CPU-baseline legs of bench.py): 'best', 'scalar', 'avx2', 'avx512'"""
    lib(native).o_set_batch_isa(ISA_CODES[isa])


SOLINAS_P = 0xFFFFFFFF00000001


class Native:
    """native{32,64,128}::Plan32 / native_binary{32,64,128}::Plan32.

    Words are numpy uint32 / uint64; 128-bit words are uint64 arrays of shape (..., n, 2), little-endian limbs.
    Residue planes are a (nprimes, n) uint32 array.
    """

    def __init__(self, h, n, bits, binary, L):
        self._h, self.n, self.bits, self.binary, self._L = h, n, bits, binary, L
        self.nprimes = L.o_native_nprimes(h)

    @classmethod
    def try_new(cls, n, bits, binary=False, native=False):
        L = lib(native)
        h = C.c_void_p()
        if L.o_native_new(n, bits, int(binary), C.byref(h)) != 0:
            return None
        return cls(h, n, bits, binary, L)

    def __del__(self):
        try:
            self._L.o_native_free(self._h)
        except Exception:
            pass

    def word_dtype(self):
        return np.uint32 if self.bits == 32 else np.uint64

    def word_shape(self, *lead):
        return tuple(lead) + ((self.n,) if self.bits != 128 else (self.n, 2))

    def fwd(self, value):
        out = np.empty((self.nprimes, self.n), np.uint32)
        self._L.o_native_fwd(self._h, _ptr(np.ascontiguousarray(value)), _ptr(out))
        return out

    def fwd_binary(self, value):
        assert self.binary
        out = np.empty((self.nprimes, self.n), np.uint32)
        self._L.o_native_fwd_binary(self._h, _ptr(np.ascontiguousarray(value)), _ptr(out))
        return out

    def inv(self, mod_p):
        """Returns value; mod_p is clobbered like in the reference."""
        assert mod_p.dtype == np.uint32 and mod_p.shape == (self.nprimes, self.n) and mod_p.flags.c_contiguous
        out = np.empty(self.word_shape(), self.word_dtype())
        self._L.o_native_inv(self._h, _ptr(out), _ptr(mod_p))
        return out

    def negacyclic_polymul(self, lhs, rhs):
        lhs = np.ascontiguousarray(lhs)
        rhs = np.ascontiguousarray(rhs)
        assert lhs.shape == rhs.shape
        out = np.empty_like(lhs)
        per = self.word_shape()
        lead = lhs.shape[: lhs.ndim - len(per)]
        assert lhs.shape[len(lead):] == per, "assert_eq!(n, lhs.len())"
        l2 = lhs.reshape((-1,) + per)
        r2 = rhs.reshape((-1,) + per)
        o2 = out.reshape((-1,) + per)
        for i in range(l2.shape[0]):
            self._L.o_native_polymul(self._h, _ptr(o2[i]), _ptr(l2[i]), _ptr(r2[i]))
        return out

    def polymul_batch(self, prod, lhs, rhs, batch, nthreads):
        self._L.o_native_polymul_batch(self._h, _ptr(prod), _ptr(lhs), _ptr(rhs), batch, nthreads)


PRIMES52 = (0x3FFFFFE770001, 0x3FFFFFEB90001, 0x3FFFFFEC80001)   # primes52::P0..P2, src/lib.rs:600-602


class Native52:
    """native32 / native64 / native_binary32 / native_binary64 ::Plan52 (src/native64.rs:1072-1165, native32.rs:435-496,
    native_binary32.rs:266-330, native_binary64.rs:447-521).  The reference's reconstruct_52bit_* functions exist only as
    AVX-512-IFMA code (no scalar twin); every intermediate they produce is canonical (mul_mod52_avx512 ends with
    small_mod, native32.rs:96-107), so they compute the mixed-radix digits v_k of the CRT and
    v_0 + v_1 P_0 + v_2 P_0 P_1 - (v_top > P_top / 2 ? P_0 .. P_top : 0) wrapped to the word (native64.rs:796-828,
    native32.rs:237-252, native_binary32.rs:117-124).  Restated here with Python integers; one polynomial per call."""

    def __init__(self, n, bits, binary, plans):
        self.n, self.bits, self.binary, self.plans = n, bits, binary, plans
        self.primes = [pl.modulus() for pl in plans]

    @classmethod
    def try_new(cls, n, bits, binary=False):
        np_ = (1 if binary else 2) if bits == 32 else (2 if binary else 3)
        plans = [Plan64.try_new(n, p) for p in PRIMES52[:np_]]
        if any(pl is None for pl in plans):
            return None
        return cls(n, bits, binary, plans)

    def fwd(self, value, binary_copy=False):
        out = np.empty((len(self.plans), self.n), np.uint64)
        for k, (pl, p) in enumerate(zip(self.plans, self.primes)):
            v = value.astype(np.uint64)
            if self.bits == 64 and not binary_copy:       # native64.rs:1110-1112 `value % P_k`; 32-bit words are < P_k
                v = v % np.uint64(p)
            out[k] = pl.fwd(v.copy())
        return out

    def inv(self, mod_p):
        res = [pl.inv(mod_p[k].copy()) for k, pl in enumerate(self.plans)]
        mask = (1 << self.bits) - 1
        out = np.empty(self.n, np.uint32 if self.bits == 32 else np.uint64)
        P = self.primes
        for i in range(self.n):
            m = [int(r[i]) for r in res]
            v = [m[0]]
            partial, pre = m[0], 1
            for k in range(1, len(P)):
                pre *= P[k - 1]
                vk = (m[k] - partial) * pow(pre, -1, P[k]) % P[k]
                v.append(vk)
                partial += vk * pre
            if v[-1] > P[-1] // 2:
                partial -= pre * P[-1] if len(P) > 1 else P[0]
            out[i] = partial & mask
        return out


class Product:
    """product::Plan (src/product.rs:139-967), restated on top of the oracle's prime plans.  One polynomial per
    call like the reference: `standard` is a (n,) uint64 array, NTT-domain buffers are (ntt_domain_len,) uint64
    arrays in the reference's packed layout (u32 planes bit-cast into the front, then the u64 planes,
    product.rs:261-278).  Scalar paths only; the SIMD bodies of the 2 x u32 case compute the same values."""
    GENERIC = ("generic", 0)

    @staticmethod
    def bounded(bound):
        return ("bounded", int(bound))

    REPLACE, ACCUMULATE = 0, 1

    def __init__(self, n, modulus, primes, p32, p64):
        self.n, self.modulus, self.primes, self.p32, self.p64 = n, modulus, primes, p32, p64
        # modular_inverses (product.rs:205-226): inverse of every earlier prime modulo p_j
        self.minv = {(j, i): pow(primes[i] % primes[j], -1, primes[j]) for j in range(len(primes)) for i in range(j)}

    @classmethod
    def try_new(cls, n, modulus, factors):               # product.rs:152-251
        if n % 2 != 0:
            return None
        primes = sorted(int(f) for f in factors)
        prev = 0
        for f in primes:                                 # zeros / duplicates
            if f == prev:
                return None
            prev = f
        primes = [f for f in primes if f != 1]
        prod = 1
        for f in primes:
            prod *= f
            if prod >= 1 << 64:                          # checked_mul
                return None
        if prod != modulus:
            return None
        p32, p64 = [], []
        for f in primes:
            pl = Plan32.try_new(n, f) if f < 1 << 32 else Plan64.try_new(n, f)
            if pl is None:
                return None
            (p32 if f < 1 << 32 else p64).append(pl)
        return cls(n, modulus, primes, p32, p64)

    def ntt_size(self):
        return self.n

    def ntt_domain_len(self):                            # product.rs:265-274
        return (self.n // 2) * len(self.p32) + self.n * len(self.p64)

    def _split(self, ntt):
        n32 = (self.n // 2) * len(self.p32)
        return ntt[:n32].view(np.uint32), ntt[n32:]

    def fwd(self, ntt, standard, mode=GENERIC):          # product.rs:276-353
        assert standard.shape == (self.n,) and ntt.shape == (self.ntt_domain_len(),)
        n, c32, c64 = self.n, len(self.p32), len(self.p64)
        ntt32, ntt64 = self._split(ntt)
        if c32 == 0 and c64 == 1:
            ntt64[:] = standard
            self.p64[0].fwd(ntt64)
            return ntt
        if c32 == 1 and c64 == 0:
            ntt32[:] = standard.astype(np.uint32)        # `standard as u32`
            self.p32[0].fwd(ntt32)
            return ntt
        if c32 == 2 and c64 == 0:
            p0, p1, p = self.primes[0], self.primes[1], self.modulus
            if mode[0] == "bounded" and mode[1] < p0 and mode[1] < p1:
                positive = standard < np.uint64(p // 2)
                s32 = standard.astype(np.uint32)
                comp = (np.uint32(p & 0xFFFFFFFF) - s32).astype(np.uint32)
                with np.errstate(over="ignore"):
                    ntt32[:n] = np.where(positive, s32, np.uint32(p0) - comp)
                    ntt32[n:] = np.where(positive, s32, np.uint32(p1) - comp)
            else:
                ntt32[:n] = (standard % np.uint64(p0)).astype(np.uint32)
                ntt32[n:] = (standard % np.uint64(p1)).astype(np.uint32)
            self.p32[0].fwd(ntt32[:n])
            self.p32[1].fwd(ntt32[n:])
            return ntt
        for k, pl in enumerate(self.p32):
            ntt32[k * n:(k + 1) * n] = (standard % np.uint64(pl.p)).astype(np.uint32)
            pl.fwd(ntt32[k * n:(k + 1) * n])
        for k, pl in enumerate(self.p64):
            ntt64[k * n:(k + 1) * n] = standard % np.uint64(pl.p)
            pl.fwd(ntt64[k * n:(k + 1) * n])
        return ntt

    def inv(self, standard, ntt, mode=REPLACE):          # product.rs:355-880
        assert standard.shape == (self.n,) and ntt.shape == (self.ntt_domain_len(),)
        n, c32, c64 = self.n, len(self.p32), len(self.p64)
        ntt32, ntt64 = self._split(ntt)
        for k, pl in enumerate(self.p32):
            pl.inv(ntt32[k * n:(k + 1) * n])
        for k, pl in enumerate(self.p64):
            pl.inv(ntt64[k * n:(k + 1) * n])
        M64 = (1 << 64) - 1

        def add_mod_u64(p, a, b):                        # product.rs:83-91
            s = (a + b) & M64
            return (s - p) & M64 if (s >= p or a + b > M64) else s

        if c32 == 0 and c64 == 0:
            if mode == self.REPLACE:
                standard[:] = 0
            return standard
        if c32 == 1 and c64 == 0:
            p = self.primes[0]
            for i in range(n):
                if mode == self.REPLACE:
                    standard[i] = int(ntt32[i])
                else:                                    # add_mod_u32(p, *standard as u32, ntt) as u64
                    a, b = int(standard[i]) & 0xFFFFFFFF, int(ntt32[i])
                    s = (a + b) & 0xFFFFFFFF
                    standard[i] = (s - p) & 0xFFFFFFFF if (s >= p or a + b > 0xFFFFFFFF) else s
            return standard
        planes = [ntt32[k * n:(k + 1) * n] for k in range(c32)] + [ntt64[k * n:(k + 1) * n] for k in range(c64)]
        for i in range(n):                               # Knuth 4.3.2 mixed radix, product.rs:806-880
            v = []
            for j, pj in enumerate(self.primes):
                x = int(planes[j][i])
                for t in range(j):
                    diff = x - v[t] if x >= v[t] else x - v[t] + pj     # sub_mod
                    x = diff * self.minv[(j, t)] % pj
                v.append(x)
            acc = 0
            for j in reversed(range(len(v))):
                acc = (acc * self.primes[j] + v[j]) & M64
            standard[i] = acc if mode == self.REPLACE else add_mod_u64(self.modulus, int(standard[i]), acc)
        return standard

    def _pw(self, name, *bufs):                          # product.rs:884-967
        n = self.n
        sp = [self._split(b) for b in bufs]
        for k, pl in enumerate(self.p32):
            getattr(pl, name)(*[s[0][k * n:(k + 1) * n] for s in sp])
        for k, pl in enumerate(self.p64):
            getattr(pl, name)(*[s[1][k * n:(k + 1) * n] for s in sp])
        return bufs[0]

    def mul_assign_normalize(self, lhs, rhs):
        return self._pw("mul_assign_normalize", lhs, rhs)

    def normalize(self, values):
        return self._pw("normalize", values)

    def mul_accumulate(self, acc, lhs, rhs):
        return self._pw("mul_accumulate", acc, lhs, rhs)


def schoolbook32(p, lhs, rhs):
    out = np.empty_like(lhs)
    lib().o_schoolbook32(lhs.size, p, _ptr(lhs), _ptr(rhs), _ptr(out))
    return out


def schoolbook64(p, lhs, rhs):
    out = np.empty_like(lhs)
    lib().o_schoolbook64(lhs.size, p, _ptr(lhs), _ptr(rhs), _ptr(out))
    return out


def schoolbook128(lhs, rhs):
    out = np.empty_like(lhs)
    lib().o_schoolbook128(lhs.shape[0], _ptr(lhs), _ptr(rhs), _ptr(out))
    return out


def negacyclic_wrapping(bits, lhs, rhs, nthreads=None):
    """wrapping negacyclic product of one polynomial pair (any n; the specification of the extended plans)"""
    out = np.empty_like(lhs)
    n = lhs.shape[0]
    lib().o_negacyclic_wrapping(n, bits, _ptr(lhs), _ptr(rhs), _ptr(out), nthreads or max_threads())
    return out


def direct_fwd64(n, p, psi, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty(n, np.uint64)
    lib().o_direct_fwd64(n, p, psi, _ptr(a), _ptr(out))
    return out


def max_threads():
    return lib().o_max_threads()
