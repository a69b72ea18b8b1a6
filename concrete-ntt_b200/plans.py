"""Host-side mirror of the reference's Plan API (same names, argument meaning and error behaviour) over
the C ABI.  One polynomial per call becomes one *batch* per call: every buffer argument is an array whose
last dimension is the polynomial (``n`` words) and whose leading dimensions are the batch.

Buffers may be
  * torch CUDA tensors  -> device-resident path, in place, asynchronous on the current torch stream;
  * numpy arrays        -> host-slice path (pinned or pageable), staged H2D/D2H inside the library.
Reference: prime32.rs:627-928, prime64.rs:701-1129, native*.rs / native_binary*.rs `impl Plan32`.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ReferencePanic, check

try:  # torch is plumbing (device memory, streams); numpy-only use stays possible
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _Buf:
    """Pointer + shape facts of one argument."""

    def __init__(self, plan_device, x, itemsize, name):
        self.is_dev = _is_torch(x)
        if self.is_dev:
            if not x.is_cuda:
                raise TypeError("%s: torch tensors must live on a CUDA device (use numpy for host slices)" % name)
            if x.is_floating_point() or x.is_complex() or x.dtype == torch.bool:
                raise TypeError("%s: expected an integer tensor of %d-byte words, got %s" % (name, itemsize, x.dtype))
            if x.device.index != plan_device:
                raise TypeError("%s lives on cuda:%s but the plan was created for cuda:%s" % (name, x.device.index, plan_device))
            if not x.is_contiguous():
                raise ValueError("%s must be contiguous" % name)
            if x.element_size() != itemsize:
                raise TypeError("%s: expected %d-byte words" % (name, itemsize))
            if x.data_ptr() % 16:
                raise ValueError("%s: device batches must be 16-byte aligned (include/cntt_b200.h)" % name)
            self.ptr = x.data_ptr()
            self.shape = tuple(x.shape)
            self.device = x.device.index
            self.t = x
        else:
            if not isinstance(x, np.ndarray):
                raise TypeError("%s must be a numpy array or a torch CUDA tensor" % name)
            if not x.flags.c_contiguous:
                raise ValueError("%s must be C-contiguous" % name)
            if x.dtype.kind not in "ui":
                raise TypeError("%s: expected an integer array of %d-byte words, got %s" % (name, itemsize, x.dtype))
            if x.dtype.itemsize != itemsize:
                raise TypeError("%s: expected %d-byte words" % (name, itemsize))
            self.ptr = x.ctypes.data
            self.shape = x.shape
            self.device = None
        self.words = int(np.prod(self.shape)) if len(self.shape) else 1


class _PrimePlan:
    _bits = None

    def __init__(self, handle, n, p, device):
        self._h, self._n, self._p, self._device = handle, n, p, device

    @classmethod
    def try_new(cls, polynomial_size, modulus, device=0):
        """Plan::try_new -> Plan or None.  Raises ReferencePanic for modulus in {0, 1} (fastdiv.rs:49,99)."""
        l = _lib.lib()
        h = C.c_void_p()
        st = getattr(l, "cntt_prime%d_plan_new" % cls._bits)(polynomial_size, modulus, device, C.byref(h))
        if st in (_lib.INVALID_SIZE, _lib.INVALID_MODULUS, _lib.NO_ROOT):
            return None
        check(st, "try_new")
        return cls(h, polynomial_size, modulus, device)

    def __del__(self):
        try:
            if self._h:
                getattr(_lib.lib(), "cntt_prime%d_plan_free" % self._bits)(self._h)
                self._h = None
        except Exception:
            pass

    def _fn(self, name):
        return getattr(_lib.lib(), "cntt_prime%d_%s" % (self._bits, name))

    def ntt_size(self):
        return self._fn("ntt_size")(self._h)

    def modulus(self):
        return self._fn("modulus")(self._h)

    def _ntt(self, name, buf):
        b = _Buf(self._device, buf, self._bits // 8, "buf")
        if len(b.shape) == 0 or b.shape[-1] != self._n:
            raise ReferencePanic("assert_eq!(buf.len(), self.ntt_size())")
        batch = b.words // self._n
        if b.is_dev:
            check(self._fn(name)(self._h, b.ptr, batch, _stream_of(b.t)), name)
        else:
            check(self._fn(name + "_host")(self._h, b.ptr, b.words, batch), name)
        return buf

    def fwd(self, buf):
        """Plan::fwd: in place, natural order in, bit-reversed order out."""
        return self._ntt("fwd", buf)

    def inv(self, buf):
        """Plan::inv: in place, bit-reversed in, natural out, not normalised (inv(fwd(x)) = n x)."""
        return self._ntt("inv", buf)

    def fwd_inv(self, buf):
        """fwd then inv with one upload/download (host slices) -- the round trip of BASELINE config 1/2."""
        b = _Buf(self._device, buf, self._bits // 8, "buf")
        if len(b.shape) == 0 or b.shape[-1] != self._n:
            raise ReferencePanic("assert_eq!(buf.len(), self.ntt_size())")
        batch = b.words // self._n
        if b.is_dev:
            check(self._fn("fwd")(self._h, b.ptr, batch, _stream_of(b.t)), "fwd")
            check(self._fn("inv")(self._h, b.ptr, batch, _stream_of(b.t)), "inv")
        else:
            check(self._fn("fwd_inv_host")(self._h, b.ptr, b.words, batch), "fwd_inv")
        return buf

    def _pw(self, name, *arrs):
        bufs = [_Buf(self._device, a, self._bits // 8, "arg%d" % i) for i, a in enumerate(arrs)]
        # izip! truncates to the shortest slice (prime32.rs:397); whole vectors only (prime32.rs:348)
        nwords = min(b.words for b in bufs)
        dev = bufs[0].is_dev
        if any(b.is_dev != dev for b in bufs):
            raise TypeError("all arguments must be on the same side (all device or all host)")
        ptrs = [b.ptr for b in bufs]
        if dev:
            check(self._fn(name)(self._h, *ptrs, nwords, _stream_of(bufs[0].t)), name)
        else:
            check(self._fn(name + "_host")(self._h, *ptrs, nwords), name)
        return arrs[0]

    def mul_assign_normalize(self, lhs, rhs):
        """lhs[i] = lhs[i] * rhs[i] * n^-1 mod p."""
        return self._pw("mul_assign_normalize", lhs, rhs)

    def normalize(self, values):
        """values[i] = values[i] * n^-1 mod p."""
        return self._pw("normalize", values)

    def mul_accumulate(self, acc, lhs, rhs):
        """acc[i] = acc[i] + lhs[i] * rhs[i] mod p."""
        return self._pw("mul_accumulate", acc, lhs, rhs)


class Plan32Prime(_PrimePlan):
    """prime32::Plan"""
    _bits = 32


class Plan64Prime(_PrimePlan):
    """prime64::Plan"""
    _bits = 64


class _NativePlan:
    """native{32,64,128}::Plan32 / native_binary{32,64,128}::Plan32.

    Words: 32/64-bit -> arrays of 4/8-byte items, last dim n.  128-bit -> 8-byte items with trailing
    dims (n, 2) = little-endian {lo, hi}.  Residue planes: uint32, shape (num_primes, batch..., n).
    """
    _bits = None
    _binary = False

    def __init__(self, handle, n, device):
        self._h, self._n, self._device = handle, n, device
        self._np = _lib.lib().cntt_native_num_primes(handle)

    @classmethod
    def _new(cls, ctor, n, device):
        h = C.c_void_p()
        st = ctor(n, cls._bits, int(cls._binary), device, C.byref(h))
        if st in (_lib.INVALID_SIZE, _lib.INVALID_MODULUS, _lib.NO_ROOT, _lib.UNSUPPORTED):
            return None
        check(st, "try_new")
        return cls(h, n, device)

    @classmethod
    def try_new(cls, n, device=0):
        """Plan32::try_new(n): None unless 32 <= n <= 32768 is a power of two (src/native64.rs:933-942)."""
        return cls._new(_lib.lib().cntt_native_plan_new, n, device)

    @classmethod
    def try_new_extended(cls, n, device=0):
        """EXTENSION, no reference counterpart: the same plan on the primes k*2^17+1 below 2^30, which also
        admit n = 65536 (BASELINE.json configs[4]).  None for native128 (needs ten primes, nine exist)."""
        return cls._new(_lib.lib().cntt_native_plan_new_ext, n, device)

    def ntt_i(self, i):
        """Plan32::ntt_0() .. ntt_9() (src/native64.rs:950-969): the prime32::Plan of the i-th prime.  A plan is a
        pure function of (n, p), so this is an equal plan on the same device -- use it for NTT-domain work on the
        residue planes (mul_accumulate against a pre-transformed key, src/prime32.rs:905-927)."""
        if not 0 <= i < self._np:
            raise IndexError(i)
        cache = self.__dict__.setdefault("_ntt", {})
        if i not in cache:
            cache[i] = Plan32Prime.try_new(self._n, self.ntt_modulus(i), device=self._device)
        return cache[i]

    def __del__(self):
        try:
            if self._h:
                _lib.lib().cntt_native_plan_free(self._h)
                self._h = None
        except Exception:
            pass

    def ntt_size(self):
        return _lib.lib().cntt_native_ntt_size(self._h)

    def num_primes(self):
        return self._np

    def ntt_modulus(self, i):
        """modulus of Plan32::ntt_i()"""
        return _lib.lib().cntt_native_prime(self._h, i)

    def _word_buf(self, x, name):
        b = _Buf(self._device, x, 8 if self._bits == 128 else self._bits // 8, name)
        tail = (self._n, 2) if self._bits == 128 else (self._n,)
        if b.shape[len(b.shape) - len(tail):] != tail:
            raise ReferencePanic("assert_eq!(n, %s.len())" % name)
        b.batch = b.words // int(np.prod(tail))
        return b

    def _planes(self, x, batch):
        b = _Buf(self._device, x, 4, "mod_p")
        if b.words != self._np * batch * self._n:
            raise ReferencePanic("residue planes must hold num_primes * batch * n words")
        return b

    def _fwd(self, name, value, mod_p):
        v = self._word_buf(value, "value")
        m = self._planes(mod_p, v.batch)
        l = _lib.lib()
        if v.is_dev != m.is_dev:
            raise TypeError("value and mod_p must be on the same side (both device tensors or both host arrays)")
        if v.is_dev:   # device-resident, asynchronous on the tensor's stream
            check(getattr(l, "cntt_native_" + name)(self._h, v.ptr, m.ptr, v.batch, _stream_of(v.t)))
        else:          # the reference's host-slice call shape, staged through the plan's arena
            check(getattr(l, "cntt_native_%s_host" % name)(self._h, v.ptr, m.ptr, v.batch * self._n, v.batch))
        return mod_p

    def fwd(self, value, mod_p):
        """Plan32::fwd(value, mod_p0, ...): residues of `value` mod each prime, forward-transformed."""
        return self._fwd("fwd", value, mod_p)

    def fwd_binary(self, value, mod_p):
        """Plan32::fwd_binary (binary plans): `value as u32` without reduction, forward-transformed."""
        if not self._binary:
            raise AttributeError("fwd_binary exists only on native_binary* plans")
        return self._fwd("fwd_binary", value, mod_p)

    def inv(self, value, mod_p):
        """Plan32::inv(value, mod_p0, ...): inverse transforms (clobbering mod_p) + Garner lift."""
        return self._fwd("inv", value, mod_p)

    def negacyclic_polymul(self, prod, lhs, rhs):
        """Plan32::negacyclic_polymul(prod, lhs, rhs)."""
        p = self._word_buf(prod, "prod")
        l = self._word_buf(lhs, "lhs")
        r = self._word_buf(rhs, "rhs")
        if not (p.batch == l.batch == r.batch):
            raise ReferencePanic("assert_eq!(n, lhs.len())")
        if p.is_dev != l.is_dev or p.is_dev != r.is_dev:
            raise TypeError("all arguments must be on the same side (all device or all host)")
        if p.is_dev:
            check(_lib.lib().cntt_native_polymul(self._h, p.ptr, l.ptr, r.ptr, p.batch, _stream_of(p.t)))
        else:
            check(_lib.lib().cntt_native_polymul_host(self._h, p.ptr, l.ptr, r.ptr, p.batch * self._n, p.batch))
        return prod


def _polymul_ntt_rhs(self, prod, lhs, rhs_planes):
    """EXTENSION: negacyclic_polymul with rhs already transformed -- `rhs_planes` is what fwd (fwd_binary on binary plans) wrote, shape
    (num_primes, batch, n) or (num_primes, 1, n) / (num_primes, n) for one key shared by the whole batch.  Device tensors only,
    256 <= n <= 4096."""
    p = self._word_buf(prod, "prod")
    l = self._word_buf(lhs, "lhs")
    r = _Buf(self._device, rhs_planes, 4, "rhs_planes")
    if not (p.is_dev and l.is_dev and r.is_dev):
        raise TypeError("negacyclic_polymul_ntt_rhs operates on device-resident tensors")
    if p.batch != l.batch or r.words not in (self._np * p.batch * self._n, self._np * self._n):
        raise ReferencePanic("assert_eq!(n, lhs.len())")
    rb = r.words // (self._np * self._n)
    check(_lib.lib().cntt_native_polymul_ntt_rhs(self._h, p.ptr, l.ptr, r.ptr, rb, p.batch, _stream_of(p.t)), "negacyclic_polymul_ntt_rhs")
    return prod


_NativePlan.negacyclic_polymul_ntt_rhs = _polymul_ntt_rhs


class _Native52Plan:
    """native32 / native64 / native_binary32 / native_binary64 ::Plan52 (src/native64.rs:29-34,1072-1165 and twins): the
    same plans on 2 / 3 / 1 / 2 ~50-bit primes (primes52) with uint64 residue planes, shape (num_primes, batch..., n).
    In the reference the type needs feature = "nightly" and try_new is None without AVX-512 IFMA; here it is always
    available.  Every method takes device tensors or host arrays; negacyclic_polymul returns exactly what Plan32 returns."""
    _bits = None
    _binary = False

    def __init__(self, handle, n, device):
        self._h, self._n, self._device = handle, n, device
        self._np = _lib.lib().cntt_native52_num_primes(handle)

    @classmethod
    def try_new(cls, n, device=0):
        h = C.c_void_p()
        st = _lib.lib().cntt_native52_plan_new(n, cls._bits, int(cls._binary), device, C.byref(h))
        if st in (_lib.INVALID_SIZE, _lib.INVALID_MODULUS, _lib.NO_ROOT):
            return None
        check(st, "try_new")
        return cls(h, n, device)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().cntt_native52_plan_free(self._h)
                self._h = None
        except Exception:
            pass

    def ntt_size(self):
        return _lib.lib().cntt_native52_ntt_size(self._h)

    def num_primes(self):
        return self._np

    def ntt_modulus(self, i):
        return _lib.lib().cntt_native52_prime(self._h, i)

    def ntt_i(self, i):
        """Plan52::ntt_0() .. : the prime64::Plan of the i-th prime"""
        if not 0 <= i < self._np:
            raise IndexError(i)
        cache = self.__dict__.setdefault("_ntt", {})
        if i not in cache:
            cache[i] = Plan64Prime.try_new(self._n, self.ntt_modulus(i), device=self._device)
        return cache[i]

    def _args(self, value, mod_p):
        v = _Buf(self._device, value, self._bits // 8, "value")
        if len(v.shape) == 0 or v.shape[-1] != self._n:
            raise ReferencePanic("assert_eq!(n, value.len())")
        batch = v.words // self._n
        m = _Buf(self._device, mod_p, 8, "mod_p")
        if m.words != self._np * batch * self._n:
            raise ReferencePanic("residue planes must hold num_primes * batch * n words")
        if v.is_dev != m.is_dev:
            raise TypeError("value and mod_p must be on the same side (both device tensors or both host arrays)")
        return v, m, batch

    def _run(self, name, value, mod_p):
        v, m, batch = self._args(value, mod_p)
        l = _lib.lib()
        if v.is_dev:
            check(getattr(l, "cntt_native52_" + name)(self._h, v.ptr, m.ptr, batch, _stream_of(v.t)))
        else:   # the reference's host-slice call shape (src/native64.rs:1108-1141)
            check(getattr(l, "cntt_native52_%s_host" % name)(self._h, v.ptr, m.ptr, batch * self._n, batch))

    def fwd(self, value, mod_p):
        self._run("fwd", value, mod_p)
        return mod_p

    def fwd_binary(self, value, mod_p):
        if not self._binary:
            raise AttributeError("fwd_binary exists only on native_binary* plans")
        self._run("fwd_binary", value, mod_p)
        return mod_p

    def inv(self, value, mod_p):
        self._run("inv", value, mod_p)
        return value

    def negacyclic_polymul(self, prod, lhs, rhs):
        bufs = [_Buf(self._device, x, self._bits // 8, nm) for x, nm in ((prod, "prod"), (lhs, "lhs"), (rhs, "rhs"))]
        for b in bufs:
            if len(b.shape) == 0 or b.shape[-1] != self._n or b.words != bufs[0].words:
                raise ReferencePanic("assert_eq!(n, lhs.len())")
        if len({b.is_dev for b in bufs}) != 1:
            raise TypeError("all arguments must be on the same side (all device or all host)")
        batch = bufs[0].words // self._n
        l = _lib.lib()
        if bufs[0].is_dev:
            check(l.cntt_native52_polymul(self._h, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, batch, _stream_of(bufs[0].t)))
        else:
            check(l.cntt_native52_polymul_host(self._h, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[0].words, batch))
        return prod


class FwdMode:
    """product::FwdMode (src/product.rs:10-14): FwdMode.Generic or FwdMode.Bounded(bound)."""

    def __init__(self, kind, bound=0):
        self.kind, self.bound = kind, bound

    @staticmethod
    def Bounded(bound):
        return FwdMode(1, int(bound))

    def __repr__(self):
        return "FwdMode.Generic" if self.kind == 0 else "FwdMode.Bounded(%d)" % self.bound


FwdMode.Generic = FwdMode(0)


class InvMode:
    """product::InvMode (src/product.rs:16-20)"""
    Replace = 0
    Accumulate = 1


class ProductPlan:
    """product::Plan (src/product.rs:139-967): negacyclic NTT plan for a modulus that is a product of distinct
    primes.  `standard` buffers are (batch..., n) and NTT-domain buffers (batch..., ntt_domain_len) u64 words in
    the reference's packed layout (u32 planes first, then u64 planes); torch CUDA int64 tensors run device-resident
    and asynchronously, numpy uint64 arrays take the host-slice path (the reference's call shape)."""

    def __init__(self, handle, n, modulus, device):
        self._h, self._n, self._modulus, self._device = handle, n, modulus, device
        self._dl = _lib.lib().cntt_product_ntt_domain_len(handle)

    @classmethod
    def try_new(cls, polynomial_size, modulus, factors, device=0):
        l = _lib.lib()
        h = C.c_void_p()
        fs = [int(f) for f in factors]
        arr = (C.c_uint64 * max(1, len(fs)))(*fs)
        st = l.cntt_product_plan_new(polynomial_size, modulus, arr, len(fs), device, C.byref(h))
        if st in (_lib.INVALID_SIZE, _lib.INVALID_MODULUS, _lib.NO_ROOT):
            return None
        check(st, "try_new")
        return cls(h, polynomial_size, modulus, device)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().cntt_product_plan_free(self._h)
                self._h = None
        except Exception:
            pass

    def ntt_size(self):
        return _lib.lib().cntt_product_ntt_size(self._h)

    def modulus(self):
        return _lib.lib().cntt_product_modulus(self._h)

    def ntt_domain_len(self):
        return self._dl

    def primes(self):
        c32, c64 = C.c_int(), C.c_int()
        check(_lib.lib().cntt_product_num_primes(self._h, C.byref(c32), C.byref(c64)))
        return [_lib.lib().cntt_product_prime(self._h, i) for i in range(c32.value + c64.value)]

    def _std(self, x, name):
        b = _Buf(self._device, x, 8, name)
        if len(b.shape) == 0 or b.shape[-1] != self._n:
            raise ReferencePanic("assert_eq!(standard.len(), self.ntt_size())")
        b.batch = b.words // self._n
        return b

    def _dom(self, x, name, like=None):
        b = _Buf(self._device, x, 8, name)
        if self._dl == 0:
            b.batch = like.batch if like is not None else 0
            if b.words != 0:
                raise ReferencePanic("assert_eq!(%s.len(), self.ntt_domain_len())" % name)
        else:
            if len(b.shape) == 0 or b.shape[-1] != self._dl:
                raise ReferencePanic("assert_eq!(%s.len(), self.ntt_domain_len())" % name)
            b.batch = b.words // self._dl
        if like is not None:
            if b.batch != like.batch:
                raise ReferencePanic("batch sizes differ")
            if b.is_dev != like.is_dev:
                raise TypeError("all buffers of one call must be device tensors or all host arrays")
        return b

    def fwd(self, ntt, standard, mode=FwdMode.Generic):
        s = self._std(standard, "standard")
        d = self._dom(ntt, "ntt", s)
        l = _lib.lib()
        if s.is_dev:
            check(l.cntt_product_fwd(self._h, d.ptr, s.ptr, mode.kind, mode.bound, s.batch, _stream_of(s.t)), "fwd")
        else:
            check(l.cntt_product_fwd_host(self._h, d.ptr, s.ptr, d.words, s.words, mode.kind, mode.bound, s.batch), "fwd")
        return ntt

    def inv(self, standard, ntt, mode=InvMode.Replace):
        s = self._std(standard, "standard")
        d = self._dom(ntt, "ntt", s)
        l = _lib.lib()
        if s.is_dev:
            check(l.cntt_product_inv(self._h, s.ptr, d.ptr, mode, s.batch, _stream_of(s.t)), "inv")
        else:
            check(l.cntt_product_inv_host(self._h, s.ptr, d.ptr, s.words, d.words, mode, s.batch), "inv")
        return standard

    def mul_assign_normalize(self, lhs, rhs):
        a = self._dom(lhs, "lhs")
        b = self._dom(rhs, "rhs", a)
        l = _lib.lib()
        if a.is_dev:
            check(l.cntt_product_mul_assign_normalize(self._h, a.ptr, b.ptr, a.batch, _stream_of(a.t)))
        else:
            check(l.cntt_product_mul_assign_normalize_host(self._h, a.ptr, b.ptr, a.words, a.batch))
        return lhs

    def normalize(self, values):
        a = self._dom(values, "values")
        l = _lib.lib()
        if a.is_dev:
            check(l.cntt_product_normalize(self._h, a.ptr, a.batch, _stream_of(a.t)))
        else:
            check(l.cntt_product_normalize_host(self._h, a.ptr, a.words, a.batch))
        return values

    def mul_accumulate(self, acc, lhs, rhs):
        a = self._dom(acc, "acc")
        x = self._dom(lhs, "lhs", a)
        y = self._dom(rhs, "rhs", a)
        l = _lib.lib()
        if a.is_dev:
            check(l.cntt_product_mul_accumulate(self._h, a.ptr, x.ptr, y.ptr, a.batch, _stream_of(a.t)))
        else:
            check(l.cntt_product_mul_accumulate_host(self._h, a.ptr, x.ptr, y.ptr, a.words, a.batch))
        return acc


class HostMulti:
    """EXTENSION: one host batch over several GPUs in ONE process (cntt_*_host_multi).  `plans` are equal plans built on
    different devices (prime32 / prime64 plans: fwd, inv, fwd_inv; native Plan32: negacyclic_polymul); the batch is cut into
    len(plans) contiguous shards, each staged and transformed on its own device concurrently, no collective."""

    def __init__(self, plans):
        if not plans:
            raise ValueError("at least one plan")
        self._plans = list(plans)
        self._arr = (C.c_void_p * len(self._plans))(*[p._h for p in self._plans])
        self._p0 = self._plans[0]

    def _prime(self, name, buf):
        p0 = self._p0
        b = _Buf(None, buf, p0._bits // 8, "buf")
        if b.is_dev:
            raise TypeError("HostMulti takes host (numpy) batches; device-resident batches are sharded by the caller")
        if len(b.shape) == 0 or b.shape[-1] != p0._n:
            raise ReferencePanic("assert_eq!(buf.len(), self.ntt_size())")
        fn = getattr(_lib.lib(), "cntt_prime%d_%s_host_multi" % (p0._bits, name))
        check(fn(self._arr, len(self._plans), b.ptr, b.words, b.words // p0._n), name)
        return buf

    def fwd(self, buf):
        return self._prime("fwd", buf)

    def inv(self, buf):
        return self._prime("inv", buf)

    def fwd_inv(self, buf):
        return self._prime("fwd_inv", buf)

    def negacyclic_polymul(self, prod, lhs, rhs):
        p0 = self._p0
        bufs = [p0._word_buf(x, nm) for x, nm in ((prod, "prod"), (lhs, "lhs"), (rhs, "rhs"))]
        if any(b.is_dev for b in bufs) or not (bufs[0].batch == bufs[1].batch == bufs[2].batch):
            raise TypeError("HostMulti.negacyclic_polymul takes three host arrays of one shape")
        check(_lib.lib().cntt_native_polymul_host_multi(self._arr, len(self._plans), bufs[0].ptr, bufs[1].ptr, bufs[2].ptr,
                                                        bufs[0].batch * p0._n, bufs[0].batch))
        return prod
