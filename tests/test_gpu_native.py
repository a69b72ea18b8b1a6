"""GPU parity of the native / native_binary plans against the CPU oracle and the wrapping schoolbook."""
import numpy as np
import pytest

from conftest import rng, rand_words

pytestmark = pytest.mark.gpu


def dev(torch, a):
    sd = np.int32 if a.dtype.itemsize == 4 else np.int64
    return torch.from_numpy(a.view(sd).copy()).cuda()


def host(t, dtype):
    return t.cpu().numpy().view(dtype)


def plan_pair(cntt, oracle, n, bits, binary):
    mod = getattr(cntt, ("native_binary%d" if binary else "native%d") % bits)
    return mod.Plan32.try_new(n), oracle.Native.try_new(n, bits, binary=binary)


def make_rhs(g, bits, shape, binary):
    r = rand_words(g, bits, shape)
    if binary:
        r = r & r.dtype.type(1)
        if bits == 128:
            r[..., 1] = 0
    return r


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("bits", [32, 64, 128])
@pytest.mark.parametrize("n", [32, 64, 256, 1024, 2048, 4096])
def test_polymul_matches_oracle(cntt, oracle, torch_cuda, n, bits, binary):
    g = rng(n * 7 + bits + int(binary))
    gp, op = plan_pair(cntt, oracle, n, bits, binary)
    batch = 5
    wdt = np.uint32 if bits == 32 else np.uint64
    lhs = rand_words(g, bits, (batch, n))
    rhs = make_rhs(g, bits, (batch, n), binary)
    ref = op.negacyclic_polymul(lhs, rhs)
    dl, dr = dev(torch_cuda, lhs), dev(torch_cuda, rhs)
    dp = torch_cuda.empty_like(dl)
    gp.negacyclic_polymul(dp, dl, dr)
    assert (host(dp, wdt) == ref).all()
    # host-slice flavour
    hp = np.empty_like(lhs)
    gp.negacyclic_polymul(hp, lhs, rhs)
    assert (hp == ref).all()
    if n <= 256:  # reference's own oracle: wrapping schoolbook (src/native64.rs:1196-1215)
        sb = [oracle.schoolbook32(0, l, r) if bits == 32 else oracle.schoolbook64(0, l, r) if bits == 64
              else oracle.schoolbook128(l, r) for l, r in zip(lhs, rhs)]
        assert (np.stack(sb) == ref).all()


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("bits", [32, 64, 128])
def test_split_fwd_inv(cntt, oracle, torch_cuda, bits, binary):
    """Plan32::fwd / fwd_binary / inv on device residue planes (src/native64.rs:971-1038)."""
    n, batch = 512, 3
    g = rng(bits + 100 * int(binary))
    gp, op = plan_pair(cntt, oracle, n, bits, binary)
    wdt = np.uint32 if bits == 32 else np.uint64
    val = rand_words(g, bits, (batch, n))
    npz = gp.num_primes()
    assert npz == op.nprimes and [gp.ntt_modulus(i) for i in range(npz)] == [oracle.primes32(i) for i in range(npz)]
    planes = torch_cuda.empty((npz, batch, n), dtype=torch_cuda.int32, device="cuda")
    gp.fwd(dev(torch_cuda, val), planes)
    ref_planes = np.stack([op.fwd(v) for v in val], axis=1)          # (np, batch, n)
    assert (host(planes, np.uint32) == ref_planes).all()
    if binary:
        bval = make_rhs(g, bits, (batch, n), True)
        gp.fwd_binary(dev(torch_cuda, bval), planes)
        assert (host(planes, np.uint32) == np.stack([op.fwd_binary(v) for v in bval], axis=1)).all()
        gp.fwd(dev(torch_cuda, val), planes)
    out = dev(torch_cuda, np.zeros_like(val))
    gp.inv(out, planes)
    ref_out = np.stack([op.inv(np.ascontiguousarray(ref_planes[:, b])) for b in range(batch)])
    assert (host(out, wdt) == ref_out).all()


@pytest.mark.parametrize("bits", [32, 64, 128])
def test_crt_on_arbitrary_residues(cntt, oracle, torch_cuda, bits):
    """The Garner lift must match the reference formula on *arbitrary* residues, not only on polymul
    outputs (sign rule v_last > P_last/2; src/native64.rs:1245-1293 pins this for SIMD vs scalar).
    Feed inv() planes that are forward transforms of arbitrary residue vectors."""
    n, batch = 64, 4
    g = rng(bits + 5)
    for binary in (False, True):
        gp, op = plan_pair(cntt, oracle, n, bits, binary)
        npz = op.nprimes
        res = np.stack([g.integers(0, oracle.primes32(k), size=(batch, n), dtype=np.uint64).astype(np.uint32) for k in range(npz)])
        # transform each residue vector forward with the per-prime oracle plan so that inv() returns them
        fw = res.copy()
        for k in range(npz):
            pk = oracle.Plan32.try_new(n, oracle.primes32(k))
            for b in range(batch):
                pk.fwd(fw[k, b])
        wdt = np.uint32 if bits == 32 else np.uint64
        shape = (batch, n) if bits != 128 else (batch, n, 2)
        out = dev(torch_cuda, np.zeros(shape, wdt))
        gp.inv(out, dev(torch_cuda, fw))
        ref = np.stack([op.inv(np.ascontiguousarray(fw[:, b])) for b in range(batch)])
        assert (host(out, wdt) == ref).all(), (bits, binary)


@pytest.mark.parametrize("n", [8192, 16384, 32768])
def test_native_large_n(cntt, oracle, torch_cuda, n):
    """4096 < N <= 32768: the three-kernel path (native_large.cuh), every kind, ragged batch of 3.
    Reference maximum is 32768 (P1 - 1 = 2^16 * odd); 65536 -> None like the reference."""
    g = rng(n)
    kinds = [(64, False), (64, True), (32, False), (32, True)] + ([(128, False), (128, True)] if n <= 16384 else [(128, True)])
    for bits, binary in kinds:
        gp, op = plan_pair(cntt, oracle, n, bits, binary)
        lhs = rand_words(g, bits, (3, n))
        rhs = make_rhs(g, bits, (3, n), binary)
        dl, dr = dev(torch_cuda, lhs), dev(torch_cuda, rhs)
        dp = torch_cuda.empty_like(dl)
        gp.negacyclic_polymul(dp, dl, dr)
        assert (host(dp, np.uint32 if bits == 32 else np.uint64) == op.negacyclic_polymul(lhs, rhs)).all()
    assert cntt.native64.Plan32.try_new(65536) is None
    assert cntt.native_binary64.Plan32.try_new(65536) is None


def test_polymul_errors(cntt, torch_cuda):
    plan = cntt.native64.Plan32.try_new(64)
    z = lambda k: torch_cuda.zeros(k, dtype=torch_cuda.int64, device="cuda")
    with pytest.raises(cntt.ReferencePanic):
        plan.negacyclic_polymul(z(64), z(63), z(64))          # src/native64.rs:1043-1045
    with pytest.raises(AttributeError):
        plan.fwd_binary(z(64), torch_cuda.zeros((5, 64), dtype=torch_cuda.int32, device="cuda"))
    assert cntt.native64.Plan32.try_new(48) is None


def test_full_size_config3_properties(cntt, oracle, torch_cuda):
    """BASELINE config 3 at full size (native64 N=2048, batch 2^16): oracle on sampled polynomials,
    plus x * 1 == x and bilinearity (a * (b + c) == a * b + a * c, wrapping) on the whole batch."""
    torch = torch_cuda
    n, batch = 2048, 1 << 16
    g = rng(3)
    plan = cntt.native64.Plan32.try_new(n)
    op = oracle.Native.try_new(n, 64)
    a = rand_words(g, 64, (batch, n))
    b = rand_words(g, 64, (batch, n))
    c = rand_words(g, 64, (batch, n))
    da, db, dc = dev(torch, a), dev(torch, b), dev(torch, c)
    ab, ac, abc = torch.empty_like(da), torch.empty_like(da), torch.empty_like(da)
    plan.negacyclic_polymul(ab, da, db)
    plan.negacyclic_polymul(ac, da, dc)
    plan.negacyclic_polymul(abc, da, db + dc)       # int64 add wraps mod 2^64
    assert torch.equal(abc, ab + ac)
    hab = host(ab, np.uint64)
    for r in g.integers(0, batch, 6):
        assert (hab[r] == op.negacyclic_polymul(a[r], b[r])).all()
    one = torch.zeros_like(da)
    one[:, 0] = 1
    plan.negacyclic_polymul(ab, da, one)
    assert torch.equal(ab, da)


# ---- extended plans (EXTENSION, no reference counterpart): primes k*2^17+1, n up to 65536 -----------------------------
EXT_PRIMES = [0x3F3A0001, 0x3F540001, 0x3F5A0001, 0x3F760001, 0x3F820001, 0x3FAC0001, 0x3FD20001, 0x3FDE0001, 0x3FFC0001]


def ext_plan(cntt, n, bits, binary):
    mod = getattr(cntt, ("native_binary%d" if binary else "native%d") % bits)
    return mod.Plan32.try_new_extended(n)


@pytest.mark.parametrize("n", [64, 2048, 4096, 16384])
def test_extended_polymul_equals_reference_plan(cntt, oracle, torch_cuda, n):
    """Wherever both constructors succeed the polymul of an extended plan is bit-identical to the reference
    plan's (the product is exact and then wrapped: it cannot depend on the primes that carried it)."""
    g = rng(900 + n)
    for bits, binary in [(32, False), (64, False), (32, True), (64, True), (128, True)]:
        gp = ext_plan(cntt, n, bits, binary)
        op = oracle.Native.try_new(n, bits, binary=binary)
        assert [gp.ntt_modulus(i) for i in range(gp.num_primes())] == EXT_PRIMES[:gp.num_primes()]
        lhs = rand_words(g, bits, (3, n))
        rhs = make_rhs(g, bits, (3, n), binary)
        dl, dr = dev(torch_cuda, lhs), dev(torch_cuda, rhs)
        dp = torch_cuda.empty_like(dl)
        gp.negacyclic_polymul(dp, dl, dr)
        assert (host(dp, np.uint32 if bits == 32 else np.uint64) == op.negacyclic_polymul(lhs, rhs)).all(), (bits, binary)
    assert cntt.native128.Plan32.try_new_extended(n) is None   # ten primes needed, nine exist


def test_extended_n65536_polymul(cntt, oracle, torch_cuda):
    """BASELINE configs[4]: native_binary64 N = 65536 (and the other kinds the nine primes can carry) against the
    wrapping negacyclic schoolbook -- the specification itself, the reference has no plan of this size."""
    n = 65536
    g = rng(65536)
    assert cntt.native_binary64.Plan32.try_new(n) is None            # reference behaviour is unchanged
    for bits, binary, batch in [(64, True, 3), (64, False, 2), (32, False, 2), (32, True, 2), (128, True, 1)]:
        gp = ext_plan(cntt, n, bits, binary)
        assert gp is not None and gp.ntt_size() == n
        lhs = rand_words(g, bits, (batch, n))
        if bits == 128:  # keep the O(n^2) u128 oracle short: sparse lhs (zero coefficients are skipped)
            keep = g.integers(0, n, 512)
            sparse = np.zeros_like(lhs)
            sparse[:, keep] = lhs[:, keep]
            lhs = sparse
        rhs = make_rhs(g, bits, (batch, n), binary)
        dl, dr = dev(torch_cuda, lhs), dev(torch_cuda, rhs)
        dp = torch_cuda.empty_like(dl)
        gp.negacyclic_polymul(dp, dl, dr)
        got = host(dp, np.uint32 if bits == 32 else np.uint64)
        for b in range(batch):
            assert (got[b] == oracle.negacyclic_wrapping(bits, lhs[b], rhs[b])).all(), (bits, binary, b)
        hp = np.empty_like(lhs)                                        # host-slice flavour
        gp.negacyclic_polymul(hp, lhs, rhs)
        assert (hp == got).all()


def test_extended_n65536_split_planes(cntt, oracle, torch_cuda):
    """fwd / inv of an extended plan at N = 65536: planes are the per-prime transforms (oracle prime32 plans on
    the extended primes) of value mod p; inv(fwd(x)) == N * x wrapping (centred lift is exact: N x < prod/2)."""
    n, batch = 65536, 2
    g = rng(17)
    gp = ext_plan(cntt, n, 64, True)
    val = rand_words(g, 64, (batch, n))
    planes = torch_cuda.empty((3, batch, n), dtype=torch_cuda.int32, device="cuda")
    gp.fwd(dev(torch_cuda, val), planes)
    hp = host(planes, np.uint32)
    for k in range(3):
        p = EXT_PRIMES[k]
        op = oracle.Plan32.try_new(n, p)
        for b in range(batch):
            assert (hp[k, b] == op.fwd((val[b] % np.uint64(p)).astype(np.uint32))).all()
    out = dev(torch_cuda, np.zeros_like(val))
    gp.inv(out, planes)
    assert (host(out, np.uint64) == val * np.uint64(n)).all()


def test_split_phase_key_switch_shape(cntt, oracle, torch_cuda):
    """SURVEY.md 8(f) rank 2 -- the TFHE usage: the key stays in the NTT domain, many mul_accumulate per inv.
    sum_k a_k * b_k (negacyclic, wrapping) via fwd planes, ntt_i().mul_accumulate on every plane, one inv.
    inv returns N * (sum), so the check is against N * sum of oracle polymuls (src/native64.rs:971-1038,
    src/prime32.rs:905-927)."""
    torch = torch_cuda
    n, batch, terms = 1024, 4, 3
    g = rng(2718)
    gp = cntt.native64.Plan32.try_new(n)
    op = oracle.Native.try_new(n, 64)
    npz = gp.num_primes()
    a = rand_words(g, 64, (terms, batch, n))
    b = rand_words(g, 64, (terms, batch, n))
    # keep the integer sum inside the centred CRT range: |sum| * N < prod(P)/2 needs small-ish operands
    a >>= np.uint64(8)
    b >>= np.uint64(8)
    acc = torch.zeros((npz, batch, n), dtype=torch.int32, device="cuda")
    pa = torch.empty_like(acc)
    pb = torch.empty_like(acc)
    for k in range(terms):
        gp.fwd(dev(torch, a[k]), pa)
        gp.fwd(dev(torch, b[k]), pb)
        for i in range(npz):
            gp.ntt_i(i).mul_accumulate(acc[i], pa[i], pb[i])
    out = torch.empty((batch, n), dtype=torch.int64, device="cuda")
    gp.inv(out, acc)
    ref = np.zeros((batch, n), np.uint64)
    for k in range(terms):
        ref += op.negacyclic_polymul(a[k], b[k])
    assert (host(out, np.uint64) == ref * np.uint64(n)).all()



@pytest.mark.parametrize("n", [2048, 4096, 32768])
def test_polymul_extreme_magnitudes(cntt, oracle, torch_cuda, n):
    """Largest |coefficient| the plans can meet (all words 2^w - 1: coefficient n-1 = +n max^2, coefficient 0 =
    -(n-2) max^2): the reconstruction's quotient estimate (native_device.cuh, reconstruct_bounded) must still land
    on the centred lift.  Checked against the oracle's exact Garner path."""
    for bits, binary in [(32, False), (64, False), (128, False), (32, True), (64, True), (128, True)]:
        if n > 16384 and bits == 128 and not binary:
            continue
        gp, op = plan_pair(cntt, oracle, n, bits, binary)
        wdt = np.uint32 if bits == 32 else np.uint64
        shape = (2, n) if bits != 128 else (2, n, 2)
        lhs = np.full(shape, np.iinfo(wdt).max, wdt)
        rhs = np.full(shape, np.iinfo(wdt).max, wdt)
        lhs[1, ::2] = 0                 # second polynomial: mixed signs / half the terms
        if binary:
            rhs = np.ones(shape, wdt)
            if bits == 128:
                rhs[..., 1] = 0
        dl, dr = dev(torch_cuda, lhs), dev(torch_cuda, rhs)
        dp = torch_cuda.empty_like(dl)
        gp.negacyclic_polymul(dp, dl, dr)
        assert (host(dp, wdt) == op.negacyclic_polymul(lhs, rhs)).all(), (bits, binary)


def test_extended_n65536_extreme_magnitudes(cntt, oracle, torch_cuda):
    n = 65536
    for bits, binary in [(64, False), (64, True), (32, False)]:
        gp = ext_plan(cntt, n, bits, binary)
        wdt = np.uint32 if bits == 32 else np.uint64
        lhs = np.full((1, n), np.iinfo(wdt).max, wdt)
        rhs = np.ones((1, n), wdt) if binary else np.full((1, n), np.iinfo(wdt).max, wdt)
        dl, dr = dev(torch_cuda, lhs), dev(torch_cuda, rhs)
        dp = torch_cuda.empty_like(dl)
        gp.negacyclic_polymul(dp, dl, dr)
        assert (host(dp, wdt)[0] == oracle.negacyclic_wrapping(bits, lhs[0], rhs[0])).all(), (bits, binary)


@pytest.mark.parametrize("bits,binary", [(32, False), (64, False), (128, False), (64, True)])
def test_split_fwd_inv_host_slices(cntt, oracle, bits, binary):
    """The reference's own call shape for Plan32::fwd / fwd_binary / inv: host slices (cntt_native_*_host), one
    polynomial and a small batch; inv also returns the clobbered residue buffers (src/native64.rs:1001-1014)."""
    n = 256
    g = rng(bits * 3 + int(binary))
    gp, op = plan_pair(cntt, oracle, n, bits, binary)
    npz = gp.num_primes()
    for batch in (1, 3):
        val = rand_words(g, bits, (batch, n))
        planes = np.zeros((npz, batch, n), np.uint32)
        gp.fwd(val, planes)
        ref_planes = np.stack([op.fwd(v) for v in val], axis=1)
        assert (planes == ref_planes).all()
        if binary:
            bval = make_rhs(g, bits, (batch, n), True)
            gp.fwd_binary(bval, planes)
            assert (planes == np.stack([op.fwd_binary(v) for v in bval], axis=1)).all()
            gp.fwd(val, planes)
        out = np.zeros_like(val)
        gp.inv(out, planes)
        ref_out, ref_clobbered = [], []
        for b in range(batch):
            pb = np.ascontiguousarray(ref_planes[:, b])
            ref_out.append(op.inv(pb))
            ref_clobbered.append(pb)
        assert (out == np.stack(ref_out)).all()
        assert (planes == np.stack(ref_clobbered, axis=1)).all()
    with pytest.raises(cntt.ReferencePanic):
        gp.fwd(rand_words(g, bits, (1, n // 2)), np.zeros((npz, 1, n // 2), np.uint32))


@pytest.mark.parametrize("bits,binary", [(32, False), (64, False), (32, True), (64, True)])
@pytest.mark.parametrize("n", [32, 64, 1024, 4096, 16384])
def test_plan52(cntt, oracle, torch_cuda, bits, binary, n):
    """Plan52 twins (primes52, u64 residue planes over prime64 plans): fwd / fwd_binary planes and inv lift against the
    oracle restatement, inv on arbitrary canonical residues (sign rule v_top > P_top / 2), and negacyclic_polymul
    equal to Plan32's (and to the wrapping schoolbook at n = 64)."""
    torch = torch_cuda
    mod = getattr(cntt, ("native_binary%d" if binary else "native%d") % bits)
    gp, op = mod.Plan52.try_new(n), oracle.Native52.try_new(n, bits, binary)
    assert gp is not None and gp.ntt_size() == n
    npz = gp.num_primes()
    assert [gp.ntt_modulus(i) for i in range(npz)] == list(oracle.PRIMES52[:npz]) == op.primes
    g = rng(52 + bits + n + int(binary))
    batch = 3
    wdt = np.uint32 if bits == 32 else np.uint64
    val = rand_words(g, bits, (batch, n))
    planes = torch.empty((npz, batch, n), dtype=torch.int64, device="cuda")
    gp.fwd(dev(torch, val), planes)
    ref = np.stack([op.fwd(v) for v in val], axis=1)
    assert (host(planes, np.uint64) == ref).all()
    if binary:
        bval = make_rhs(g, bits, (batch, n), True)
        gp.fwd_binary(dev(torch, bval), planes)
        assert (host(planes, np.uint64) == np.stack([op.fwd(v, binary_copy=True) for v in bval], axis=1)).all()
    # inv on arbitrary canonical residues: forward-transform random residue vectors so that inv() returns them
    res = np.stack([g.integers(0, p, size=(batch, n), dtype=np.uint64) for p in op.primes])
    res[:, 0, :4] = np.array([[0, 1, p // 2, p - 1] for p in op.primes], dtype=np.uint64)
    fw = np.stack([np.stack([op.plans[k].fwd(res[k, b].copy()) for b in range(batch)]) for k in range(npz)])
    out = dev(torch, np.zeros((batch, n), wdt))
    gp.inv(out, dev(torch, fw))
    want = np.stack([op.inv(np.ascontiguousarray(fw[:, b])) for b in range(batch)])
    assert (host(out, wdt) == want).all()
    # polymul: identical to Plan32 (exact product, wrapped), device and host flavours
    lhs, rhs = rand_words(g, bits, (batch, n)), make_rhs(g, bits, (batch, n), binary)
    p32 = oracle.Native.try_new(n, bits, binary=binary).negacyclic_polymul(lhs, rhs)
    dp = torch.empty_like(dev(torch, lhs))
    gp.negacyclic_polymul(dp, dev(torch, lhs), dev(torch, rhs))
    assert (host(dp, wdt) == p32).all()
    hp = np.empty_like(lhs)
    gp.negacyclic_polymul(hp, lhs, rhs)
    assert (hp == p32).all()
    # and the split path composes to the same product: fwd, mul_assign_normalize per prime, inv
    pl_, pr_ = torch.empty_like(planes), torch.empty_like(planes)
    gp.fwd(dev(torch, lhs), pl_)
    (gp.fwd_binary if binary else gp.fwd)(dev(torch, rhs), pr_)
    for k in range(npz):
        gp.ntt_i(k).mul_assign_normalize(pl_[k], pr_[k])
    out2 = torch.empty_like(dp)
    gp.inv(out2, pl_)
    assert (host(out2, wdt) == p32).all()
    # host-slice flavours (the reference's call shape): same planes, same lift, planes clobbered like the device call
    hplanes = np.zeros((npz, batch, n), np.uint64)
    gp.fwd(val, hplanes)
    assert (hplanes == ref).all()
    if binary:
        gp.fwd_binary(bval, hplanes)
        assert (hplanes == np.stack([op.fwd(v, binary_copy=True) for v in bval], axis=1)).all()
    hfw, hout = fw.copy(), np.zeros((batch, n), wdt)
    gp.inv(hout, hfw)
    assert (hout == want).all()
    dfw = dev(torch, fw)
    gp.inv(torch.empty_like(out), dfw)
    assert (hfw == host(dfw, np.uint64)).all()


@pytest.mark.parametrize("bits,binary", [(32, False), (64, False), (128, False), (32, True), (64, True), (128, True)])
@pytest.mark.parametrize("n", [256, 2048, 4096])
def test_polymul_with_pretransformed_rhs(cntt, oracle, torch_cuda, bits, binary, n):
    """EXTENSION cntt_native_polymul_ntt_rhs: rhs handed over as the residue planes Plan32::fwd / fwd_binary wrote (the TFHE shape, key
    kept in the NTT domain) -- the product must be what negacyclic_polymul returns, per-product keys and one shared key."""
    torch = torch_cuda
    mod = getattr(cntt, ("native_binary%d" if binary else "native%d") % bits)
    gp, op = mod.Plan32.try_new(n), oracle.Native.try_new(n, bits, binary=binary)
    g = rng(900 + bits + n + int(binary))
    batch = 5
    lhs, rhs = rand_words(g, bits, (batch, n)), make_rhs(g, bits, (batch, n), binary)
    want = op.negacyclic_polymul(lhs, rhs)
    planes = torch.empty((gp.num_primes(), batch, n), dtype=torch.int32, device="cuda")
    (gp.fwd_binary if binary else gp.fwd)(dev(torch, rhs), planes)
    dl = dev(torch, lhs)
    prod = torch.empty_like(dl)
    gp.negacyclic_polymul_ntt_rhs(prod, dl, planes)
    wdt = np.uint32 if bits == 32 else np.uint64
    assert (host(prod, wdt) == want).all()
    # one key for the whole batch
    key = torch.empty((gp.num_primes(), 1, n), dtype=torch.int32, device="cuda")
    (gp.fwd_binary if binary else gp.fwd)(dev(torch, rhs[:1]), key)
    gp.negacyclic_polymul_ntt_rhs(prod, dl, key)
    assert (host(prod, wdt) == op.negacyclic_polymul(lhs, np.repeat(rhs[:1], batch, axis=0))).all()
    small = mod.Plan32.try_new(64)
    with pytest.raises(cntt.CnttError):
        small.negacyclic_polymul_ntt_rhs(torch.empty((1, 64) + ((2,) if bits == 128 else ()), dtype=dl.dtype, device="cuda"),
                                         torch.empty((1, 64) + ((2,) if bits == 128 else ()), dtype=dl.dtype, device="cuda"),
                                         torch.empty((small.num_primes(), 1, 64), dtype=torch.int32, device="cuda"))
