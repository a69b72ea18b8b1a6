"""GPU parity of product::Plan (src/product.rs:139-967) against the CPU oracle -- bit-exact, through the C ABI.
The oracle takes one polynomial per call like the reference; the device calls take the slices concatenated."""
import numpy as np
import pytest

from conftest import rng, rand_mod
from test_oracle import _product_cases, _prod

pytestmark = pytest.mark.gpu


def dev(torch, a):
    return torch.from_numpy(a.view(np.int64).copy()).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


CASES = ["u64x1", "u32x1", "u32x2", "u30x2", "u32x4", "u32x2_u64x1"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("n", [256, 2048])
def test_product_fwd_inv_pointwise(cntt, oracle, torch_cuda, case, n):
    torch = torch_cuda
    ps = _product_cases(oracle, n)[case]
    p = _prod(ps)
    op = oracle.Product.try_new(n, p, ps)
    gp = cntt.product.Plan.try_new(n, p, ps)
    assert gp is not None and gp.ntt_size() == n and gp.modulus() == p
    assert gp.ntt_domain_len() == op.ntt_domain_len() and gp.primes() == sorted(ps)
    dl, batch = op.ntt_domain_len(), 5
    g = rng(n + len(case))
    std = rand_mod(g, p, (batch, n), np.uint64)
    std2 = rand_mod(g, p, (batch, n), np.uint64)
    # fwd
    ref_a = np.zeros((batch, dl), dtype=np.uint64)
    ref_b = np.zeros((batch, dl), dtype=np.uint64)
    for b in range(batch):
        op.fwd(ref_a[b], std[b], op.GENERIC)
        op.fwd(ref_b[b], std2[b], op.GENERIC)
    da = torch.zeros((batch, dl), dtype=torch.int64, device="cuda")
    db = torch.zeros_like(da)
    gp.fwd(da, dev(torch, std))
    gp.fwd(db, dev(torch, std2))
    assert (host(da) == ref_a).all(), (case, n, "fwd")
    assert (host(db) == ref_b).all(), (case, n, "fwd")
    # mul_accumulate, normalize, mul_assign_normalize on NTT-domain buffers
    ref_acc = ref_a.copy()
    dacc = da.clone()
    for b in range(batch):
        op.mul_accumulate(ref_acc[b], ref_a[b], ref_b[b])
    gp.mul_accumulate(dacc, da, db)
    assert (host(dacc) == ref_acc).all(), (case, n, "mul_accumulate")
    for b in range(batch):
        op.normalize(ref_acc[b])
    gp.normalize(dacc)
    assert (host(dacc) == ref_acc).all(), (case, n, "normalize")
    for b in range(batch):
        op.mul_assign_normalize(ref_a[b], ref_b[b])
    gp.mul_assign_normalize(da, db)
    assert (host(da) == ref_a).all(), (case, n, "mul_assign_normalize")
    # inv: Replace, then Accumulate on top of a random reduced vector
    out_ref = np.zeros((batch, n), dtype=np.uint64)
    keep = ref_a.copy()
    for b in range(batch):
        op.inv(out_ref[b], ref_a[b], op.REPLACE)
    dkeep = da.clone()
    dout = torch.zeros((batch, n), dtype=torch.int64, device="cuda")
    gp.inv(dout, da, cntt.product.InvMode.Replace)
    assert (host(dout) == out_ref).all(), (case, n, "inv replace")
    assert (host(da) == ref_a).all(), (case, n, "inv clobbers ntt like the reference")
    acc0 = rand_mod(g, p, (batch, n), np.uint64)
    acc_ref = acc0.copy()
    for b in range(batch):
        op.inv(acc_ref[b], keep[b], op.ACCUMULATE)
    dacc2 = dev(torch, acc0)
    gp.inv(dacc2, dkeep, cntt.product.InvMode.Accumulate)
    assert (host(dacc2) == acc_ref).all(), (case, n, "inv accumulate")


@pytest.mark.parametrize("n", [64, 1024])
def test_product_bounded(cntt, oracle, torch_cuda, n):
    """FwdMode::Bounded fast path of the two-u32-prime plan (product.rs:305-322), and the fallthrough to the
    generic reduction when the bound is not below both primes."""
    torch = torch_cuda
    f = oracle.largest_prime_in_arithmetic_progression64
    p0 = f(2 * n, 1, 0, 1 << 31)
    p1 = f(2 * n, 1, 0, p0 - 1)
    p = p0 * p1
    op, gp = oracle.Product.try_new(n, p, [p0, p1]), cntt.product.Plan.try_new(n, p, [p1, p0])
    g = rng(5 + n)
    batch = 4
    for bound in (1 << 20, p1 - 1, p0 + 5):
        lim = min(bound, 1 << 62)
        small = g.integers(-lim + 1, lim, size=(batch, n))
        std = np.array([[int(x) % p for x in row] for row in small], dtype=np.uint64)
        ref = np.zeros((batch, op.ntt_domain_len()), dtype=np.uint64)
        for b in range(batch):
            op.fwd(ref[b], std[b], op.bounded(bound))
        d = torch.zeros((batch, op.ntt_domain_len()), dtype=torch.int64, device="cuda")
        gp.fwd(d, dev(torch, std), cntt.product.FwdMode.Bounded(bound))
        assert (host(d) == ref).all(), bound


def test_product_try_new_failures(cntt, oracle):
    f = oracle.largest_prime_in_arithmetic_progression64
    n = 256
    p0, p1 = f(2 * n, 1, 0, 1 << 33), f(2 * n, 1, 0, 1 << 15)
    P = cntt.product.Plan
    assert P.try_new(n, 0, [p0, 0]) is None                       # src/product.rs:1155-1160
    assert P.try_new(n, p0 * p1 * p1 % 2**64, [p1, p0, p1]) is None   # src/product.rs:1162-1169
    assert P.try_new(n, p0 * p1 + 1, [p0, p1]) is None
    assert P.try_new(n, 15 * p1, [15, p1]) is None
    assert P.try_new(n, p0 * p1, [p0, p1, 1]) is not None


def test_product_pbs_shape_roundtrip_large(cntt, oracle, torch_cuda):
    """tfhe-rs NTT-PBS shape (product.rs:444-445: two primes < 2^31), N=2048, batch 4096: size-independent
    property inv(fwd(x)) * n^-1 == x, checked on the device output with Python integers on a sample and with the
    oracle on the first polynomials."""
    torch = torch_cuda
    n, batch = 2048, 4096
    f = oracle.largest_prime_in_arithmetic_progression64
    p0 = f(2 * n, 1, 0, 1 << 31)
    p1 = f(2 * n, 1, 0, p0 - 1)
    p = p0 * p1
    gp = cntt.product.Plan.try_new(n, p, [p0, p1])
    g = rng(4242)
    std = rand_mod(g, p, (batch, n), np.uint64)
    d = torch.zeros((batch, gp.ntt_domain_len()), dtype=torch.int64, device="cuda")
    gp.fwd(d, dev(torch, std))
    out = torch.zeros((batch, n), dtype=torch.int64, device="cuda")
    gp.inv(out, d, cntt.product.InvMode.Replace)
    got = host(out)
    n_inv = pow(n, -1, p)
    for b in (0, 1, batch // 2, batch - 1):
        assert [int(x) * n_inv % p for x in got[b]] == [int(x) for x in std[b]]
    # n * x mod p for every polynomial, vectorised through the two residues
    for q in (p0, p1):
        assert ((got % np.uint64(q)) == (std % np.uint64(q)) * np.uint64(n % q) % np.uint64(q)).all()


def test_product_host_slices(cntt, oracle):
    """The reference's call shape: numpy host slices through the *_host entry points (staged inside the library)."""
    n = 512
    ps = _product_cases(oracle, n)["u32x2_u64x1"]
    p = _prod(ps)
    op, gp = oracle.Product.try_new(n, p, ps), cntt.product.Plan.try_new(n, p, ps)
    g = rng(31337)
    dl = op.ntt_domain_len()
    a, b = rand_mod(g, p, (n,), np.uint64), rand_mod(g, p, (n,), np.uint64)
    ra, rb = np.zeros(dl, np.uint64), np.zeros(dl, np.uint64)
    ga, gb = np.zeros(dl, np.uint64), np.zeros(dl, np.uint64)
    op.fwd(ra, a, op.GENERIC); op.fwd(rb, b, op.GENERIC)
    gp.fwd(ga, a); gp.fwd(gb, b)
    assert (ga == ra).all() and (gb == rb).all()
    racc, gacc = ra.copy(), ga.copy()
    op.mul_accumulate(racc, ra, rb); gp.mul_accumulate(gacc, ga, gb)
    assert (gacc == racc).all()
    op.normalize(racc); gp.normalize(gacc)
    assert (gacc == racc).all()
    op.mul_assign_normalize(ra, rb); gp.mul_assign_normalize(ga, gb)
    assert (ga == ra).all()
    ro, go = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
    keep_r, keep_g = ra.copy(), ga.copy()
    op.inv(ro, ra, op.REPLACE); gp.inv(go, ga, cntt.product.InvMode.Replace)
    assert (go == ro).all() and (ga == ra).all()
    op.inv(ro, keep_r, op.ACCUMULATE); gp.inv(go, keep_g, cntt.product.InvMode.Accumulate)
    assert (go == ro).all()
    with pytest.raises(cntt.ReferencePanic):
        gp.fwd(np.zeros(dl + 1, np.uint64), a)
    # batch of 3 concatenated polynomials through the same host entry points
    A = rand_mod(g, p, (3, n), np.uint64)
    G = np.zeros((3, dl), np.uint64)
    gp.fwd(G, A)
    for i in range(3):
        r = np.zeros(dl, np.uint64)
        op.fwd(r, A[i], op.GENERIC)
        assert (G[i] == r).all()


@pytest.mark.parametrize("n", [256, 512, 1024, 4096])
@pytest.mark.parametrize("hi", [1 << 30, 1 << 31])
def test_product_fused_two_u32_primes(cntt, oracle, torch_cuda, n, hi):
    """The fused kernels of the hot shape (two u32 primes of one class, product_fused.hpp), both classes (< 2^30,
    < 2^31), every fused size, ragged batch: fwd Generic and Bounded, inv Replace (incl. the clobbered ntt) and
    Accumulate, against the oracle."""
    torch = torch_cuda
    f = oracle.largest_prime_in_arithmetic_progression64
    p0 = f(2 * n, 1, 0, hi)
    p1 = f(2 * n, 1, 0, p0 - 1)
    p = p0 * p1
    op, gp = oracle.Product.try_new(n, p, [p0, p1]), cntt.product.Plan.try_new(n, p, [p0, p1])
    dl, batch = op.ntt_domain_len(), 11
    g = rng(n + hi % 1000)
    std = rand_mod(g, p, (batch, n), np.uint64)
    std[0, :4] = [0, 1, p - 1, p // 2]
    ref = np.zeros((batch, dl), dtype=np.uint64)
    for b in range(batch):
        op.fwd(ref[b], std[b], op.GENERIC)
    d = torch.zeros((batch, dl), dtype=torch.int64, device="cuda")
    gp.fwd(d, dev(torch, std))
    assert (host(d) == ref).all()
    small = g.integers(-(1 << 20) + 1, 1 << 20, size=(batch, n))
    sstd = np.array([[int(x) % p for x in row] for row in small], dtype=np.uint64)
    refb = np.zeros((batch, dl), dtype=np.uint64)
    for b in range(batch):
        op.fwd(refb[b], sstd[b], op.bounded(1 << 20))
    db = torch.zeros_like(d)
    gp.fwd(db, dev(torch, sstd), cntt.product.FwdMode.Bounded(1 << 20))
    assert (host(db) == refb).all()
    out_ref = np.zeros((batch, n), dtype=np.uint64)
    keep = ref.copy()
    for b in range(batch):
        op.inv(out_ref[b], ref[b], op.REPLACE)
    dkeep = d.clone()
    out = torch.zeros((batch, n), dtype=torch.int64, device="cuda")
    gp.inv(out, d, cntt.product.InvMode.Replace)
    assert (host(out) == out_ref).all()
    assert (host(d) == ref).all()                      # inverse transforms left in ntt, like the reference
    acc0 = rand_mod(g, p, (batch, n), np.uint64)
    acc_ref = acc0.copy()
    for b in range(batch):
        op.inv(acc_ref[b], keep[b], op.ACCUMULATE)
    dacc = dev(torch, acc0)
    gp.inv(dacc, dkeep, cntt.product.InvMode.Accumulate)
    assert (host(dacc) == acc_ref).all()


@pytest.mark.parametrize("case", ["u64x1", "u32x1", "u32x2", "u30x2"])
def test_product_n8192(cntt, oracle, torch_cuda, case):
    """N = 8192: the u32 planes take the single-launch 8192-point kernel with a polynomial stride, the u64 planes the
    strided level + 4096-word blocks; same packed layout and results as the oracle (src/product.rs:261-353, 355-880)."""
    torch = torch_cuda
    n = 8192
    f = oracle.largest_prime_in_arithmetic_progression64
    d = 2 * n
    p0 = f(d, 1, 0, 2**32 - 1)
    q0 = f(d, 1, 0, 1 << 30)
    ps = {"u64x1": [f(d, 1, 0, 2**64 - 1)], "u32x1": [p0], "u32x2": [p0, f(d, 1, 0, p0 - 1)], "u30x2": [q0, f(d, 1, 0, q0 - 1)]}[case]
    p = _prod(ps)
    op, gp = oracle.Product.try_new(n, p, ps), cntt.product.Plan.try_new(n, p, ps)
    assert op is not None and gp is not None
    dl, batch = op.ntt_domain_len(), 3
    std = rand_mod(rng(8192 + len(case)), p, (batch, n), np.uint64)
    ref = np.zeros((batch, dl), dtype=np.uint64)
    out_ref = np.zeros((batch, n), dtype=np.uint64)
    for b in range(batch):
        op.fwd(ref[b], std[b], op.GENERIC)
    da = torch.zeros((batch, dl), dtype=torch.int64, device="cuda")
    gp.fwd(da, dev(torch, std))
    assert (host(da) == ref).all(), (case, "fwd")
    for b in range(batch):
        op.inv(out_ref[b], ref[b], op.REPLACE)
    dout = torch.zeros((batch, n), dtype=torch.int64, device="cuda")
    gp.inv(dout, da, cntt.product.InvMode.Replace)
    assert (host(dout) == out_ref).all(), (case, "inv")
