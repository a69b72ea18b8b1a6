"""Host-side arithmetic behind the fused polymul kernels (no GPU): the primes a fused kernel carries the product on must
satisfy prod P_k > 2 |x| for every product coefficient x, and the quotient estimate of reconstruct_bounded
(csrc/native_device.cuh) needs |x| / M well below 1/2.  native128 runs on NINE of the reference's ten primes for
N <= 4096 (csrc/native.hpp, native_fused_np); every other kind uses the reference's prime count (src/native32.rs:8-12,
native64.rs:16-22, native128.rs:6-17, native_binary32.rs:11, native_binary64.rs:17-21, native_binary128.rs:4-10)."""
from fractions import Fraction

import pytest

FUSED_NP = {(32, False): 3, (64, False): 5, (128, False): 9, (32, True): 2, (64, True): 3, (128, True): 5}
REF_NP = {(32, False): 3, (64, False): 5, (128, False): 10, (32, True): 2, (64, True): 3, (128, True): 5}


@pytest.mark.parametrize("bits,binary", list(FUSED_NP))
def test_fused_prime_count_carries_every_coefficient(oracle, bits, binary):
    primes = [oracle.primes32(i) for i in range(10)]
    n = 4096                                   # largest fused size (kFusedMaxLogN = 12)
    bound = n * (2**bits - 1) * ((1 if binary else 2**bits - 1))     # max |coefficient|, attained by all-ones operands
    M = 1
    for p in primes[:FUSED_NP[(bits, binary)]]:
        M *= p
    assert 2 * bound < M
    # float32 quotient estimate: the true fraction x / M must stay clear of +-1/2 by far more than the float error
    assert Fraction(bound, M) < Fraction(3, 10)
    # one prime fewer would NOT do (the count is minimal), and the reference's own count is never smaller
    assert 2 * bound >= M // primes[FUSED_NP[(bits, binary)] - 1] or FUSED_NP[(bits, binary)] == 1
    assert FUSED_NP[(bits, binary)] <= REF_NP[(bits, binary)]


def test_native128_nine_primes_stop_at_4096(oracle):
    """The nine-prime product does not cover N = 8192 (|x| < 2^269 there), so the large-N path keeps all ten."""
    primes = [oracle.primes32(i) for i in range(10)]
    M9 = 1
    for p in primes[:9]:
        M9 *= p
    assert 2 * 8192 * (2**128 - 1) ** 2 > M9
    M10 = M9 * primes[9]
    assert 2 * 32768 * (2**128 - 1) ** 2 < M10
