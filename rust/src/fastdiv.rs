//! `concrete_ntt::fastdiv` -- division by a constant (reference: src/fastdiv.rs:28-150).
//!
//! Host-only plan-time helpers.  The reference precomputes a double-word reciprocal so that `div` / `rem` need no hardware
//! division; what is observable is the exact quotient and remainder, which is what these return.  The public fields of
//! the reference's structs (`double_reciprocal`, `single_reciprocal`, `divisor`) are kept, with the same values:
//! `floor(2^(2w) / d) + 1`-style reciprocals as computed in `Div32::new` / `Div64::new`.

/// src/fastdiv.rs:28-34
#[derive(Copy, Clone, Debug)]
pub struct Div32 {
    pub double_reciprocal: u128,
    pub single_reciprocal: u64,
    pub divisor: u32,
}

/// src/fastdiv.rs:36-41.  The reference stores a 256-bit `double_reciprocal` in its private `u256` type; the facade keeps
/// its four little-endian limbs.
#[derive(Copy, Clone, Debug)]
pub struct Div64 {
    pub double_reciprocal: [u64; 4],
    pub single_reciprocal: u128,
    pub divisor: u64,
}

impl Div32 {
    /// src/fastdiv.rs:48-60 -- panics for `divisor <= 1` like the reference's `assert!(divisor > 1)`.
    pub const fn new(divisor: u32) -> Self {
        assert!(divisor > 1);
        let single_reciprocal = (u64::MAX / divisor as u64) + 1;
        let double_reciprocal = (u128::MAX / divisor as u128) + 1;
        Self { double_reciprocal, single_reciprocal, divisor }
    }
    /// src/fastdiv.rs:62-66
    #[inline(always)]
    pub const fn div(n: u32, d: Self) -> u32 {
        n / d.divisor
    }
    /// src/fastdiv.rs:68-73
    #[inline(always)]
    pub const fn rem(n: u32, d: Self) -> u32 {
        n % d.divisor
    }
    /// src/fastdiv.rs:75-79
    #[inline(always)]
    pub const fn div_u64(n: u64, d: Self) -> u64 {
        n / d.divisor as u64
    }
    /// src/fastdiv.rs:81-86
    #[inline(always)]
    pub const fn rem_u64(n: u64, d: Self) -> u32 {
        (n % d.divisor as u64) as u32
    }
    /// src/fastdiv.rs:88-90
    #[inline(always)]
    pub const fn divisor(&self) -> u32 {
        self.divisor
    }
}

impl Div64 {
    /// src/fastdiv.rs:98-119 -- panics for `divisor <= 1`.
    pub const fn new(divisor: u64) -> Self {
        assert!(divisor > 1);
        let single_reciprocal = (u128::MAX / divisor as u128) + 1;
        // floor((2^256 - 1) / d) + 1 by schoolbook long division on 64-bit limbs, most significant first
        let mut q = [0u64; 4];
        let mut rem: u128 = 0;
        let mut i = 4;
        while i > 0 {
            i -= 1;
            let cur = (rem << 64) | u64::MAX as u128;
            q[i] = (cur / divisor as u128) as u64;
            rem = cur % divisor as u128;
        }
        // + 1 with carry
        let mut j = 0;
        while j < 4 {
            let (v, c) = q[j].overflowing_add(1);
            q[j] = v;
            if !c {
                break;
            }
            j += 1;
        }
        Self { double_reciprocal: q, single_reciprocal, divisor }
    }
    /// src/fastdiv.rs:121-125
    #[inline(always)]
    pub const fn div(n: u64, d: Self) -> u64 {
        n / d.divisor
    }
    /// src/fastdiv.rs:127-132
    #[inline(always)]
    pub const fn rem(n: u64, d: Self) -> u64 {
        n % d.divisor
    }
    /// src/fastdiv.rs:134-138
    #[inline(always)]
    pub const fn div_u128(n: u128, d: Self) -> u128 {
        n / d.divisor as u128
    }
    /// src/fastdiv.rs:140-145
    #[inline(always)]
    pub const fn rem_u128(n: u128, d: Self) -> u64 {
        (n % d.divisor as u128) as u64
    }
    /// src/fastdiv.rs:147-149
    #[inline(always)]
    pub const fn divisor(&self) -> u64 {
        self.divisor
    }
}
