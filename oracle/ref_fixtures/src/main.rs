//! Reference fixtures: outputs of the real `concrete-ntt` crate on seeded inputs, one JSON object per line.
//! TEST INFRASTRUCTURE (oracle/): the records are consumed by tests/test_ref_fixtures.py, which regenerates the same
//! inputs (splitmix64, below) and compares the CPU oracle and the CUDA library with them bit for bit.
//! Record: {"kind", "n", "p"|"bits", "seed", then per output "<name>_hash" (FNV-1a over the u64-widened words),
//! "<name>_head" (first 8 words) and, for n == 32, "<name>_full"}.
use concrete_ntt::prime::largest_prime_in_arithmetic_progression64 as lpap;
use concrete_ntt::{native128, native32, native64, native_binary128, native_binary32, native_binary64, prime32, prime64, product};

struct Rng(u64);
impl Rng {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
}
fn fnv(words: impl Iterator<Item = u64>) -> u64 {
    let mut h = 0xcbf29ce484222325u64;
    for w in words {
        h ^= w;
        h = h.wrapping_mul(0x100000001b3);
    }
    h
}
fn field(name: &str, words: &[u64], full: bool) -> String {
    let list = |w: &[u64]| w.iter().map(|x| format!("\"{x}\"")).collect::<Vec<_>>().join(",");
    let mut s = format!("\"{name}_hash\":\"{}\",\"{name}_head\":[{}]", fnv(words.iter().copied()), list(&words[..8.min(words.len())]));
    if full {
        s += &format!(",\"{name}_full\":[{}]", list(words));
    }
    s
}
fn w32(v: &[u32]) -> Vec<u64> { v.iter().map(|&x| x as u64).collect() }
fn w128(v: &[u128]) -> Vec<u64> { v.iter().flat_map(|&x| [x as u64, (x >> 64) as u64]).collect() }

fn prime32_case(n: usize, p: u32, seed: u64) {
    let plan = match prime32::Plan::try_new(n, p) { Some(pl) => pl, None => { println!("{{\"kind\":\"prime32\",\"n\":{n},\"p\":\"{p}\",\"none\":true}}"); return; } };
    let mut r = Rng(seed);
    let mut gen = || (0..n).map(|_| (r.next() % p as u64) as u32).collect::<Vec<u32>>();
    let (a, b, c) = (gen(), gen(), gen());
    let mut f = a.clone();
    plan.fwd(&mut f);
    let mut i = f.clone();
    plan.inv(&mut i);
    let mut man = a.clone();
    plan.mul_assign_normalize(&mut man, &b);
    let mut nm = a.clone();
    plan.normalize(&mut nm);
    let mut acc = c.clone();
    plan.mul_accumulate(&mut acc, &a, &b);
    let full = n == 32;
    println!("{{\"kind\":\"prime32\",\"n\":{n},\"p\":\"{p}\",\"seed\":\"{seed}\",{},{},{},{},{}}}", field("fwd", &w32(&f), full), field("inv", &w32(&i), full),
             field("mul_assign_normalize", &w32(&man), false), field("normalize", &w32(&nm), false), field("mul_accumulate", &w32(&acc), false));
}
fn prime64_case(n: usize, p: u64, seed: u64) {
    let plan = match prime64::Plan::try_new(n, p) { Some(pl) => pl, None => { println!("{{\"kind\":\"prime64\",\"n\":{n},\"p\":\"{p}\",\"none\":true}}"); return; } };
    let mut r = Rng(seed);
    let mut gen = || (0..n).map(|_| r.next() % p).collect::<Vec<u64>>();
    let (a, b, c) = (gen(), gen(), gen());
    let mut f = a.clone();
    plan.fwd(&mut f);
    let mut i = f.clone();
    plan.inv(&mut i);
    let mut man = a.clone();
    plan.mul_assign_normalize(&mut man, &b);
    let mut nm = a.clone();
    plan.normalize(&mut nm);
    let mut acc = c.clone();
    plan.mul_accumulate(&mut acc, &a, &b);
    let full = n == 32;
    println!("{{\"kind\":\"prime64\",\"n\":{n},\"p\":\"{p}\",\"seed\":\"{seed}\",{},{},{},{},{}}}", field("fwd", &f, full), field("inv", &i, full),
             field("mul_assign_normalize", &man, false), field("normalize", &nm, false), field("mul_accumulate", &acc, false));
}
macro_rules! polymul_case {
    ($kind:expr, $bits:expr, $plan:ty, $word:ty, $n:expr, $seed:expr, $binary:expr, $widen:expr) => {{
        let n: usize = $n;
        let seed: u64 = $seed;
        if let Some(plan) = <$plan>::try_new(n) {
            let mut r = Rng(seed);
            let mut word = |binary: bool| -> $word {
                if $bits == 128 { let lo = r.next() as u128; let hi = r.next() as u128; let v = (hi << 64) | lo; (if binary { v & 1 } else { v }) as $word }
                else { let v = r.next(); (if binary { v & 1 } else { v }) as $word }
            };
            let lhs: Vec<$word> = (0..n).map(|_| word(false)).collect();
            let rhs: Vec<$word> = (0..n).map(|_| word($binary)).collect();
            let mut prod = vec![0 as $word; n];
            plan.negacyclic_polymul(&mut prod, &lhs, &rhs);
            println!("{{\"kind\":\"{}\",\"n\":{n},\"bits\":{},\"seed\":\"{seed}\",{}}}", $kind, $bits, field("prod", &$widen(&prod), n == 32));
        } else {
            println!("{{\"kind\":\"{}\",\"n\":{n},\"bits\":{},\"none\":true}}", $kind, $bits);
        }
    }};
}
fn native64_split_case(n: usize, seed: u64) {
    let plan = native64::Plan32::try_new(n).unwrap();
    let mut r = Rng(seed);
    let value: Vec<u64> = (0..n).map(|_| r.next()).collect();
    let mut m: Vec<Vec<u32>> = (0..5).map(|_| vec![0u32; n]).collect();
    {
        let (m0, rest) = m.split_at_mut(1);
        let (m1, rest) = rest.split_at_mut(1);
        let (m2, rest) = rest.split_at_mut(1);
        let (m3, m4) = rest.split_at_mut(1);
        plan.fwd(&value, &mut m0[0], &mut m1[0], &mut m2[0], &mut m3[0], &mut m4[0]);
    }
    let planes: Vec<u64> = m.iter().flat_map(|v| w32(v)).collect();
    let mut back = vec![0u64; n];
    {
        let (m0, rest) = m.split_at_mut(1);
        let (m1, rest) = rest.split_at_mut(1);
        let (m2, rest) = rest.split_at_mut(1);
        let (m3, m4) = rest.split_at_mut(1);
        plan.inv(&mut back, &mut m0[0], &mut m1[0], &mut m2[0], &mut m3[0], &mut m4[0]);
    }
    println!("{{\"kind\":\"native64_split\",\"n\":{n},\"bits\":64,\"seed\":\"{seed}\",{},{}}}", field("planes", &planes, false), field("inv", &back, false));
}
fn product_case(n: usize, seed: u64) {
    let p0 = lpap(2 * n as u64, 1, 0, 1 << 31).unwrap();
    let p1 = lpap(2 * n as u64, 1, 0, p0 - 1).unwrap();
    let modulus = p0 * p1;
    let plan = product::Plan::try_new(n, modulus, [p0, p1]).unwrap();
    let mut r = Rng(seed);
    let std: Vec<u64> = (0..n).map(|_| r.next() % modulus).collect();
    let mut ntt = vec![0u64; plan.ntt_domain_len()];
    plan.fwd(&mut ntt, &std, product::FwdMode::Generic);
    let dom = ntt.clone();
    let mut back = vec![0u64; n];
    plan.inv(&mut back, &mut ntt, product::InvMode::Replace);
    println!("{{\"kind\":\"product\",\"n\":{n},\"p0\":\"{p0}\",\"p1\":\"{p1}\",\"seed\":\"{seed}\",{},{}}}", field("fwd", &dom, false), field("inv", &back, false));
}

fn main() {
    println!("{{\"kind\":\"header\",\"crate\":\"concrete-ntt\",\"version\":\"0.2.0\",\"format\":1}}");
    let p32: Vec<u32> = vec![
        concrete_ntt::prime32::Plan::try_new(32, 1062862849).map(|p| p.modulus()).unwrap(), // primes32::P0 (< 2^30)
        lpap(1 << 16, 1, 1 << 29, 1 << 30).unwrap() as u32,
        lpap(1 << 16, 1, 1 << 30, 1 << 31).unwrap() as u32,
        lpap(1 << 16, 1, 1 << 31, 1 << 32).unwrap() as u32,
    ];
    let p64: Vec<u64> = vec![
        lpap(1 << 16, 1, 1 << 49, 1 << 50).unwrap(),
        lpap(1 << 16, 1, 1 << 50, 1 << 51).unwrap(),
        lpap(1 << 16, 1, 1 << 61, 1 << 62).unwrap(),
        lpap(1 << 16, 1, 1 << 62, 1 << 63).unwrap(),
        prime64::Solinas::P,
        lpap(1 << 16, 1, 1 << 63, u64::MAX).unwrap(),
    ];
    let mut seed = 0xC0FFEEu64;
    for &n in &[32usize, 1024, 4096, 32768] {
        for &p in &p32 { seed += 1; prime32_case(n, p, seed); }
        for &p in &p64 { seed += 1; prime64_case(n, p, seed); }
    }
    prime64_case(16, prime64::Solinas::P, 0xBEEF);
    for &n in &[32usize, 1024] {
        seed += 1; polymul_case!("native", 32, native32::Plan32, u32, n, seed, false, w32);
        seed += 1; polymul_case!("native", 64, native64::Plan32, u64, n, seed, false, |v: &Vec<u64>| v.clone());
        seed += 1; polymul_case!("native", 128, native128::Plan32, u128, n, seed, false, |v: &Vec<u128>| w128(v));
        seed += 1; polymul_case!("native_binary", 32, native_binary32::Plan32, u32, n, seed, true, w32);
        seed += 1; polymul_case!("native_binary", 64, native_binary64::Plan32, u64, n, seed, true, |v: &Vec<u64>| v.clone());
        seed += 1; polymul_case!("native_binary", 128, native_binary128::Plan32, u128, n, seed, true, |v: &Vec<u128>| w128(v));
    }
    native64_split_case(2048, 0xABCD);
    product_case(1024, 0x1234);
    product_case(2048, 0x1235);
}
