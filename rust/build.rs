// Links libcntt_b200.so.  CNTT_B200_DIR = directory holding the library (concrete-ntt_b200/ of this repository
// after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("CNTT_B200_DIR").expect("set CNTT_B200_DIR to the directory containing libcntt_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cntt_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=CNTT_B200_DIR");
}
