// Link against the in-tree shared library built by `make -C mental-poker_b200/csrc`.
fn main() {
    let dir = std::env::var("MPSHUFFLE_LIB_DIR").unwrap_or_else(|_| "../../mental-poker_b200/lib".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=mpshuffle");
}
