// Points the linker at the in-tree build of the library (3photons-rust_b200/_build/libtp3.so).
fn main() {
    let dir = std::env::var("TP3_LIB_DIR").unwrap_or_else(|_| "../../3photons-rust_b200/_build".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=tp3");
    println!("cargo:rerun-if-env-changed=TP3_LIB_DIR");
}
