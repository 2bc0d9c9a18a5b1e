// Links against libtexpresso_b200.so (built by `python -m texpresso_b200.build`).
// TEXPRESSO_B200_LIB_DIR must point at the directory that holds it.
fn main() {
    let dir = std::env::var("TEXPRESSO_B200_LIB_DIR").expect("set TEXPRESSO_B200_LIB_DIR to the directory of libtexpresso_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=texpresso_b200");
    println!("cargo:rerun-if-env-changed=TEXPRESSO_B200_LIB_DIR");
}
