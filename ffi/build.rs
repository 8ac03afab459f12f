// Points rustc at the prebuilt CUDA library (tiny-object-detection_b200/lib/libtod_b200.so).
fn main() {
    let dir = std::env::var("TOD_B200_LIB_DIR").unwrap_or_else(|_| "../tiny-object-detection_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tod_b200");
    println!("cargo:rerun-if-env-changed=TOD_B200_LIB_DIR");
}
