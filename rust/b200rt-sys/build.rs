// UNCOMPILED SOURCE (no Rust toolchain in the build image).
fn main() {
    // directory that holds libb200rt.so (homework-18-graphics-raytracer_b200/_lib after `__graft_entry__.build()`)
    if let Ok(dir) = std::env::var("B200RT_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=b200rt");
    println!("cargo:rerun-if-env-changed=B200RT_LIB_DIR");
}
