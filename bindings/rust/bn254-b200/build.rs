// Links against libbn254_b200.so.  BN254_B200_LIB_DIR names the directory that holds it (the repository's bn254_b200/).
fn main() {
    if let Ok(dir) = std::env::var("BN254_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=bn254_b200");
    println!("cargo:rerun-if-env-changed=BN254_B200_LIB_DIR");
}
