// Links the prebuilt CUDA library.  AWRY_B200_LIB_DIR points at the directory holding
// libawry_b200.so (awry_b200/ in this repository after `make -C awry_b200/csrc`).
fn main() {
    let dir = std::env::var("AWRY_B200_LIB_DIR").unwrap_or_else(|_| "../../awry_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=awry_b200");
    println!("cargo:rerun-if-env-changed=AWRY_B200_LIB_DIR");
}
