// Links libact_b200.so (built by `make -C anonymous-credit-tokens_b200/csrc`, nvcc, sm_100a).
fn main() {
    let dir = std::env::var("ACT_B200_LIB_DIR").expect("set ACT_B200_LIB_DIR to the directory holding libact_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=act_b200");
    println!("cargo:rerun-if-env-changed=ACT_B200_LIB_DIR");
}
