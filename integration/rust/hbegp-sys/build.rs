// Links the prebuilt libhbegp.so (make -C hbetune_rs_b200/csrc).  HBEGP_LIB_DIR points at the directory holding it.
fn main() {
    let dir = std::env::var("HBEGP_LIB_DIR").expect("set HBEGP_LIB_DIR to the directory containing libhbegp.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=hbegp");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=HBEGP_LIB_DIR");
}
