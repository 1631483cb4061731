// Links the CUDA library built by `python -c "import __graft_entry__ as g; g.build()"`.
fn main() {
    let dir = std::env::var("B200DF_LIB_DIR").expect("set B200DF_LIB_DIR to the directory that holds libb200df.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=b200df");
    println!("cargo:rerun-if-env-changed=B200DF_LIB_DIR");
}
