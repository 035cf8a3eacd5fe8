// Link against the C ABI library.  TMC2GPU_LIB_DIR = directory that holds libtmc2gpu.so
// (built by `python -c "import __graft_entry__ as g; g.build()"` -> tmc2-rs_b200/libtmc2gpu.so).
fn main() {
    let dir = std::env::var("TMC2GPU_LIB_DIR").expect("set TMC2GPU_LIB_DIR to the directory of libtmc2gpu.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tmc2gpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=TMC2GPU_LIB_DIR");
}
