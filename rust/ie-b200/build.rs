// Links against the in-tree shared library built by `make -C interpolation_engine_b200/csrc`.
// IE_B200_LIB_DIR overrides the default location (the directory that holds libie_b200.so).
fn main() {
    let dir = std::env::var("IE_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../../interpolation_engine_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ie_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=IE_B200_LIB_DIR");
}
