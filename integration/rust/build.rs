// Links libbamscan.so.  BAMSCAN_LIB_DIR points at the directory that holds it (datafusion-bio-formats_b200/ in this repository).
fn main() {
    if let Ok(dir) = std::env::var("BAMSCAN_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=bamscan");
    println!("cargo:rerun-if-env-changed=BAMSCAN_LIB_DIR");
}
