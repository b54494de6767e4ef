// Links against a prebuilt libocrb.so (built by `make -C ocr_rs_b200/csrc`, sm_100a only).  OCRB_LIB_DIR points at the
// directory holding it; the default is the in-tree location relative to this crate.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("OCRB_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../ocr_rs_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=ocrb");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=OCRB_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/ocrb.h");
}
