// Same pattern as cellular_raza-core/build.rs:9-19 (the `cara` placeholder), with the sm_100a arch flags
// that file lacks.  CRB_SRC may point at a checkout of this repository (default: two levels up).
use std::path::PathBuf;

fn main() {
    let root = std::env::var("CRB_SRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../.."));
    let csrc = root.join("cellular_raza_b200/csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/crb.h").display());
    cc::Build::new()
        .cuda(true)
        .ccbin(false)
        .cudart("shared")
        .flag("-O3")
        .flag("-std=c++17")
        .flag("-lineinfo")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .include(root.join("include"))
        .include(&csrc)
        .file(csrc.join("crb_api.cu")) // single translation unit: includes every .cuh
        .compile("crb200");
    println!("cargo:rustc-link-lib=cudart");
}
