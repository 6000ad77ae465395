// build.rs for the `cuda` feature: compiles the kernel tree with nvcc for sm_100a and links it.
// Replaces the reference's build.rs (which only prints a deprecation warning, build.rs:11-22).
// NOT COMPILED in the build container (no Rust toolchain there).
use std::{env, path::PathBuf, process::Command};

fn main() {
    if env::var_os("CARGO_FEATURE_CUDA").is_none() {
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("cuda/csrc"); // = hades252_b200/csrc of the engine repository
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let units = [
        "hades_engine.cu", "hades_w3.cu", "hades_w5.cu", "hades_w9.cu", "hades_w3_dense.cu", "hades_w5_dense.cu",
        "hades_w9_dense.cu", "hades_w3_ccf.cu", "hades_w5_ccf.cu", "hades_w9_ccf.cu", "hades_generic.cu", "hades_frtest.cu",
    ];
    let mut objs = Vec::new();
    for u in units {
        let obj = out.join(u.replace(".cu", ".o"));
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"])
            .args(["-Xcompiler", "-fPIC", "-c", "-o"])
            .arg(&obj)
            .arg(csrc.join(u))
            .status()
            .expect("nvcc not found")
            .success();
        assert!(ok, "nvcc failed on {u}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(u).display());
    }
    let lib = out.join("libhades_b200.so");
    let ok = Command::new(&nvcc)
        .args(["-shared", "-cudart", "static", "-o"])
        .arg(&lib)
        .args(&objs)
        .args(["-ldl", "-lpthread"]) // libnccl.so.2 is dlopen'ed on demand by multi-device contexts
        .status()
        .unwrap()
        .success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=hades_b200");
}
