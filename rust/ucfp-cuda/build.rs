// Builds libucfp_cuda from the CUDA sources of this repository with nvcc for sm_100a and links it.
// NOT compiled in the development container of this repository (no cargo/rustc there); the same sources are
// built by `python -m ucfp_b200.build`, which is what the test-suite and bench exercise.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("ucfp_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objs = Vec::new();
    for (src, extra) in [("api.cu", ""), ("corpus.cu", ""), ("batcher.cu", ""), ("group.cu", ""), ("multihash.cu", ""), ("jpeg.cu", ""),
                         ("hamming.cu", ""), ("jaccard.cu", ""), ("cosine.cu", "-fmad=false"), ("image.cu", "-fmad=false"), ("merge.cu", "")] {
        let obj = out.join(src.replace(".cu", ".o"));
        let mut c = Command::new(&nvcc);
        c.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                "--expt-relaxed-constexpr", "-c"]);
        if !extra.is_empty() { c.arg(extra); }
        c.arg(csrc.join(src)).arg("-o").arg(&obj);
        assert!(c.status().expect("nvcc not found").success(), "nvcc failed on {src}");
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
        objs.push(obj);
    }
    let lib = out.join("libucfp_cuda.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=ucfp_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=cuda");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    // NCCL (multi-GPU groups) and nvJPEG (device-side JPEG decode) are resolved with dlopen at first use; link them here as well
    // when the host wants them pinned at load time:
    if env::var("UCFP_LINK_NCCL").is_ok() { println!("cargo:rustc-link-lib=dylib=nccl"); }
    if env::var("UCFP_LINK_NVJPEG").is_ok() { println!("cargo:rustc-link-lib=dylib=nvjpeg"); }
}
