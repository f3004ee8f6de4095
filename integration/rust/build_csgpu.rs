// build_csgpu.rs — `include!("build_csgpu.rs");` from build.rs and call `build_csgpu()` in main() (see build.rs.patch).
//
// Compiles the vendored CUDA sources (this repository's codesearch_b200/csrc + include/, copied to <crate>/csgpu/) into
// $OUT_DIR/libcsgpu.so for sm_100a ONLY and links it. No nvcc => the build fails: the GPU vector store has no CPU
// fallback. Every translation unit of the library is listed (tests/test_rust_binding.py keeps this list equal to the
// Makefile's OBJS; round 1's snippet compiled csgpu.cu alone and could not have linked).
const CSGPU_SOURCES: &[&str] = &["csgpu.cu", "scan_multi.cu", "scan_filtered.cu", "gemm_topk.cu", "snapshot.cu", "scan_i8.cu"];

fn build_csgpu() {
    use std::path::PathBuf;
    use std::process::Command;
    let root = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("csgpu");
    let csrc = root.join("csrc");
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    println!("cargo:rerun-if-env-changed=NVCC");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").display());
    let mut objects = Vec::new();
    let mut children = Vec::new();
    for src in CSGPU_SOURCES {
        let obj = out.join(format!("{}.o", src.trim_end_matches(".cu")));
        let child = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
                   "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .spawn()
            .expect("nvcc not found (set NVCC): the GPU vector store has no CPU fallback");
        children.push((src, child));
        objects.push(obj);
    }
    for (src, mut child) in children {
        assert!(child.wait().expect("nvcc did not run").success(), "nvcc failed on {src}");
    }
    let status = Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"])
        .args(&objects)
        .arg("-o")
        .arg(out.join("libcsgpu.so"))
        .args(["-ldl", "-lpthread"])
        .status()
        .expect("nvcc link step did not run");
    assert!(status.success(), "linking libcsgpu.so failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=csgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
