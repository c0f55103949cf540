// Builds the CUDA side (include/lkm.h, linfa_b200/csrc/*.cu) into a static library with nvcc for
// sm_100a and links it together with the CUDA runtime.  NCCL is dlopen'ed by the library at
// lkm_comm_init time, so it is not a link-time dependency.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    // LKM_SRC points at the repository that holds include/lkm.h and linfa_b200/csrc
    let src = PathBuf::from(env::var("LKM_SRC").unwrap_or_else(|_| "../..".into()));
    let lib = out.join("liblkm.a");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objects = Vec::new();
    for unit in ["lkm_api", "lkm_extras"] {
        let cu = src.join(format!("linfa_b200/csrc/{unit}.cu"));
        let obj = out.join(format!("{unit}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
            .args(["-Xcompiler", "-fPIC", "-c", "-o"])
            .arg(&obj)
            .arg(&cu)
            .status()
            .expect("nvcc not found: the B200 K-Means path has no CPU fallback");
        assert!(status.success(), "nvcc failed");
        println!("cargo:rerun-if-changed={}", cu.display());
        objects.push(obj);
    }
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().expect("ar");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=lkm");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rerun-if-changed={}", src.join("include/lkm.h").display());
}
