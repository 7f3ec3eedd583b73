// Links the nvcc-built static library (make -C toy_cpu_pathtracing_b200/csrc -> lib/libtcpt.a) and the CUDA runtime.
// TCPT_LIB_DIR points at toy_cpu_pathtracing_b200/lib of the B200 backend checkout.
fn main() {
    let lib_dir = std::env::var("TCPT_LIB_DIR").expect("set TCPT_LIB_DIR to <b200 backend>/toy_cpu_pathtracing_b200/lib");
    let cuda = std::env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={lib_dir}");
    println!("cargo:rustc-link-lib=static=tcpt");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-env-changed=TCPT_LIB_DIR");
}
