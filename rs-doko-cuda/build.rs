fn main() {
    // DOKO_CUDA_LIB_DIR = directory that holds libdoko_cuda.so (master_doko_reinforcement_learning_b200/);
    // CUDA_HOME/lib64 holds libcudart.so (device memory for the batch wrapper).
    if let Ok(dir) = std::env::var("DOKO_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    let cuda = std::env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=dylib=doko_cuda");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-env-changed=DOKO_CUDA_LIB_DIR");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
}
