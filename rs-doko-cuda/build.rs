fn main() {
    // DOKO_CUDA_LIB_DIR = directory that holds libdoko_cuda.so (master_doko_reinforcement_learning_b200/)
    if let Ok(dir) = std::env::var("DOKO_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=doko_cuda");
}
