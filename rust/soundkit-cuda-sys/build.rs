// Links libsoundkit_cuda.so; SOUNDKIT_CUDA_LIB_DIR points at the directory that holds it
// (soundkit_b200/ in this repository).
fn main() {
    if let Ok(dir) = std::env::var("SOUNDKIT_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=soundkit_cuda");
    println!("cargo:rerun-if-env-changed=SOUNDKIT_CUDA_LIB_DIR");
}
