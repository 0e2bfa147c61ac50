// raytracer/build.rs — links librtb200.so.  Written against include/rtb200.h; NOT compiled in the
// build image (no rustc/cargo there).
fn main() {
    // librtb200.so is built by `make -C ray_tracer_archive_b200/csrc` (nvcc -gencode arch=compute_100a,code=sm_100a)
    let dir = std::env::var("RTB200_LIB_DIR").expect("set RTB200_LIB_DIR to the directory holding librtb200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rtb200");
    println!("cargo:rerun-if-env-changed=RTB200_LIB_DIR");
}
