// raytracer/build.rs — builds librtb200.so with nvcc (through the library's own Makefile) and links it.
// The reference crate has no build script (raytracer/Cargo.toml:1-12); this file and the `[build-dependencies]`-free
// `build = "build.rs"` line are the only Cargo-level additions.  NOT compiled in the build image (no rustc/cargo there);
// the C++ mirror (include/rtb200_scene.hpp + tests/cpp/test_scene_mirror.cpp) pushes byte-identical records through the
// same C ABI and IS tested.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    // RTB200_DIR = checkout of the rtb200 repository (holds ray_tracer_archive_b200/csrc and include/rtb200.h)
    let root = PathBuf::from(env::var("RTB200_DIR").expect("set RTB200_DIR to the rtb200 checkout"));
    let csrc = root.join("ray_tracer_archive_b200").join("csrc");
    // nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (csrc/Makefile); NVCC may point at a specific nvcc
    let status = Command::new("make")
        .arg("-C")
        .arg(&csrc)
        .arg("-j")
        .status()
        .expect("could not run make (is the CUDA toolkit installed?)");
    assert!(status.success(), "building librtb200.so failed");
    let lib_dir = root.join("ray_tracer_archive_b200");
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=rtb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:rerun-if-env-changed=RTB200_DIR");
    for f in ["rtb_api.cu", "rtb_kernels.cu", "rtb_device.cuh", "rtb_internal.hpp", "rtb_launch.hpp", "rtb_nccl.hpp", "flatten.cpp", "bvh_build.cpp"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("rtb200.h").display());
}
