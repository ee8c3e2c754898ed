// UNCOMPILED (no Rust toolchain in the build image).  Links the in-tree libvxp.so.
fn main() {
    let dir = std::env::var("VXP_LIB_DIR").unwrap_or_else(|_| "../../voxelphile_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=vxp");
    println!("cargo:rerun-if-env-changed=VXP_LIB_DIR");
}
