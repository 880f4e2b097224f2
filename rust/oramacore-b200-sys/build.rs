// NOTE: source-only crate — the build image has no cargo/rustc, so this was never compiled here.
fn main() {
    let dir = std::env::var("ORAMACORE_B200_LIB_DIR").unwrap_or_else(|_| "../../oramacore_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=oramacore_b200");
}
