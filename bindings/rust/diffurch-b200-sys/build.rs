// Point rustc at the directory that holds libdiffurch_b200.so (built by `python __graft_entry__.py`).
fn main() {
    let dir = std::env::var("DIFFURCH_B200_LIB_DIR").unwrap_or_else(|_| "../../../diffurch_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=diffurch_b200");
    println!("cargo:rerun-if-env-changed=DIFFURCH_B200_LIB_DIR");
}
