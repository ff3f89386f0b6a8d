// BZ2B200_LIB_DIR = directory that holds libbz2b200.so (bzip2_rust_b200/ after `python -m bzip2_rust_b200.build`)
fn main() {
    let dir = std::env::var("BZ2B200_LIB_DIR").expect("set BZ2B200_LIB_DIR to the directory of libbz2b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=bz2b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=BZ2B200_LIB_DIR");
}
