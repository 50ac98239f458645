// Links the B200 library.  FRUID_B200_LIB_DIR must point at the directory holding
// libfruid_b200.so (fruid_b200/lib/ in this repository).
fn main() {
    let dir = std::env::var("FRUID_B200_LIB_DIR").expect("set FRUID_B200_LIB_DIR to the directory of libfruid_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fruid_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
