// ESIM_B200_LIB_DIR = the directory that holds libesim_b200.so and libesim_host.so (epidemicsimulator_b200/ of the build tree)
fn main() {
    let dir = std::env::var("ESIM_B200_LIB_DIR").expect("set ESIM_B200_LIB_DIR to the directory of libesim_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=esim_b200");
    println!("cargo:rustc-link-lib=dylib=esim_host");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=ESIM_B200_LIB_DIR");
}
