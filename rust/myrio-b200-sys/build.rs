// Builds libmyrio_b200.so with nvcc for sm_100a and links it.  No other architecture,
// no CPU fallback.  MYRIO_B200_ROOT points at a checkout of this repository.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("MYRIO_B200_ROOT").expect("set MYRIO_B200_ROOT to the myrio-b200 checkout"));
    let csrc = root.join("myrio_b200").join("csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("make (nvcc) failed to start");
    assert!(status.success(), "nvcc build of libmyrio_b200.so failed");
    println!("cargo:rustc-link-search=native={}", root.join("myrio_b200").display());
    println!("cargo:rustc-link-lib=dylib=myrio_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").join("myrio_b200.h").display());
}
