//! Compiles the CUDA sources with `nvcc -gencode arch=compute_100a,code=sm_100a` into a static library and links
//! it (with cudart and the C++ runtime).  The translation units are the list in `translation_units.txt`, which
//! `rust/gen_sys.py` derives from `radiorust_b200/build.py` -- the same objects the Python build produces.
//!
//! Environment:
//!   CUDA_HOME                 CUDA toolkit root (default /usr/local/cuda); nvcc >= 12.8 for sm_100a
//!   RADIORUST_B200_LIB_DIR    with feature `prebuilt`: directory that holds libradiorust_b200.a
//!   RADIORUST_B200_JOBS       parallel nvcc processes (default: available parallelism)
use std::path::{Path, PathBuf};
use std::process::Command;
use std::sync::{Arc, Mutex};
use std::{env, fs, thread};

const COMMON_FLAGS: &[&str] = &[
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
];

struct Unit {
    object: String,
    source: String,
    flags: Vec<String>,
}

fn units(manifest: &Path) -> Vec<Unit> {
    let text = fs::read_to_string(manifest.join("translation_units.txt")).expect("translation_units.txt");
    text.lines()
        .filter(|l| !l.trim().is_empty() && !l.starts_with('#'))
        .map(|l| {
            let mut it = l.split_whitespace().map(str::to_owned);
            Unit { object: it.next().unwrap(), source: it.next().unwrap(), flags: it.collect() }
        })
        .collect()
}

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
    println!("cargo:rerun-if-env-changed=RADIORUST_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=translation_units.txt");
    println!("cargo:rerun-if-changed=csrc");
    println!("cargo:rerun-if-changed=include");

    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("RADIORUST_B200_LIB_DIR").expect("feature `prebuilt` needs RADIORUST_B200_LIB_DIR");
        println!("cargo:rustc-link-search=native={dir}");
    } else {
        let csrc = manifest.join("csrc");
        let nvcc = format!("{cuda}/bin/nvcc");
        let work = Arc::new(Mutex::new(units(&manifest)));
        let objects = Arc::new(Mutex::new(Vec::<PathBuf>::new()));
        let jobs = env::var("RADIORUST_B200_JOBS")
            .ok()
            .and_then(|v| v.parse::<usize>().ok())
            .unwrap_or_else(|| thread::available_parallelism().map(|n| n.get()).unwrap_or(4));
        let mut handles = Vec::new();
        for _ in 0..jobs.max(1) {
            let (work, objects, csrc, out, nvcc) = (work.clone(), objects.clone(), csrc.clone(), out.clone(), nvcc.clone());
            handles.push(thread::spawn(move || loop {
                let Some(unit) = work.lock().unwrap().pop() else { return };
                let obj = out.join(&unit.object);
                let status = Command::new(&nvcc)
                    .args(COMMON_FLAGS)
                    .args(&unit.flags)
                    .arg("-c")
                    .arg(csrc.join(&unit.source))
                    .arg("-o")
                    .arg(&obj)
                    .status()
                    .unwrap_or_else(|e| panic!("cannot run {nvcc}: {e} (there is no CPU fallback: the CUDA toolkit is required)"));
                assert!(status.success(), "nvcc failed for {} {:?}", unit.source, unit.flags);
                objects.lock().unwrap().push(obj);
            }));
        }
        for h in handles {
            h.join().expect("nvcc worker");
        }
        let lib = out.join("libradiorust_b200.a");
        let _ = fs::remove_file(&lib);
        let objs = objects.lock().unwrap();
        assert!(Command::new("ar").arg("rcs").arg(&lib).args(objs.iter()).status().expect("ar").success(), "ar failed");
        println!("cargo:rustc-link-search=native={}", out.display());
    }
    println!("cargo:rustc-link-lib=static=radiorust_b200");
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    // the header, for dependents that want to run bindgen themselves
    println!("cargo:include={}", manifest.join("include").display());
}
