#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY — builds oracle/_ref/main_meth_ref from the reference's own sources.

The reference (/root/reference/src/{main_meth,vamp,utilities,data,options}.cpp; vamp_probit.cpp is textually
included by vamp.cpp:13) is compiled where it lies against oracle/ref_shims/ (single-rank mpi.h, Boost stand-ins).
It cannot run as shipped (SURVEY.md "five facts" #2/#3), so four scripted edits are applied to a *temporary*
copy that never enters the repo; only the binary lands in oracle/_ref/ (git-ignored, travels to the GPU box):

  P1  vamp.cpp:70,77      size x1_hat / r1 to M (un-comment the authors' own lines)      — otherwise OOB write at :205
  P2  vamp.cpp:296, vamp_probit.cpp:298   Hutchinson probe from a counter hash of (seed, it, S+i) instead of
      std::random_device — sharding-invariant and reproducible (seed from env VAMPOMI_SEED)
  P3  vamp_probit.cpp:53  probit start p1 from a hashed Box-Muller of (seed, i)
  P4  data.cpp:297,351    size_t column offsets so one rank can hold N*M >= 2^31
  P5  main_meth.cpp:48    the inference branch calls dataset.read_covariates(--cov-file, --C) after constructing the dataset: the
      shipped main never loads covariates, so `--C > 0` would index an empty matrix; no effect when C = 0

Flags follow README.md:28 minus -D_GLIBCXX_DEBUG -g, and -march=x86-64-v3 instead of native because the binary is
built in this container but timed on the GPU box's host CPU.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("VAMPOMI_REFERENCE_SRC", "/root/reference/src")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_BIN = os.path.join(OUT_DIR, "main_meth_ref")
OUT_BIN_STRICT = os.path.join(OUT_DIR, "main_meth_ref_O2")      # IEEE-strict build (-O2, no fast-math), see DESIGN.md "parity floor"
OUT_BIN_V4 = os.path.join(OUT_DIR, "main_meth_ref_v4")         # README flags for an AVX-512 host (-march=x86-64-v4): the TIMING binary


OUT_BIN_GPUDATA = os.path.join(OUT_DIR, "main_meth_ref_gpudata")   # the reference's main + vamp over OUR `class data` (ref_shims/data_gpu.cpp)


def build_gpu_data(force=False, verbose=True):
    """The reference's own main_meth.cpp / vamp.cpp / utilities.cpp / options.cpp (patches P1-P3) linked against the `class data`
    adapter of INTEGRATION.md §2 (ref_shims/data_gpu.cpp) and libvampomi_cuda.so instead of src/data.cpp."""
    if not available():
        return os.path.isfile(OUT_BIN_GPUDATA)
    root = os.path.dirname(HERE)
    lib = os.path.join(root, "vampomi_b200", "lib", "libvampomi_cuda.so")
    if not os.path.isfile(lib):
        raise RuntimeError("libvampomi_cuda.so is not built yet")
    deps = [os.path.abspath(__file__), os.path.join(HERE, "ref_shims", "data_gpu.cpp"), os.path.join(root, "include", "vampomi.h")]
    if not force and os.path.isfile(OUT_BIN_GPUDATA) and all(os.path.getmtime(d) <= os.path.getmtime(OUT_BIN_GPUDATA) for d in deps):
        return True
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vampomi_ref_")
    try:
        for f in os.listdir(REF_SRC):
            if f.endswith((".cpp", ".hpp")):
                with open(os.path.join(REF_SRC, f)) as fh:
                    text = fh.read()
                for old, new, count in PATCHES.get(f, []):
                    text = text.replace(old, new)
                with open(os.path.join(tmp, f), "w") as fh:
                    fh.write(text)
        shims = os.path.join(HERE, "ref_shims")
        cmd = ["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-fopenmp", "-w", "-I", shims, "-I", tmp, "-I", os.path.join(root, "include"),
               "-include", os.path.join(shims, "oracle_hooks.h")]
        cmd += [os.path.join(tmp, u) for u in UNITS if u != "data.cpp"] + [os.path.join(shims, "data_gpu.cpp")]
        cmd += ["-L", os.path.dirname(lib), "-lvampomi_cuda", "-Wl,-rpath,$ORIGIN/../../vampomi_b200/lib", "-o", OUT_BIN_GPUDATA + ".tmp"]
        if verbose:
            print("[oracle/_ref] " + " ".join(cmd))
        subprocess.run(cmd, check=True)
        os.replace(OUT_BIN_GPUDATA + ".tmp", OUT_BIN_GPUDATA)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


def timing_binary():
    """The reference build to TIME on this host: the README asks for -march=native (README.md:28), but the binary is built in
    the build container and timed on the GPU box's CPU, so two ISA levels are built and the widest one this CPU runs is chosen
    (x86-64-v4 = AVX-512 F/BW/CD/DQ/VL, else x86-64-v3 = AVX2 + FMA, the build the fixtures were made with)."""
    try:
        with open("/proc/cpuinfo") as f:
            flags = set(next(l for l in f if l.startswith("flags")).split(":", 1)[1].split())
    except Exception:
        flags = set()
    if {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags and os.path.isfile(OUT_BIN_V4):
        return OUT_BIN_V4, "-Ofast -march=x86-64-v4 (AVX-512: this host's level; README.md:28 asks for -march=native)"
    return OUT_BIN, "-Ofast -march=x86-64-v3 (AVX2 + FMA)"


PATCHES = {
    "main_meth.cpp": [
        ("        // Initialize model hyperparameters\n", "        dataset.read_covariates(opt.get_cov_file(), C);\n        // Initialize model hyperparameters\n", 1),
    ],
    "vamp.cpp": [
        ("//x1_hat = std::vector<double> (M, 0.0);", "x1_hat = std::vector<double> (M, 0.0);", 1),
        ("//r1 = std::vector<double> (M, 0.0);", "r1 = std::vector<double> (M, 0.0);", 1),
        ("bern_vec[i] = (2*bern(rd) - 1) / sqrt(Mt);", "bern_vec[i] = vampomi_oracle_probe(it, (long) S + i) / sqrt(Mt);", 1),
    ],
    "vamp_probit.cpp": [
        ("bern_vec[i] = (2*bern(rd) - 1) / sqrt(Mt);", "bern_vec[i] = vampomi_oracle_probe(it, (long) S + i) / sqrt(Mt);", 1),
        ("p1 = simulate(N, std::vector<double> {1.0}, std::vector<double> {1.0});", "p1 = vampomi_oracle_p1(N);", 1),
    ],
    "data.cpp": [
        ("&meth_data[mloc * N]", "&meth_data[size_t(mloc) * size_t(N)]", 1),
        ("&meth_data[i * N]", "&meth_data[size_t(i) * size_t(N)]", 1),
    ],
}
UNITS = ["main_meth.cpp", "vamp.cpp", "utilities.cpp", "data.cpp", "options.cpp"]


def available():
    return os.path.isdir(REF_SRC) and all(os.path.isfile(os.path.join(REF_SRC, u)) for u in UNITS)


def up_to_date(out_bin=None):
    out_bin = out_bin or OUT_BIN
    if not os.path.isfile(out_bin):
        return False
    t = os.path.getmtime(out_bin)
    deps = [os.path.abspath(__file__)]
    for root, _, files in os.walk(os.path.join(HERE, "ref_shims")):
        deps += [os.path.join(root, f) for f in files]
    if available():
        deps += [os.path.join(REF_SRC, f) for f in os.listdir(REF_SRC)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=True, strict=False, v4=False):
    out_bin = OUT_BIN_STRICT if strict else OUT_BIN_V4 if v4 else OUT_BIN
    opt = ["-O2"] if strict else ["-Ofast"]
    march = "-march=x86-64-v4" if v4 else "-march=x86-64-v3"
    if not available():
        if verbose:
            print(f"[oracle/_ref] reference sources not present at {REF_SRC}; keeping prebuilt binary (if any)")
        return os.path.isfile(out_bin)
    if not force and up_to_date(out_bin):
        return True
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vampomi_ref_")
    try:
        for f in os.listdir(REF_SRC):
            if f.endswith((".cpp", ".hpp")):
                with open(os.path.join(REF_SRC, f)) as fh:
                    text = fh.read()
                for old, new, count in PATCHES.get(f, []):
                    if text.count(old) != count:
                        raise RuntimeError(f"patch anchor {old!r} found {text.count(old)}x in {f}, expected {count}")
                    text = text.replace(old, new)
                with open(os.path.join(tmp, f), "w") as fh:
                    fh.write(text)
        shims = os.path.join(HERE, "ref_shims")
        cmd = ["g++", "-std=c++17"] + opt + [march, "-fopenmp", "-w",
               "-I", shims, "-include", os.path.join(shims, "oracle_hooks.h")]
        cmd += [os.path.join(tmp, u) for u in UNITS]
        cmd += ["-o", out_bin + ".tmp"]
        if verbose:
            print("[oracle/_ref] " + " ".join(cmd))
        subprocess.run(cmd, check=True)
        os.replace(out_bin + ".tmp", out_bin)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    if "--strict" in sys.argv:
        ok = build(force="--force" in sys.argv, strict=True) and ok
    ok = build(force="--force" in sys.argv, v4=True) and ok
    if "--gpu-data" in sys.argv:
        ok = build_gpu_data(force="--force" in sys.argv) and ok
    print(OUT_BIN if ok else "unavailable")
    sys.exit(0 if ok else 1)
