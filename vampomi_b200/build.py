"""Builds libvampomi_cuda.so and main_meth in-tree with nvcc for sm_100a (see csrc/Makefile)."""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "lib", "libvampomi_cuda.so")
MAIN_METH = os.path.join(PKG_DIR, "bin", "main_meth")
MAIN_METH_PROBIT = os.path.join(PKG_DIR, "bin", "main_meth_probit")


def build(verbose=False, jobs=None):
    env = dict(os.environ)
    if "/usr/local/cuda/bin" not in env.get("PATH", ""):
        env["PATH"] = "/usr/local/cuda/bin:" + env.get("PATH", "")
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libvampomi_cuda.so failed (see output above)")
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
