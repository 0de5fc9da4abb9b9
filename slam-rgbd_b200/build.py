"""Builds the native libraries in-tree with make (nvcc -gencode arch=compute_100a,code=sm_100a)."""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def build_all(verbose: bool = False) -> None:
    env = dict(os.environ)
    env.setdefault("PATH", "")
    env["PATH"] = "/usr/local/cuda/bin:" + env["PATH"]
    res = subprocess.run(["make", "-C", PKG_DIR, "all"], env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("native build failed (see output above)")
    for rel in ("lib/libyouth_cuda.so", "lib/libAlgorithmModule.so", "lib/libyouth_synth.so", "bin/youth_harness", "bin/youth_multi"):
        if not os.path.exists(os.path.join(PKG_DIR, rel)):
            raise RuntimeError(f"native build did not produce {rel}")
