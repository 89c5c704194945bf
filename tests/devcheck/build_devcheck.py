"""TEST-ONLY: nvcc build of the device-side self checks (see sincos_check.cu). Same flags as the
classic-control kernels (`-fmad=false`), sm_100a."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(HERE, "_build", "sincos_check")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "sincos_check.cu")
    deps = [src, os.path.join(ROOT, "carl_b200", "csrc", "physics_classic.h"), os.path.join(ROOT, "carl_b200", "csrc", "rng.h")]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
                    "-I", os.path.join(ROOT, "include"), "-o", BIN, src], check=True)
    return BIN


if __name__ == "__main__":
    print(build())
