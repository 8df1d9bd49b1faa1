"""Builds tests/_build/libsvo_emu.so: the product's .cu sources compiled by g++ against cuda_emu.h.
TEST INFRASTRUCTURE ONLY -- see cuda_emu.h.  Never imported by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sparsevoxeloctree_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_build", "libsvo_emu.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                                os.path.join(ROOT, "include", "svo.h")]
    newest = max(os.path.getmtime(p) for p in srcs)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["g++", "-x", "c++", "-std=c++20", "-O1", "-g", "-DSVO_EMU", "-ffp-contract=off", "-pthread", "-fPIC",
               "-shared", "-fvisibility=hidden", "-I" + HERE, "-I" + CSRC, "-o", OUT, os.path.join(CSRC, "svo_b200.cu")]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
