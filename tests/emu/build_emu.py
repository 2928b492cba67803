"""Builds tests/emu/libviwb_emu.so: the DEVICE SOURCE of libviwb.so compiled with g++ (-DVIWB_HOST_EMU), every block
function executed by one host thread per block.  TEST INFRASTRUCTURE ONLY -- it lets the `not gpu` suite exercise the
kernels' indexing / algebra and the host lowering on the CPU-only CI box.  It is not built by __graft_entry__.build(),
not looked up by viwb.lib (which only loads csrc/libviwb.so), and is not a fallback for anything."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
SRC = os.path.join(ROOT, "viw-fusion_b200", "csrc")
OUT = os.path.join(HERE, "libviwb_emu.so")


def build():
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh", ".inl"))] + [os.path.join(ROOT, "include", "viwb.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-DVIWB_HOST_EMU", "-x", "c++", "-Wno-unknown-pragmas",
                           "-o", OUT, os.path.join(SRC, "viwb.cu")])
    return OUT


if __name__ == "__main__":
    print(build())
