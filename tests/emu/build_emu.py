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


def build_sanitized():
    """The same source with -fsanitize=address,undefined (left shifts of negative values excepted: the LK fixed-point arithmetic restates OpenCV's
    `x << n` on signed operands, well defined on the GPU and for gcc): out-of-bounds indices of the kernels' local arrays, shared-memory carving and
    workspace offsets show up here, on the CPU box.  Returns (library, libasan to LD_PRELOAD) or None when the toolchain has no libasan."""
    out = os.path.join(HERE, "libviwb_emu_asan.so")
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        return None
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh", ".inl"))] + [os.path.join(ROOT, "include", "viwb.h")]
    if not (os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps)):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-DVIWB_HOST_EMU", "-x", "c++", "-Wno-unknown-pragmas",
                               "-DVIWB_EMU_STRICT", "-fsanitize=address,undefined", "-fno-sanitize=shift", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                               "-o", out, os.path.join(SRC, "viwb.cu")])
    return out, asan


if __name__ == "__main__":
    print(build())
