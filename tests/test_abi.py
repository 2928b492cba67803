"""`not gpu`: the C-ABI library loads and exports every symbol include/viwb.h declares (no compute calls),
the ctypes mirror has the C struct sizes, and the product refuses to run without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from viwb import abi, lib as viwb_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "viwb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(viwb_[a-z_0-9]+)\s*\(", src))
    return sorted(n for n in names if not n.startswith("viwb_block_"))     # static inline helpers


@pytest.fixture(scope="module")
def product_lib():
    if not os.path.exists(viwb_lib.DEFAULT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(viwb_lib.DEFAULT_LIB)


def test_every_declared_symbol_is_exported(product_lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(product_lib, s), "libviwb.so does not export %s" % s
    assert set(viwb_lib.EXPORTS) == set(syms)


def test_struct_sizes_match_the_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "viwb.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(viwb_prior), sizeof(viwb_globals), '
                    'sizeof(viwb_problem), sizeof(viwb_options), sizeof(viwb_summary)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(abi.Prior), C.sizeof(abi.Globals), C.sizeof(abi.Problem), C.sizeof(abi.Options), C.sizeof(abi.Summary)]


def test_block_tables_match_the_header():
    total_s = sum(abi.block_size(b) for b in range(abi.NUM_FIXED_BLOCKS))
    total_t = sum(abi.block_tsize(b) for b in range(abi.NUM_FIXED_BLOCKS))
    assert total_s == abi.STATE_FIXED and total_t == abi.TANGENT_FIXED
    for b in range(abi.NUM_FIXED_BLOCKS - 1):
        assert abi.block_offset(b) + abi.block_size(b) == abi.block_offset(b + 1)
        assert abi.block_toffset(b) + abi.block_tsize(b) == abi.block_toffset(b + 1)


def test_no_cpu_fallback_without_a_device(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(viwb_lib.ViwbError):
        viwb_lib.Context(0)


def test_product_does_not_reference_the_oracle():
    """only tests/, smoke() and bench.py's cpu_baseline legs may touch oracle/"""
    pkg = os.path.join(ROOT, "viw-fusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "viw_oracle" not in txt and "oracle/" not in txt.replace("not the oracle", ""), os.path.join(dp, f)
