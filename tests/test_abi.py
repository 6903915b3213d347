"""CPU-side checks of the C-ABI boundary: the shared library loads and exports every symbol that
include/smsut_b200.h declares, and the ctypes prototypes cover all of them (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smsut_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smsut_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = declared_symbols()
    assert "smsut_conv_tc" in names and "smsut_wgrad_tc" in names and len(names) >= 50


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_prototypes_cover_header(pkg):
    assert sorted(pkg._lib.exported_names()) == declared_symbols()


def test_abi_version_and_error_string(pkg):
    assert pkg._lib.lib.smsut_abi_version() == 1
    assert isinstance(pkg._lib.lib.smsut_last_error(), bytes)


def test_struct_layouts_match_header(pkg, tmp_path):
    # compile the header with gcc and compare struct sizes with the ctypes mirrors
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu\\n",'
                   'sizeof(smsut_conv_tc_args),sizeof(smsut_wgrad_tc_args),sizeof(smsut_conv_direct_args),'
                   'sizeof(smsut_pack_entry));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    L = pkg._lib
    assert sizes == [ctypes.sizeof(L.ConvTcArgs), ctypes.sizeof(L.WgradTcArgs), ctypes.sizeof(L.ConvDirectArgs),
                     ctypes.sizeof(L.PackEntry)]


def test_ops_refuse_cpu_tensors(pkg):
    import pytest
    import torch
    from smsut_b200 import ops
    with pytest.raises(pkg._lib.SmsutError):
        ops.in_stats(torch.zeros(1, 4, 4, 16, dtype=torch.bfloat16))
