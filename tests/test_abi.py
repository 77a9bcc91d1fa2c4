"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the public header
declares, and the header-generated ctypes structs are sane.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from mednet_b200 import _abi


def test_library_loads_and_exports_every_declared_symbol():
    lib = _abi.lib()
    text = open(_abi.HEADER_PATH).read()
    declared = set(re.findall(r"\b(mednet_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", " ", text, flags=re.S)))
    assert len(declared) >= 40
    assert declared == set(_abi.FUNCTIONS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mednet_abi_version() == _abi.MEDNET_ABI_VERSION
    assert lib.mednet_error_string(-2).decode().startswith("unsupported")


def test_struct_layouts_follow_the_header(tmp_path):
    """sizeof/offsetof of every ctypes struct equal what gcc computes from the header itself."""
    import subprocess
    s = _abi.STRUCTS["mednet_conv3d_params"]
    names = [f for f, _ in s._fields_]
    assert names[:5] == ["x", "w", "bias", "addend", "y"] and names[-3:] == ["impl", "y_f32", "addend_f32"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mednet_b200.h"', 'int main(void) {']
    for name, fields in _abi._STRUCT_FIELDS.items():
        lines.append(f'  printf("{name} %zu %zu\\n", sizeof({name}), offsetof({name}, {fields[-1][0]}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "sizes.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.dirname(_abi.HEADER_PATH), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        if not line.strip():
            continue
        name, size, last_off = line.split()
        st = _abi.STRUCTS[name]
        assert ctypes.sizeof(st) == int(size), name
        assert getattr(st, _abi._STRUCT_FIELDS[name][-1][0]).offset == int(last_off), name
        seen += 1
    assert seen == len(_abi.STRUCTS) >= 26


def test_host_side_queries_need_no_gpu():
    lib = _abi.lib()
    p = _abi.make("mednet_gn_fwd_params", N=2, S=4096, C=64, G=8, dtype=_abi.MEDNET_BF16)
    assert lib.mednet_groupnorm_fwd_workspace_bytes(ctypes.byref(p)) > 0
    q = _abi.make("mednet_conv3d_params", N=1, Di=8, Hi=8, Wi=8, Do=8, Ho=8, Wo=8, K=64, Nout=64, dtype=_abi.MEDNET_F32,
                  impl=_abi.MEDNET_IMPL_AUTO)
    assert lib.mednet_conv3d_select_impl(ctypes.byref(q)) == _abi.MEDNET_IMPL_SIMT      # fp32 -> validation path
    q.impl = _abi.MEDNET_IMPL_TCGEN05
    assert lib.mednet_conv3d_select_impl(ctypes.byref(q)) == _abi.MEDNET_EUNSUPPORTED   # explicit impl never falls back
    with pytest.raises(KeyError):
        _abi.make("mednet_conv3d_params", nonsense=1)


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", os.path.join(os.path.dirname(_abi.LIB_PATH), "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _abi.lib()


def test_every_entry_point_is_documented_for_integrators():
    """INTEGRATION.md maps each launcher of include/mednet_b200.h to the reference call site it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "mednet_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = set(re.findall(r"\b(mednet_[a-z0-9_]+)\s*\(", header))
    missing = sorted(n for n in names if not n.endswith("_workspace_bytes") and n not in doc)
    assert not missing, missing
