"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly the entry
points declared in include/spt_b200.h (no compute calls — there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'spt_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(spt_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    from superpoint_transformer_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/spt_b200.h but not exported'
    # and the ctypes table binds every declared symbol, nothing else
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.spt_abi_version() == 1
    assert b'sm_100a' in lib.spt_build_info()


def test_library_contains_sm100a_sass_only():
    from superpoint_transformer_b200 import _lib
    _lib.load()
    out = subprocess.run(['cuobjdump', '--list-elf', _lib.library_path()],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_(\d+a?)', out.stdout))
    assert archs == {'100a'}, archs


def test_workspace_queries_are_host_only_and_monotone():
    from superpoint_transformer_b200 import _lib
    lib = _lib.load()
    a = lib.spt_group_index_workspace_bytes(1000, 100)
    b = lib.spt_group_index_workspace_bytes(1000, 100000)
    assert 0 < a < b
    assert lib.spt_graphnorm_workspace_bytes(2, 128) > 0
    assert lib.spt_unitsphere_workspace_bytes(10) >= 120


def test_ops_refuse_cpu_tensors_loudly():
    """no CPU fallback: the product path raises instead of computing on the host"""
    import torch
    import superpoint_transformer_b200 as S
    x = torch.randn(10, 8)
    idx = torch.randint(0, 3, (10,))
    with pytest.raises(RuntimeError, match='CUDA'):
        S.ops.segment_pool(x, idx, 3, 'max')
    with pytest.raises(RuntimeError, match='CUDA'):
        S.nn.GraphNorm(8)(x)
    with pytest.raises(RuntimeError, match='CUDA'):
        S.nn.SelfAttentionBlock(8, num_heads=2, qk_dim=2)(x, torch.randint(0, 10, (2, 30)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'superpoint_transformer_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the header's structs have the C compiler's size and field offsets
    (gcc on include/spt_b200.h: the header is plain C)."""
    import ctypes
    from superpoint_transformer_b200 import _lib
    structs = {'spt_select_level': _lib.SelectLevel, 'spt_attn_extras': _lib.AttnExtras}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'  printf("%zu\\n", sizeof({cname}));\n')
        lines += [f'  printf("%zu\\n", offsetof({cname}, {f}));\n' for f, _ in cls._fields_]
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include "spt_b200.h"\nint main(void) {\n' +
                   ''.join(lines) + '  printf("%d\\n", (int)SPT_SEL_ROWS);\n  return 0;\n}\n')
    exe = tmp_path / 'layout'
    cc = subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o',
                         str(exe)], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    out = [int(v) for v in subprocess.run([str(exe)], capture_output=True,
                                          text=True).stdout.split()]
    for cname, cls in structs.items():
        n = len(cls._fields_)
        assert out[0] == ctypes.sizeof(cls), cname
        assert out[1:n + 1] == [getattr(cls, f).offset for f, _ in cls._fields_], cname
        out = out[n + 1:]
    assert out == [_lib.SEL_ROWS]
