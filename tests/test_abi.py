"""CPU: the C-ABI library loads and exports every symbol include/pepflow_b200.h declares; argument
validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT, golden_state_dict_spec


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pepflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from pepflowww_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_lib.SIGNATURES) == names


def test_version_strerror_and_config_check():
    from pepflowww_b200 import _lib
    lib = _lib.load()
    assert lib.pf_version() == 4
    assert lib.pf_strerror(0) == b"ok"
    assert b"shape" in lib.pf_strerror(-1)
    assert lib.pf_check_config(128, 64, 128, 8, 8, 12, 4, 2) == 0
    assert lib.pf_check_config(256, 64, 128, 8, 8, 12, 4, 2) == -2
    assert lib.pf_set_option(b"no_such_option", 1) == -7
    assert lib.pf_set_option(b"edge_impl", 7) == -7
    assert lib.pf_get_option(b"edge_impl") in (0, 1, 2)


def test_null_and_shape_errors_without_gpu():
    from pepflowww_b200 import _lib
    lib = _lib.load()
    assert lib.pf_linear(None, None, None, None, None, None, 4, 4, 4, 0, None) == -3
    assert lib.pf_so3_log(None, None, 4, None) == -3
    assert lib.pf_ga_encoder_workspace_bytes(2, 24) > 2 * 24 * 24 * 64 * 4
    assert lib.pf_edge_transition_workspace_bytes(2, 24) > 0


def test_enum_sizes_match_header():
    from pepflowww_b200 import _lib
    text = open(os.path.join(ROOT, "include", "pepflow_b200.h")).read()
    g = re.search(r"enum pf_ga_global_slot \{(.*?)\};", text, re.S).group(1)
    b = re.search(r"enum pf_ga_block_slot \{(.*?)\};", text, re.S).group(1)
    strip = lambda s: [x.split("=")[0].strip() for x in re.sub(r"/\*.*?\*/", "", s, flags=re.S).split(",") if x.strip()]
    gs, bs = strip(g), strip(b)
    assert gs[-1] == "PF_G_NSLOTS" and bs[-1] == "PF_B_NSLOTS"
    assert [x[len("PF_G_"):] for x in gs[:-1]] == _lib.G_SLOTS
    assert [x[len("PF_B_"):] for x in bs[:-1]] == _lib.B_SLOTS


def test_state_dict_surface_matches_reference():
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    cfg, name = load_config()
    assert name == "learn_angle"
    sd = FlowModel(cfg.model).state_dict()
    spec = golden_state_dict_spec()
    assert set(sd) == set(spec)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == shape, k
    assert sum(v.numel() for k, v in sd.items() if not k.endswith("freq_bands")) == 6880353


def test_no_cpu_fallback():
    import torch
    from pepflowww_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.zeros(4, 8), torch.zeros(8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.so3_log(torch.eye(3)[None])


def test_plain_c_consumer(tmp_path):
    """include/pepflow_b200.h compiles as C99 with -Wall -Werror and a dlopen consumer (tests/c/abi_smoke.c) reaches the
    entry points with plain pointers and sizes - the binding surface a cgo / JNI / ctypes stub sees.  No GPU needed."""
    import shutil
    import subprocess
    from pepflowww_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c", "abi_smoke.c"), "-ldl", "-o", exe], check=True)
    r = subprocess.run([exe, _lib.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi ok" in r.stdout


def test_struct_layouts_match_the_header(tmp_path):
    """pf_sampler / pf_ga_weights as C sees them (gcc, offsetof / sizeof) == the ctypes Structures the Python mirror passes:
    a field added on one side only would shift every pointer after it."""
    import shutil
    import subprocess
    from pepflowww_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    fields = [n for n, _ in _lib.Sampler._fields_]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "pepflow_b200.h"', 'int main(void) {',
            '  printf("sizeof_sampler %zu\\n", sizeof(pf_sampler));',
            '  printf("sizeof_weights %zu\\n", sizeof(pf_ga_weights));',
            '  printf("w_g %zu\\n", offsetof(pf_ga_weights, g));', '  printf("w_blk %zu\\n", offsetof(pf_ga_weights, blk));',
            '  printf("w_prepacked %zu\\n", offsetof(pf_ga_weights, prepacked));',
            '  printf("w_prepacked_bytes %zu\\n", offsetof(pf_ga_weights, prepacked_bytes));']
    prog += [f'  printf("{f} %zu\\n", offsetof(pf_sampler, {f}));' for f in fields]
    prog += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = str(tmp_path / "layout")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    out = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof_sampler"]) == ctypes.sizeof(_lib.Sampler)
    for f in fields:
        assert int(out[f]) == getattr(_lib.Sampler, f).offset, f
    assert int(out["sizeof_weights"]) == ctypes.sizeof(_lib.GaWeights)
    for c_name, py_name in (("w_g", "g"), ("w_blk", "blk"), ("w_prepacked", "prepacked"), ("w_prepacked_bytes", "prepacked_bytes")):
        assert int(out[c_name]) == getattr(_lib.GaWeights, py_name).offset, py_name
