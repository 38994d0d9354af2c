"""CPU tier: liblq.so loads, exports exactly what include/lq.h declares, and its host-only
entry points (version, workspace arithmetic, argument checks) behave without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lq.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lq_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def pkg():
    import learned_quantization_b200 as p
    return p


def test_header_and_library_agree(pkg):
    names = declared_functions()
    assert len(names) >= 20
    lib = C.CDLL(pkg._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/lq.h but not exported by liblq.so"
    # the ctypes table binds every declaration, and nothing that is not declared
    assert sorted(pkg._lib.SIGNATURES) == names


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code and "at::" not in code and "Tensor" not in code and "#include <cuda" not in code
    assert 'extern "C"' in code


def test_version_and_workspace_arithmetic(pkg):
    lib, L = pkg._lib.lib, pkg._lib
    assert lib.lq_version() == 200
    descs = (L.TensorDesc * 3)()
    for d, (o, c, i) in zip(descs, [(256, 512, 512), (784, 128, 1), (0, 4, 4)]):
        d.outer, d.C, d.inner = o, c, i
    out = C.c_size_t()
    sizes = {}
    for phase in (L.LQ_PHASE_FWD, L.LQ_PHASE_BWD, L.LQ_PHASE_PENALTY_FWD, L.LQ_PHASE_PENALTY_BWD):
        assert lib.lq_workspace_bytes(descs, 3, phase, C.byref(out)) == 0
        sizes[phase] = out.value
    assert sizes[L.LQ_PHASE_FWD] == 0
    # ROW regime: one float4 partial per (row, segment) = 256*512 units; COL: nchunks*W
    assert sizes[L.LQ_PHASE_BWD] >= 256 * 512 * 16 and sizes[L.LQ_PHASE_BWD] % 256 == 0
    assert sizes[L.LQ_PHASE_PENALTY_FWD] == sizes[L.LQ_PHASE_PENALTY_BWD] > sizes[L.LQ_PHASE_BWD]
    assert lib.lq_host_workspace_bytes(256, 512, 512, C.byref(out)) == 0 and out.value > 3 * (16 << 20)
    assert lib.lq_host_workspace_bytes(0, 5, 5, C.byref(out)) == 0 and out.value == 0


def test_argument_errors_are_status_codes(pkg):
    lib, L = pkg._lib.lib, pkg._lib
    out = C.c_size_t()
    assert lib.lq_workspace_bytes(None, 1, 0, C.byref(out)) == -1          # LQ_E_NULLPTR
    assert b"null" in lib.lq_last_error()
    descs = (L.TensorDesc * 1)()
    descs[0].outer, descs[0].C, descs[0].inner = -1, 2, 2
    assert lib.lq_workspace_bytes(descs, 1, 1, C.byref(out)) == -2         # LQ_E_BADSHAPE
    descs[0].outer = 1
    assert lib.lq_workspace_bytes(descs, 1, 9, C.byref(out)) == -3         # LQ_E_BADMODE
    assert lib.lq_fq_fwd(None, None, None, None, 0, 2, 2, 2, 0, float("-inf"), float("inf"), None) == -1
    assert lib.lq_fq_fwd(None, None, None, None, 0, 2, -2, 2, 0, float("-inf"), float("inf"), None) == -2
    with pytest.raises(L.LqError):
        L.check(-4)


def test_product_has_no_cpu_path(pkg):
    """CPU tensors are refused loudly; the product never imports the oracle."""
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.ops.fq_fwd(torch.zeros(4, 4), torch.ones(1))
    pkg_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                for line in open(os.path.join(dirpath, f)):
                    code = line.split("#")[0].split("//")[0]
                    assert not re.search(r"\b(import|from|include|dlopen|CDLL)\b.*oracle", code), (f, line)


def test_header_is_plain_c_and_the_c_demo_links(pkg, tmp_path):
    """include/lq.h compiles as C (not just C++) and a C program links against liblq.so without Python."""
    import shutil
    import subprocess
    if not (shutil.which("gcc") and os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA headers not available")
    exe = tmp_path / "lq_demo"
    cmd = ["gcc", "-O1", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "lq_demo.c"), "-o", str(exe),
           "-L" + os.path.dirname(pkg._lib.LIB_PATH), "-llq", "-L/usr/local/cuda/lib64", "-lcudart", "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
