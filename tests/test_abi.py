"""The C-ABI library loads on a CPU-only box, exports every symbol include/mitransient_b200.h declares, has
struct layouts identical to the ctypes mirror, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mitransient_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tr_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from mitransient_b200 import _abi, _lib
    lib = _lib.load()
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_abi.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported by the library"
    assert lib.tr_abi_version() == _abi.TR_ABI_VERSION


def test_struct_layouts_match_c(tmp_path):
    """sizeof/offsetof of every ABI struct as the C compiler sees it == the ctypes mirror."""
    from mitransient_b200 import _abi
    structs = ["TrShapeDesc", "TrBsdfDesc", "TrEmitterDesc", "TrMediumDesc", "TrSensorDesc", "TrFilmDesc",
               "TrSceneDesc", "TrRenderParams", "TrCounters", "TrDeposit"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s in structs:
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for fname, _ in getattr(_abi, s)._fields_:
            lines.append(f'printf("{s}.{fname} %zu\\n", offsetof({s}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout
    for line in out.strip().splitlines():
        key, val = line.split()
        if "." in key:
            s, f = key.split(".")
            assert getattr(getattr(_abi, s), f).offset == int(val), key
        else:
            assert C.sizeof(getattr(_abi, key)) == int(val), key


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mitransient_b200 as mitr
    from mitransient_b200 import _abi, _lib
    lib = _lib.load()
    assert lib.tr_device_check() == _abi.TR_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.tr_last_error()
    scene = mitr.load_dict(mitr.cornell_box_config(8, 8, 16, 1))
    with pytest.raises(mitr.TransientError):
        mitr.render(scene, spp=1)
    h = C.c_void_p()
    desc = scene.describe(scene.sensors()[0])
    assert lib.tr_scene_create(C.byref(desc), C.byref(h)) == _abi.TR_ERR_NO_DEVICE and not h.value


def test_argument_validation_without_gpu():
    """NULL / malformed arguments are rejected before any device work."""
    from mitransient_b200 import _abi, _lib
    lib = _lib.load()
    assert lib.tr_film_develop(None, None, 16, 4, None) == _abi.TR_ERR_INVALID_ARGUMENT
    fd = _abi.TrFilmDesc()
    assert lib.tr_film_scatter_add(None, C.byref(fd), 4, None, 0, None) == _abi.TR_ERR_INVALID_ARGUMENT
    assert lib.tr_render(None, None, None, None, None, None) == _abi.TR_ERR_INVALID_ARGUMENT
    assert lib.tr_scene_info(None, None, None, None) == _abi.TR_ERR_INVALID_ARGUMENT
    lib.tr_scene_destroy(None)      # must be a no-op


def test_product_never_imports_oracle():
    """Nothing under mitransient_b200/ may import, link or call the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "mitransient_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "transient_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
    so = os.path.join(pkg, "lib", "libmitransient_b200.so")
    if os.path.exists(so):
        syms = subprocess.run(["nm", "-D", so], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle_" not in syms
