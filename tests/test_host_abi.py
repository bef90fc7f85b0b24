"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/rxmd_b200.h declares, the ctypes
mirrors have the C layout, and every entry point fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from rxmd_b200.host import binding, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rxmd_b200.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rxg_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    lib = engine.load_library()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rxmd_b200.h but not exported"


def test_struct_layouts_match_c(built):
    fields = {"rxg_config": [f[0] for f in binding.RxgConfig._fields_],
              "rxg_box": [f[0] for f in binding.RxgBox._fields_],
              "rxg_ff": [f[0] for f in binding.RxgFF._fields_]}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for st, fl in fields.items():
        src.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fl:
            src.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    src.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "l.c")
        open(c, "w").write("\n".join(src))
        subprocess.check_call(["/usr/bin/gcc", c, "-o", os.path.join(d, "l")])
        out = subprocess.check_output([os.path.join(d, "l")], text=True)
    got = dict(line.split() for line in out.strip().splitlines())
    for st, cls in (("rxg_config", binding.RxgConfig), ("rxg_box", binding.RxgBox), ("rxg_ff", binding.RxgFF)):
        assert int(got[st]) == C.sizeof(cls)
        for name, _ in cls._fields_:
            assert int(got[f"{st}.{name}"]) == getattr(cls, name).offset, (st, name)


def test_no_cpu_fallback(built, rdx_paths):
    """Without a usable sm_100 device rxg_create must fail with a message; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU box")
    from rxmd_b200.host.system import build_system
    s = build_system(*rdx_paths)
    with pytest.raises(engine.RxmdError) as ei:
        engine.Engine(s, s.config())
    assert "CUDA" in str(ei.value) or "device" in str(ei.value)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: no file of the product package may reference it."""
    pkg = os.path.join(ROOT, "rxmd_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "rxmd_oracle" not in txt and "librxmd_oracle" not in txt, f


def test_hint_flags_agree_across_header_python_and_shim():
    """rxg_hint's flags are promises a shim makes (include/rxmd_b200.h): the ctypes mirror and the Fortran shim must carry the
    header's values, and every flag the header defines must be known to both."""
    hdr = dict(re.findall(r"#define\s+(RXG_HINT_[A-Z_]+)\s+(\d+)", open(HEADER).read()))
    assert len(hdr) >= 4 and len(set(hdr.values())) == len(hdr)
    for name, val in hdr.items():
        assert int(val) & (int(val) - 1) == 0, f"{name} is not a single bit"
        assert getattr(engine, name.replace("RXG_", "")) == int(val), name
    shim = open(os.path.join(ROOT, "rxmd_b200", "gpu_shim.F90")).read()
    for name, val in hdr.items():
        assert re.search(name + r"\s*=\s*" + val + r"\b", shim), f"{name} = {val} missing from gpu_shim.F90"
