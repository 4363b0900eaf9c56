"""The C-ABI library loads without a GPU, exports every symbol the header declares, its struct
layouts agree with the header as gcc sees it, and compute entry points fail loudly (no CPU path)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cloud_renderer_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crn_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(pkg):
    lib = pkg.load_library()
    declared = header_functions()
    assert declared, "no declarations parsed"
    assert sorted(pkg.EXPORTS) == declared, "cloud_renderer_b200.EXPORTS out of sync with the header"
    nm = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (crn_[a-z0-9_]+)", nm))
    for name in declared:
        assert name in exported, f"{name} declared in the header but not exported"
        getattr(lib, name)


def test_struct_layouts_match_the_header(pkg):
    names = {"crn_volume_desc": pkg.VolumeDesc, "crn_sun": pkg.Sun, "crn_sun_derived": pkg.SunDerived, "crn_camera": pkg.Camera,
             "crn_trace_params": pkg.TraceParams, "crn_trace_stats": pkg.TraceStats, "crn_timings": pkg.Timings}
    prog = '#include <stdio.h>\n#include "cloud_renderer_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(src, "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        n, size = line.split()
        assert C.sizeof(names[n]) == int(size), f"{n}: ctypes {C.sizeof(names[n])} vs C {size}"


def test_version_string(pkg):
    assert b"sm_100a" in pkg.load_library().crn_version()


def test_no_cpu_fallback(pkg):
    """without a CUDA device the library refuses to create a context instead of computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CrnError) as e:
        pkg.Renderer(0)
    assert e.value.code in (pkg.CRN_ERR_NO_DEVICE, pkg.CRN_ERR_CUDA)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """nothing under cloud-renderer_b200/ may import, link or dlopen oracle/"""
    pkg_dir = os.path.join(ROOT, "cloud-renderer_b200")
    for dp, _, files in os.walk(pkg_dir):
        if "build" in dp.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text and "import_oracle" not in text, f"{f} references the oracle"
    ldd = subprocess.run(["ldd", os.path.join(pkg_dir, "libcloud_renderer_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_default_trace_params_are_the_reference_defaults(pkg):
    p = pkg.default_trace_params()       # src/Shaders/ConeTraceShader.hpp:15-36
    assert (p.numOctaves, p.minNoiseSteps, p.maxNoiseSteps, p.vctSteps) == (4, 2, 8, 16)
    assert np.allclose([p.stepSize, p.noiseOpacity, p.freqStep, p.persStep, p.adjustSize, p.minNoiseColor, p.noiseColorScale],
                       [0.01, 4.0, 3.0, 0.5, 40.0, 0.2, 0.45])
    assert np.allclose([p.vctConeAngle, p.vctConeInitialHeight, p.vctLodOffset, p.vctDownScaling], [0.9, 0.1, 0.0, 1.0])
    assert np.allclose(list(p.windVel), [0.01, 0, 0]) and np.allclose(list(p.clearColor), [0.2, 0.3, 0.5, 1.0])
    assert (p.showQuad, p.doConeTrace, p.doNoiseSample) == (0, 1, 1)
    assert p.transmittanceCutoff == 0.0 and p.sampler == pkg.SAMPLER_EXPLICIT
