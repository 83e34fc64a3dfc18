"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sdgpu.h declares, POD layouts match, and compute entry points fail loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from seekdeep_b200 import _abi
from seekdeep_b200.aligner import GpuAligner, SdgpuError
from seekdeep_b200.scoring import make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("sdgpu.h", "sdgpu_qluster.h", "sdgpu_stitch.h"):
        path = os.path.join(ROOT, "include", hdr)
        if not os.path.exists(path):
            continue
        txt = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names |= set(re.findall(r"\b(sdgpu_[a-z0-9_]+|sdq_[a-z0-9_]+|sdst_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_abi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 24
    for name in sorted(syms):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by libsdgpu.so"


def test_python_signatures_cover_header():
    hdr = {s for s in declared_symbols() if s.startswith("sdgpu_")}
    assert hdr == set(_abi.SIGNATURES), hdr ^ set(_abi.SIGNATURES)


def test_pod_layouts():
    assert C.sizeof(_abi.GapPars) == 40
    assert C.sizeof(_abi.Params) == 40 + 4 * 128 * 128 + 24
    assert C.sizeof(_abi.Comparison) == 104 and C.sizeof(_abi.Gap) == 12
    assert _abi.load().sdgpu_abi_version() == 3


def test_no_cpu_fallback_without_device():
    lib = _abi.load()
    if lib.sdgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(SdgpuError) as ei:
        GpuAligner(make_params())
    assert ei.value.code == _abi.E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_reference_oracle():
    """The shipped path must not include, link or import anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "seekdeep_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle|sdo_|libsdoracle", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_headers_compile_as_c99_and_cxx():
    """include/*.h are the drop-in boundary: plain C (no C++-isms) and usable from C++."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for hdr in ("sdgpu.h", "sdgpu_qluster.h", "sdgpu_stitch.h"):
        for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c"], ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++"]):
            r = subprocess.run(cmd + [os.path.join(inc, hdr)], capture_output=True, text=True)
            assert r.returncode == 0, f"{' '.join(cmd)} {hdr}:\n{r.stderr}"


def test_cxx_example_builds_against_the_library():
    import __graft_entry__ as g
    assert os.path.exists(g.build_example())


def test_python_struct_layouts_match_the_headers(tmp_path):
    """sizeof / offsetof of the PODs as the C compiler sees them vs the ctypes / numpy mirrors."""
    import subprocess
    import numpy as np
    import importlib
    Q = importlib.import_module("seekdeep_b200.qluster")  # (the package also exports a function of that name)
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "sdgpu_qluster.h"
#define S(t) printf(#t " %zu\n", sizeof(t))
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(sdgpu_params); S(sdgpu_gap); S(sdgpu_comparison); S(sdgpu_iter_par); S(sdgpu_pair_summary); S(sdgpu_chimera_probe);
  S(sdgpu_profile); S(sdq_opts); S(sdq_stats); S(sdgpu_grid_hit);
  O(sdgpu_comparison, oneBaseIndel); O(sdgpu_comparison, queryCoverage); O(sdgpu_iter_par, perId);
  O(sdq_opts, markChimeras); O(sdq_opts, chiParentFreqs); O(sdq_opts, chiOverlap);
  O(sdq_opts, kmerCutOff); O(sdq_opts, useKmerBinning); O(sdq_opts, kCompareLen); O(sdq_opts, binPars); O(sdq_opts, nBinPars);
  O(sdq_stats, kmerPairs); O(sdq_stats, kmerBins); O(sdq_stats, streamedGrids);
  O(sdq_stats, chimeraPairs); O(sdq_stats, secondsGpu); O(sdq_stats, msResident);
  O(sdgpu_chimera_probe, flags); O(sdgpu_profile, evaluatedCells); O(sdgpu_profile, bandKernelMs); O(sdgpu_profile, screenCertCells);
  O(sdgpu_grid_hit, mask); O(sdq_stats, passingPairs);
  return 0;
}
''')
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip().splitlines())
    c = {k: int(v) for k, v in out.items()}
    assert c["sdgpu_params"] == C.sizeof(_abi.Params) and c["sdgpu_gap"] == C.sizeof(_abi.Gap)
    assert c["sdgpu_comparison"] == C.sizeof(_abi.Comparison) == _abi.COMPARISON_DTYPE.itemsize
    assert c["sdgpu_comparison.oneBaseIndel"] == _abi.COMPARISON_DTYPE.fields["oneBaseIndel"][1]
    assert c["sdgpu_comparison.queryCoverage"] == _abi.COMPARISON_DTYPE.fields["queryCoverage"][1]
    assert c["sdgpu_iter_par"] == C.sizeof(Q.IterParC) and c["sdgpu_iter_par.perId"] == Q.IterParC.perId.offset
    assert c["sdgpu_pair_summary"] == _abi.PAIR_SUMMARY_DTYPE.itemsize
    assert c["sdgpu_chimera_probe"] == _abi.CHIMERA_PROBE_DTYPE.itemsize
    assert c["sdgpu_chimera_probe.flags"] == _abi.CHIMERA_PROBE_DTYPE.fields["flags"][1]
    assert c["sdq_opts"] == C.sizeof(Q.QlusterOptsC)
    for f in ("markChimeras", "chiParentFreqs", "chiOverlap", "kmerCutOff", "useKmerBinning", "kCompareLen", "binPars", "nBinPars"):
        assert c["sdq_opts." + f] == getattr(Q.QlusterOptsC, f).offset, f
    assert c["sdq_stats"] == C.sizeof(Q.StatsC)
    for f in ("chimeraPairs", "secondsGpu", "msResident", "kmerPairs", "kmerBins", "streamedGrids"):
        assert c["sdq_stats." + f] == getattr(Q.StatsC, f).offset, f
    # sdgpu_profile is mirrored inside GpuAligner.profile_read: 16 x 8 bytes
    assert c["sdgpu_profile"] == 128 and c["sdgpu_profile.evaluatedCells"] == 48 and c["sdgpu_profile.bandKernelMs"] == 72
    assert c["sdgpu_profile.screenCertCells"] == 112
    assert c["sdgpu_grid_hit"] == _abi.GRID_HIT_DTYPE.itemsize and c["sdgpu_grid_hit.mask"] == _abi.GRID_HIT_DTYPE.fields["mask"][1]
    assert c["sdq_stats.passingPairs"] == Q.StatsC.passingPairs.offset
    assert np.dtype(np.uint64).itemsize == 8
