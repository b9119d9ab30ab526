"""The reference-side adapter (integration/search_gpu.cpp) as a compiled artefact: built against the reference's own
gps/gps.h, exporting the six entry points with the reference's prototypes, linked with libacq_b200.so."""
import os
import subprocess

import pytest

from flydog_sdr_gps_b200 import adapter

HAVE_REF = os.path.exists(os.path.join(adapter.REF, "gps", "gps.h"))


def test_adapter_builds_and_exports_the_reference_symbols():
    lib = adapter.build()
    if lib is None:
        pytest.skip("no reference tree and no prebuilt adapter")
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for proto, mangled in adapter.REFERENCE_SYMBOLS.items():
        assert (" T " + mangled + "\n") in syms, "adapter does not define %s" % proto
    dem = subprocess.run(["nm", "-D", "-C", "--defined-only", lib], capture_output=True, text=True).stdout
    for proto in adapter.REFERENCE_SYMBOLS:
        assert proto in dem, proto
    # it links the engine, not FFTW
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    assert "libacq_b200.so" in needed and "fftw" not in needed.lower()


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_adapter_prototypes_are_the_reference_headers():
    """The mangled names above encode the prototypes; gps.h is where they come from -- check the header still says so."""
    text = open(os.path.join(adapter.REF, "gps", "gps.h")).read()
    for decl in ("void SearchInit();", "void SearchFree();", "void SearchTask(void *param);", "void SearchTaskRun();",
                 "void SearchEnable(int sat);", "void SearchParams(int argc, char *argv[]);"):
        assert decl in text, decl


def test_adapter_fails_loudly_without_a_gpu():
    """SearchInit on a box without a usable GPU: the engine refuses (no CPU fallback) and the adapter exits through
    kiwi_exit, like the reference does on fatal configuration errors."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    if adapter.build() is None:
        pytest.skip("no reference tree and no prebuilt adapter")
    a = adapter.Adapter()
    assert a.init() == -1
