"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol the header declares.
No compute call is made here (there is no GPU in the development container and no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from flydog_sdr_gps_b200 import _build, _lib, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "acq_b200.h")


DROPIN_HEADER = os.path.join(ROOT, "include", "search_dropin.h")


def declared_functions(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acq_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = _build.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    names = declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), "libacq_b200.so does not export %s" % n
    # the ctypes binding covers the same set
    assert sorted(_lib.exported_symbols()) == names
    # the host shim header (drop-in scheduling + capture sources) is exported by the same library
    shim = declared_functions(DROPIN_HEADER)
    assert len(shim) >= 14
    for n in shim:
        assert hasattr(L, n), "libacq_b200.so does not export %s" % n


def test_struct_layouts_match_header():
    assert engine.RECORD_DTYPE.itemsize == 24  # acq_record
    assert engine.CELL_DTYPE.itemsize == 16    # acq_cell
    assert engine.FINE_DTYPE.itemsize == 16    # acq_fine
    assert C.sizeof(_lib.AcqSat) == 16
    assert C.sizeof(_lib.AcqParams) == 64  # ABI version 3: struct_size first, reserved[6]


def test_defaults_are_the_reference_constants():
    p = engine.default_params()
    assert (p.dop_lo, p.dop_hi, p.half_bin, p.k_noncoh) == (-20, 20, 0, 1)  # gps/search.cpp:465
    assert p.thr_l1 == 16.0 and p.thr_e1b == 16.0                            # gps/gps.h:60, search.cpp:549
    assert p.wrap_mode == engine.WRAP_REFERENCE
    assert p.sample_bits == 1                                                # I_sign only, search.cpp:408-411
    assert p.code_doppler == 0 and list(p.reserved) == [0] * 6 and p.struct_size == 64
    assert _lib.load().acq_abi_version() == 3


def test_no_cpu_fallback():
    """Without a CUDA device engine creation must fail loudly, never compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from flydog_sdr_gps_b200 import sats
    with pytest.raises(engine.AcqError) as ei:
        engine.AcqEngine(sats.navstar())
    assert ei.value.code == -3  # ACQ_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_argument_errors_without_gpu():
    L = _lib.load()
    assert L.acq_params_default(None) == -1
    h = C.c_void_p()
    assert L.acq_create(C.byref(h), None, None, 0, 0) == -1
    assert b"satellite table" in L.acq_last_error()
    assert L.acq_destroy(None) == 0
    # parameter validation precedes the device probe: an unknown capture format is an argument error everywhere
    sat = _lib.AcqSat(1, 2, 6, 0)
    assert L.acq_create(C.byref(h), C.byref(engine.default_params(sample_bits=3)), C.byref(sat), 1, 0) == -1
    assert b"sample_bits" in L.acq_last_error()
    assert L.acq_create(C.byref(h), C.byref(engine.default_params(code_doppler=2)), C.byref(sat), 1, 0) == -1
    # code-Doppler compensation is refused (not silently truncated) when it would need too many shifted spectra
    big = engine.default_params(code_doppler=1, k_noncoh=255, dop_lo=-2048, dop_hi=2048)
    assert L.acq_create(C.byref(h), C.byref(big), C.byref(sat), 1, 0) == -4  # ACQ_ERR_UNSUPPORTED
    assert b"shifted copies" in L.acq_last_error()


def test_params_struct_size_guard():
    """A caller built against another ABI version (different sizeof(acq_params)) is refused before the library reads
    past the end of its structure (v1 callers passed 32 bytes, v2 48)."""
    L = _lib.load()
    h = C.c_void_p()
    sat = _lib.AcqSat(1, 2, 6, 0)
    for bad in (0, 32, 48, 128):
        p = engine.default_params()
        p.struct_size = bad
        assert L.acq_create(C.byref(h), C.byref(p), C.byref(sat), 1, 0) == -1
        assert b"struct_size" in L.acq_last_error()
    p = engine.default_params()
    p.reserved[5] = 1
    assert L.acq_create(C.byref(h), C.byref(p), C.byref(sat), 1, 0) == -1
    assert b"reserved" in L.acq_last_error()


def test_product_library_reads_no_environment_and_holds_no_experiment_kernels():
    """The drop-in library's behaviour must not depend on the host process's environment: A/B kernel
    forms and launch-policy switches exist only in the variant libraries tools/build_variants.py builds."""
    import subprocess
    csrc = os.path.join(ROOT, "flydog_sdr_gps_b200", "csrc")
    for f in ("acq_api.cu", "acq_kernels.cu", "acq_fft.cuh", "search_dropin.cpp", "acq_microbench.cu"):
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    # (the statically linked CUDA runtime imports getenv for its own CUDA_* variables: only the sources can be checked)
    for k in ("k_search_l1_ldg", "k_search_l1_x3", "k_search_e1b_ldg", "k_search_l1_sp", "k_search_l1_st"):
        assert k not in syms, k
    sass = subprocess.run(["cuobjdump", "-elf", _lib.lib_path()], capture_output=True, text=True).stdout
    for k in ("k_search_l1_ldg", "k_search_l1_x3", "k_search_e1b_ldg", "k_search_l1_sp", "k_search_l1_st"):
        assert k not in sass, k
    assert "k_search_l1" in sass and "k_search_e1b" in sass and "k_pick_small" in sass


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "flydog_sdr_gps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in text and "acq_oracle" not in text and "orc_fft" not in text, f


def test_capture_sources(tmp_path):
    """SPI packet assembly and the raw capture file reader (gps/search.cpp:361-406): host-only code."""
    import numpy as np
    from flydog_sdr_gps_b200 import capture
    rng = np.random.default_rng(3)
    caps = rng.integers(0, 256, (3, 8192), dtype=np.uint8)
    # 16 x 512-byte packets in arrival order == the capture
    assert np.array_equal(capture.from_packets([caps[0][k * 512:(k + 1) * 512] for k in range(16)]), caps[0])
    with pytest.raises(ValueError):
        capture.from_packets([caps[0][:512]] * 15)
    path = tmp_path / "if.4092.dat"
    path.write_bytes(caps.tobytes() + b"\x55" * 100)  # a trailing partial capture is never returned
    with capture.CaptureFile(path) as f:
        assert f.remaining() == 3
        assert np.array_equal(f.next(), caps[0])
        assert np.array_equal(f.next(2), caps[1:].reshape(-1))
        assert f.next() is None and f.remaining() == 0
        f.rewind()
        assert f.next(4) is None  # non-coherent capture of 4 blocks: not enough data, nothing consumed
        assert np.array_equal(f.next(3), caps.reshape(-1))
    with pytest.raises(OSError):
        capture.CaptureFile(tmp_path / "missing.dat")


def test_bench_reference_arm_contract():
    """bench.py --impl reference runs on the host cores alone (no GPU) and prints the one-line JSON contract."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cells/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("acquisition cells/sec") and line["config"]["workload"] == "cfg5"
    assert line["config"]["captures_total"] == 1024 and line["scaling"] == "strong"
    # both arms print the same `config` object (the workload only)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.config_dict("cfg5", 1)
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_workload_descriptions():
    """bench.config_dict: the workload sizes of SURVEY 8(d) / BASELINE.md section 3, identical for both arms."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    want = {"cfg1": (5_368_704, 1_312), "cfg2": (21_081_984, 103_040), "cfg3": (66_290_400, 4_050), "cfg4": (38_923_104, 3_362),
            "cfg5": (5_497_552_896, 1_343_488)}
    for cfg, (cells, tiles) in want.items():
        d = bench.config_dict(cfg, 1)
        assert (d["cells_per_step"], d["tiles_per_step"]) == (cells, tiles), cfg
        assert d["workload"] == cfg and d["l2"] == bench.L2_NOTE
    assert bench.config_dict("cfg5", 8)["captures_total"] == 1024
    # the farm's captures are defined by their index alone (a rank generating only its shard gets the same bytes)
    assert bench.farm_signals(5) == bench.farm_signals(5) and bench.farm_signals(5) != bench.farm_signals(6)
    a = bench.farm_captures_numpy([3])[0]
    assert a.shape == (8192,) and np.array_equal(a, bench.farm_captures_numpy([2, 3])[1])


def test_launch_plan_covers_every_tile_once():
    """acq_plan_launch (host arithmetic of launch_search, no device): which kernel, grid, claimed or strided, and the chunk
    schedule of k_search_l1_cr.  The schedule must tile [0, n_tiles) exactly -- a hole or an overlap would be a silently
    wrong or missing cell -- with single tiles over the last two rounds, and a launch claims from six rounds up."""
    L = _lib.load()
    plan = _lib.AcqLaunchPlan()

    def ask(k, half, e1b, n_tiles, sms=148):
        assert L.acq_plan_launch(k, half, e1b, n_tiles, sms, C.byref(plan)) == 0, L.acq_last_error()
        return {f: getattr(plan, f) for f, _ in plan._fields_}

    # the BASELINE shapes on a B200
    assert ask(1, 0, 0, 1024 * 32 * 41)["kernel"] == 6 and plan.grid == 296 and plan.claims == 1   # cfg5: k_search_l1_cr, chunks claimed
    assert ask(1, 0, 0, 32 * 41) == dict(kernel=6, grid=296, claims=0, chunk_big=16, chunk_mid=4, n_big=0, n_mid=0, n_chunks=1312)  # cfg1
    assert ask(1, 0, 0, 41)["grid"] == 41 and plan.claims == 0                                      # one satellite
    assert ask(20, 1, 0, 32 * 161)["kernel"] == 3 and plan.grid == 296 and plan.claims == 1         # cfg2: k_search_l1_multi
    assert ask(1, 1, 0, 32 * 161)["kernel"] == 0                                                    # half-bin K = 1: k_search_l1
    assert ask(1, 0, 1, 50 * 81)["kernel"] == 1 and plan.grid == 296 and plan.claims == 1           # cfg3: k_search_e1b
    assert ask(1, 0, 1, 30)["kernel"] == 2 and plan.grid == 30 and plan.claims == 0                 # a cluster per tile
    assert L.acq_plan_launch(1, 0, 0, 0, 148, C.byref(plan)) != 0 and L.acq_plan_launch(1, 0, 0, 2 ** 31, 148, C.byref(plan)) != 0

    rng = np.random.default_rng(5)
    sizes = [1, 2, 295, 296, 297, 1775, 1776, 1777, 2368, 4143, 4144, 4160, 5248, 167936, 2 ** 31 - 1]
    sizes += [int(x) for x in rng.integers(1, 3_000_000, 300)]
    for sms in (148, 132, 7):
        for n in sizes:
            d = ask(1, 0, 0, n, sms)
            g2 = 2 * sms
            big, mid = d["chunk_big"], d["chunk_mid"]
            singles = n - big * d["n_big"] - mid * d["n_mid"]
            assert singles >= 0 and d["n_chunks"] == d["n_big"] + d["n_mid"] + singles, (n, sms, d)
            assert d["grid"] == min(d["n_chunks"], g2), (n, sms, d)
            assert d["claims"] == (1 if n >= 6 * d["grid"] else 0), (n, sms, d)
            if not d["claims"]:
                assert d["n_big"] == 0 and d["n_mid"] == 0, (n, sms, d)     # static stride: single tiles
            else:
                assert singles >= 2 * g2 and singles < 2 * g2 + mid, (n, sms, d)        # the last two rounds go out one by one
                assert mid * d["n_mid"] < 2 * mid * g2 + big, (n, sms, d)               # about two rounds' worth in middle chunks
            # chunk c -> [start, start + len): consecutive, gap-free (the device's chunk_of, restated)
            starts = [0, big * d["n_big"], big * d["n_big"] + mid * d["n_mid"], n]
            assert starts == sorted(starts), (n, sms, d)
