"""ctypes bindings for the CPU oracle (oracle/_build/liboracle.so) and, when it has been built,
for the unmodified reference search (oracle/_ref/libref_search.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_search.so")
REF50_SO = os.path.join(HERE, "_ref", "libref_search_e1b50.so")  # same search.cpp over a 50-row Galileo table
REF_OFAST_SO = os.path.join(HERE, "_ref", "libref_search_ofast.so")  # same sources at the reference's own -Ofast (+AVX2/FMA)

N = 16384
BLOCK_BYTES = 8192
NAVSTAR, SBAS, QZSS, E1B = 0, 1, 2, 3
WRAP_REFERENCE, WRAP_CIRCULAR = 0, 1


class OrcSat(C.Structure):
    _fields_ = [("prn", C.c_int32), ("t1", C.c_int32), ("t2", C.c_int32), ("type", C.c_int32)]


class OrcParams(C.Structure):
    _fields_ = [("dop_lo", C.c_int32), ("dop_hi", C.c_int32), ("half_bin", C.c_int32),
                ("k_noncoh", C.c_int32), ("thr_l1", C.c_float), ("thr_e1b", C.c_float),
                ("wrap_mode", C.c_int32), ("sample_bits", C.c_int32), ("code_doppler", C.c_int32)]


class OrcSignal(C.Structure):
    _fields_ = [("sat", C.c_int32), ("tau", C.c_int32), ("doppler_hz", C.c_double),
                ("cn0_dbhz", C.c_double), ("phase", C.c_double), ("flip_ms", C.c_int32), ("code_doppler", C.c_int32)]


RECORD_DTYPE = np.dtype([("sat", "<i4"), ("lag", "<i4"), ("dop", "<i4"),
                         ("peak", "<f4"), ("noise", "<f4"), ("snr", "<f4")])
FINE_DTYPE = np.dtype([("dop_hz", "<f4"), ("code_fs", "<f4"), ("peak", "<f4"), ("ca_shift", "<i4")])
CELL_DTYPE = np.dtype([("peak", "<f4"), ("noise", "<f4"), ("snr", "<f4"), ("lag", "<i4")])
EVENT_DTYPE = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("d", "<i4"),
                        ("e", "<i4"), ("x", "<f8"), ("y", "<f8")])

_lib = None
_ref = None


def build(ref=True):
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-C", HERE, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)
    if ref and os.path.exists("/root/reference/gps/search.cpp"):
        subprocess.run(["make", "-C", HERE, "--no-print-directory", "ref"], check=True,
                       stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        fp = C.POINTER(C.c_float)
        u8 = C.POINTER(C.c_uint8)
        L.orc_params_default.argtypes = [C.POINTER(OrcParams)]
        L.orc_ca_chips.argtypes = [C.c_int, C.c_int, u8]
        L.orc_e1b_chips.argtypes = [C.c_int, u8]
        L.orc_hb_decimate.argtypes = [C.c_int, fp]
        L.orc_hb_decimate.restype = C.c_int
        L.orc_code_baseband.argtypes = [C.POINTER(OrcSat), fp]
        L.orc_code_spectrum.argtypes = [C.POINTER(OrcSat), fp]
        L.orc_capture_baseband.argtypes = [u8, C.c_int, fp]
        L.orc_capture_baseband_sm.argtypes = [u8, C.c_int, C.c_int, fp]
        L.orc_capture_spectrum.argtypes = [u8, C.c_int, fp]
        L.orc_fft16384.argtypes = [fp, C.c_int]
        L.orc_search_pre.argtypes = [u8, C.POINTER(OrcSat), C.c_int, fp, C.POINTER(C.c_int32), C.c_int,
                                     C.POINTER(OrcParams), C.c_void_p, C.c_void_p, C.c_int]
        L.orc_search_pre.restype = C.c_int
        L.orc_refine.argtypes = [u8, C.POINTER(OrcSat), C.c_int, C.POINTER(OrcParams), C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_int]
        L.orc_refine.restype = C.c_int
        L.orc_gen_capture.argtypes = [C.c_uint64, C.c_int, C.POINTER(OrcSat), C.c_int,
                                      C.POINTER(OrcSignal), C.c_int, u8]
        L.orc_gen_capture.restype = C.c_int
        L.orc_gen_capture_sm.argtypes = [C.c_uint64, C.c_int, C.POINTER(OrcSat), C.c_int,
                                         C.POINTER(OrcSignal), C.c_int, C.c_int, C.c_double, u8]
        L.orc_gen_capture_sm.restype = C.c_int
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(REF_SO)


def _bind_ref(path):
    R = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    u8 = C.POINTER(C.c_uint8)
    i32 = C.POINTER(C.c_int32)
    R.ref_init.restype = C.c_int
    R.ref_n_sats.restype = C.c_int
    R.ref_sat.argtypes = [C.c_int, i32, i32, i32, i32]
    R.ref_code_spectrum.argtypes = [C.c_int, fp]
    R.ref_code_baseband.argtypes = [C.c_int, fp]
    R.ref_code_baseband.restype = C.c_int
    R.ref_sample.argtypes = [u8, fp, fp]
    R.ref_correlate.argtypes = [C.c_int, i32, i32]
    R.ref_correlate.restype = C.c_float
    R.ref_search.argtypes = [u8, i32, C.c_int, i32, i32, fp]
    R.ref_search_task.argtypes = [u8, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    R.ref_search_task.restype = C.c_int
    R.ref_get_events.argtypes = [C.c_void_p, C.c_int]
    R.ref_e1b_chips.argtypes = [C.c_int, u8]
    R.ref_ca_chips.argtypes = [C.c_int, C.c_int, u8]
    R.ref_init()
    return R


def ref():
    """The unmodified reference search.cpp behind stubs (None if oracle/_ref was not built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            return None
        _ref = _bind_ref(REF_SO)
    return _ref


_ref_bench = None
REF_BENCH_BUILD = None


def _cpu_has(*flags):
    try:
        words = set(open("/proc/cpuinfo").read().split())
    except OSError:
        return False
    return all(f in words for f in flags)


def ref_bench():
    """The unmodified reference for TIMING (bench.py's CPU arms): the -Ofast + AVX2/FMA build where it exists and the host
    CPU can run it (the reference's own build compiles gps/ with -Ofast), else the strict build.  Never used for parity.
    REF_BENCH_BUILD says which."""
    global _ref_bench, REF_BENCH_BUILD
    if _ref_bench is None:
        if os.path.exists(REF_OFAST_SO) and _cpu_has("avx2", "fma"):
            _ref_bench = _bind_ref(REF_OFAST_SO)
            REF_BENCH_BUILD = "-Ofast -mavx2 -mfma (gps/ is built with -Ofast in the reference, Makefile.comp.inc)"
        else:
            _ref_bench = ref()
            REF_BENCH_BUILD = "-O2 strict IEEE (the parity build; no -Ofast/AVX2 build usable on this host)"
    return _ref_bench


_ref50 = None


def ref50():
    """The unmodified search.cpp over a table of all 50 Galileo E1-B codes (None if not built).  Same API as ref();
    sat index = PRN - 1."""
    global _ref50
    if _ref50 is None:
        if not os.path.exists(REF50_SO):
            return None
        _ref50 = _bind_ref(REF50_SO)
    return _ref50


def ref_e1b_chips(prn):
    """4092 chips of E1-B PRN prn from the reference's E1BCODE (gps/e1bcode.h:63-92)."""
    out = np.zeros(4092, np.uint8)
    ref().ref_e1b_chips(int(prn), _u8(out))
    return out


def ref_ca_chips(t1, t2):
    """1023 chips from the reference's CACODE (gps/cacode.h:23-64)."""
    out = np.zeros(1023, np.uint8)
    ref().ref_ca_chips(int(t1), int(t2), _u8(out))
    return out


def ref50_search(packed, sel):
    """Sample()+Correlate() of the unmodified reference over the 50-row Galileo table (sat = PRN - 1)."""
    packed = np.ascontiguousarray(packed, np.uint8)
    sel = np.ascontiguousarray(sel, np.int32)
    dop, lag, snr = np.zeros(len(sel), np.int32), np.zeros(len(sel), np.int32), np.zeros(len(sel), np.float32)
    i32 = C.POINTER(C.c_int32)
    ref50().ref_search(_u8(packed), sel.ctypes.data_as(i32), len(sel), dop.ctypes.data_as(i32), lag.ctypes.data_as(i32), _fp(snr))
    return dop, lag, snr


def ref50_code_spectrum(sat):
    out = np.zeros(2 * N, np.float32)
    ref50().ref_code_spectrum(sat, _fp(out))
    return out.view(np.complex64)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def sat_array(sats):
    """sats: iterable of (prn, t1, t2, type) -> ctypes array."""
    arr = (OrcSat * len(sats))()
    for i, s in enumerate(sats):
        arr[i].prn, arr[i].t1, arr[i].t2, arr[i].type = [int(v) for v in s]
    return arr


def default_params(**kw):
    p = OrcParams()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def ca_chips(t1, t2):
    out = np.zeros(1023, np.uint8)
    lib().orc_ca_chips(t1, t2, _u8(out))
    return out


def e1b_chips(prn):
    out = np.zeros(4092, np.uint8)
    lib().orc_e1b_chips(prn, _u8(out))
    return out


def code_baseband(sat):
    out = np.zeros(2 * N, np.float32)
    s = OrcSat(*[int(v) for v in sat])
    lib().orc_code_baseband(C.byref(s), _fp(out))
    return out.view(np.complex64)


def code_spectrum(sat):
    out = np.zeros(2 * N, np.float32)
    s = OrcSat(*[int(v) for v in sat])
    lib().orc_code_spectrum(C.byref(s), _fp(out))
    return out.view(np.complex64)


def block_bytes(sample_bits):
    """Bytes of one 65536-sample block: 8192 (sign only) or 16384 (sign plane + magnitude plane)."""
    return 2 * BLOCK_BYTES if sample_bits == 2 else BLOCK_BYTES


def capture_baseband(packed, half_rot=0, sample_bits=1):
    packed = np.ascontiguousarray(packed, np.uint8)
    assert packed.size == block_bytes(sample_bits)
    out = np.zeros(2 * N, np.float32)
    lib().orc_capture_baseband_sm(_u8(packed), sample_bits, half_rot, _fp(out))
    return out.view(np.complex64)


def capture_spectrum(packed, half_rot=0):
    packed = np.ascontiguousarray(packed, np.uint8)
    out = np.zeros(2 * N, np.float32)
    lib().orc_capture_spectrum(_u8(packed), half_rot, _fp(out))
    return out.view(np.complex64)


def fft16384(x, sign):
    buf = np.ascontiguousarray(x, np.complex64).copy()
    lib().orc_fft16384(_fp(buf.view(np.float32)), sign)
    return buf


def search(packed, sats, sel=None, params=None, spectra=None, want_grid=False, nthreads=0):
    """Run the oracle search on one capture (k_noncoh consecutive 8192-byte blocks).
    Returns records (structured array, one per selected sat) and optionally the (n_sel, n_dop) grid."""
    packed = np.ascontiguousarray(packed, np.uint8)
    p = params or default_params()
    assert packed.size == p.k_noncoh * block_bytes(p.sample_bits), (packed.size, p.k_noncoh, p.sample_bits)
    arr = sat_array(sats)
    n_sel = len(sats) if sel is None else len(sel)
    sel_a = None if sel is None else np.ascontiguousarray(sel, np.int32)
    out = np.zeros(n_sel, RECORD_DTYPE)
    n_dop = p.dop_hi - p.dop_lo + 1
    grid = np.zeros((n_sel, n_dop), CELL_DTYPE) if want_grid else None
    sp = None
    if spectra is not None:
        sp = np.ascontiguousarray(spectra, np.complex64)
        assert sp.shape == (len(sats), N)
    rc = lib().orc_search_pre(_u8(packed), arr, len(sats), _fp(sp.view(np.float32)) if sp is not None else None,
                              sel_a.ctypes.data_as(C.POINTER(C.c_int32)) if sel_a is not None else None,
                              n_sel, C.byref(p), out.ctypes.data, grid.ctypes.data if want_grid else None,
                              nthreads)
    if rc != 0:
        raise RuntimeError("orc_search_pre failed: %d" % rc)
    return (out, grid) if want_grid else out


def refine(packed, sats, records, params=None, nthreads=0):
    """Oracle refinement of `records` (from search() on the same capture): FINE_DTYPE array, one per record."""
    packed = np.ascontiguousarray(packed, np.uint8)
    p = params or default_params()
    assert packed.size == p.k_noncoh * block_bytes(p.sample_bits), (packed.size, p.k_noncoh, p.sample_bits)
    rec = np.ascontiguousarray(records, RECORD_DTYPE)
    out = np.zeros(rec.size, FINE_DTYPE)
    rc = lib().orc_refine(_u8(packed), sat_array(sats), len(sats), C.byref(p), rec.ctypes.data, rec.size,
                          out.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError("orc_refine failed: %d" % rc)
    return out


def gen_capture(seed, n_blocks, sats, signals, sample_bits=1, mag_thr=0.98, code_doppler=False):
    """signals: iterable of dicts/tuples (sat, tau, doppler_hz, cn0_dbhz, phase[, flip_ms]).
    sample_bits=2: sign plane + magnitude plane per block, mag = |s| > mag_thr noise sigmas (same sign bits)."""
    arr = sat_array(sats)
    sig = (OrcSignal * max(1, len(signals)))()
    for i, s in enumerate(signals):
        if isinstance(s, dict):
            s = (s["sat"], s["tau"], s["doppler_hz"], s["cn0_dbhz"], s.get("phase", 0.0), s.get("flip_ms", 0))
        s = tuple(s) + (0,) * (6 - len(s))
        sig[i].sat, sig[i].tau = int(s[0]), int(s[1])
        sig[i].doppler_hz, sig[i].cn0_dbhz, sig[i].phase = float(s[2]), float(s[3]), float(s[4])
        sig[i].flip_ms = int(s[5])
        sig[i].code_doppler = 1 if code_doppler else 0
    out = np.zeros(n_blocks * block_bytes(sample_bits), np.uint8)
    rc = lib().orc_gen_capture_sm(int(seed), n_blocks, arr, len(sats), sig, len(signals), int(sample_bits),
                                  float(mag_thr), _u8(out))
    if rc != 0:
        raise RuntimeError("orc_gen_capture_sm failed: %d" % rc)
    return out


# ---------------------------------------------------------------- reference (oracle/_ref) helpers
def ref_sats():
    R = ref()
    out = []
    for i in range(R.ref_n_sats()):
        v = [C.c_int32() for _ in range(4)]
        R.ref_sat(i, *[C.byref(x) for x in v])
        out.append(tuple(x.value for x in v))
    return out


def ref_code_spectrum(sat):
    out = np.zeros(2 * N, np.float32)
    ref().ref_code_spectrum(sat, _fp(out))
    return out.view(np.complex64)


def ref_code_baseband(sat):
    out = np.zeros(2 * N, np.float32)
    rc = ref().ref_code_baseband(sat, _fp(out))
    assert rc == 0
    return out.view(np.complex64)


def ref_sample(packed):
    packed = np.ascontiguousarray(packed, np.uint8)
    assert packed.size == BLOCK_BYTES
    x2 = np.zeros(2 * N, np.float32)
    D = np.zeros(2 * N, np.float32)
    ref().ref_sample(_u8(packed), _fp(x2), _fp(D))
    return x2.view(np.complex64), D.view(np.complex64)


def ref_search(packed, sel, lib=None):
    """Sample()+Correlate() of the unmodified reference for each sat index in sel (lib: a handle from ref_bench())."""
    packed = np.ascontiguousarray(packed, np.uint8)
    assert packed.size == BLOCK_BYTES
    sel = np.ascontiguousarray(sel, np.int32)
    dop = np.zeros(len(sel), np.int32)
    lag = np.zeros(len(sel), np.int32)
    snr = np.zeros(len(sel), np.float32)
    i32 = C.POINTER(C.c_int32)
    (lib or ref()).ref_search(_u8(packed), sel.ctypes.data_as(i32), len(sel), dop.ctypes.data_as(i32),
                              lag.ctypes.data_as(i32), _fp(snr))
    return dop, lag, snr


def ref_search_task(blocks, passes=1, free_chans=12, min_sig=16, acq=(1, 1, 1)):
    """Run the literal SearchTask loop; returns the event log (structured array)."""
    blocks = np.ascontiguousarray(blocks, np.uint8)
    n_blocks = blocks.size // BLOCK_BYTES
    n = ref().ref_search_task(_u8(blocks), n_blocks, passes, free_chans, min_sig, *[int(a) for a in acq])
    ev = np.zeros(n, EVENT_DTYPE)
    if n:
        ref().ref_get_events(ev.ctypes.data, n)
    return ev
