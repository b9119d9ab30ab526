"""Capture sources on the input side of the path (ctypes binding of include/search_dropin.h, "capture sources").

The wire format is the FPGA sampler's (gps/search.cpp:389-411): 65536 one-bit samples per capture, LSB first,
delivered as 16 SPI packets of 512 bytes, or -- with the reference's GPS_SAMPLES_FROM_FILE switch
(gps/search.cpp:361-380) -- read from a raw file of consecutive captures."""
import ctypes as C

import numpy as np

from . import _lib

BLOCK_BYTES = 8192
CAPTURE_EOF = 1


def _L():
    L = _lib.load()
    L.acq_capture_from_packets.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    L.acq_capture_file_open.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
    L.acq_capture_file_next.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.acq_capture_file_remaining.argtypes = [C.c_void_p]
    L.acq_capture_file_remaining.restype = C.c_longlong
    L.acq_capture_file_rewind.argtypes = [C.c_void_p]
    L.acq_capture_file_close.argtypes = [C.c_void_p]
    return L


def from_packets(packets):
    """Concatenate SPI packets (sequence of equal-length uint8 arrays, 16 x 512 on the hardware) into one block."""
    pk = [np.ascontiguousarray(p, np.uint8) for p in packets]
    n, size = len(pk), (pk[0].size if pk else 0)
    if any(p.size != size for p in pk):
        raise ValueError("packets must have equal length")
    ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pk])
    out = np.empty(BLOCK_BYTES, np.uint8)
    rc = _L().acq_capture_from_packets(ptrs, n, size, out.ctypes.data)
    if rc != 0:
        raise ValueError("acq_capture_from_packets: %d packets x %d bytes is not one 8192-byte capture" % (n, size))
    return out


class CaptureFile:
    """Raw 1-bit capture file in the GPS_SAMPLES_FROM_FILE format: next(n_blocks) returns the next capture
    (n_blocks x 8192 bytes) or None at end of file (where the reference exits, gps/search.cpp:375-378)."""

    def __init__(self, path):
        self._L = _L()
        self._h = C.c_void_p()
        if self._L.acq_capture_file_open(C.byref(self._h), str(path).encode()) != 0:
            raise OSError("cannot open capture file %s" % path)

    def next(self, n_blocks=1):
        out = np.empty(n_blocks * BLOCK_BYTES, np.uint8)
        rc = self._L.acq_capture_file_next(self._h, out.ctypes.data, n_blocks)
        if rc == CAPTURE_EOF:
            return None
        if rc != 0:
            raise OSError("read error in capture file")
        return out

    def remaining(self):
        return int(self._L.acq_capture_file_remaining(self._h))

    def rewind(self):
        self._L.acq_capture_file_rewind(self._h)

    def close(self):
        if self._h:
            self._L.acq_capture_file_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
