"""ctypes binding of the host drop-in shim (include/search_dropin.h) with Python callbacks as the host
side (ChanReset / ChanStart / GPSstat / SPI capture / timer).  Used by the tests to replay the
reference's SearchTask loop against a mock receiver."""
import ctypes as C

import numpy as np

from . import _lib

LITERAL, BATCH = 0, 1

_CHAN_RESET = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int)
_CHAN_START = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
_STAT_SAT = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int)
_STAT_DOP = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)
_CAPTURE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8))
_TIMER = C.CFUNCTYPE(C.c_uint, C.c_void_p)
_YIELD = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


class HostIface(C.Structure):
    _fields_ = [("user", C.c_void_p), ("chan_reset", _CHAN_RESET), ("chan_start", _CHAN_START),
                ("stat_sat", _STAT_SAT), ("stat_dop", _STAT_DOP), ("capture", _CAPTURE), ("timer_us", _TIMER),
                ("yield_", _YIELD)]


class MockReceiver:
    """Stands in for gps/channel.cpp + gps/stat.cpp + the SPI sampler: logs every call."""

    def __init__(self, blocks, free_chans=12):
        self.blocks = np.ascontiguousarray(blocks, np.uint8).reshape(-1, 8192)
        self.free = free_chans
        self.next_ch = 0
        self.time = 0
        self.samples = 0
        self.events = []

    def chan_reset(self, _u, sat, init):
        ch = self.next_ch if self.free > 0 else -1
        self.events.append(("chan_reset", sat, init, ch))
        return ch

    def chan_start(self, _u, ch, sat, t_sample, lo_shift, ca_shift, snr):
        self.events.append(("chan_start", ch, sat, lo_shift, ca_shift, snr))
        self.free -= 1
        self.next_ch += 1

    def stat_sat(self, _u, snr, ch, sat, weak, us):
        self.events.append(("stat_sat", ch, sat, weak, snr))

    def stat_dop(self, _u, ch, lo_hz, ca_shift):
        self.events.append(("stat_dop", ch, lo_hz, ca_shift))

    def capture(self, _u, dst):
        blk = self.blocks[self.samples % len(self.blocks)]
        C.memmove(dst, blk.ctypes.data, 8192)
        self.samples += 1
        return 0

    def timer_us(self, _u):
        self.time += 1000
        return self.time

    def yield_(self, _u, where):
        pass


class Dropin:
    def __init__(self, sats, receiver, device=0):
        L = _lib.load()
        self._L = L
        L.acq_dropin_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_lib.AcqSat), C.c_int, C.POINTER(HostIface), C.c_int]
        L.acq_dropin_pass.argtypes = [C.c_void_p, C.c_int]
        L.acq_dropin_destroy.argtypes = [C.c_void_p]
        L.acq_dropin_enable.argtypes = [C.c_void_p, C.c_int]
        L.acq_dropin_is_busy.argtypes = [C.c_void_p, C.c_int]
        L.acq_dropin_set_acq.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.acq_dropin_set_refine.argtypes = [C.c_void_p, C.c_int]
        L.acq_dropin_params.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
        self.rx = receiver
        self._cbs = HostIface(None, _CHAN_RESET(receiver.chan_reset), _CHAN_START(receiver.chan_start),
                              _STAT_SAT(receiver.stat_sat), _STAT_DOP(receiver.stat_dop), _CAPTURE(receiver.capture),
                              _TIMER(receiver.timer_us), _YIELD(receiver.yield_))
        arr = (_lib.AcqSat * len(sats))()
        for i, s in enumerate(sats):
            arr[i].prn, arr[i].t1, arr[i].t2, arr[i].type = [int(v) for v in s]
        h = C.c_void_p()
        rc = L.acq_dropin_create(C.byref(h), arr, len(sats), C.byref(self._cbs), device)
        if rc != 0:
            raise RuntimeError("acq_dropin_create failed: %d %s" % (rc, L.acq_last_error().decode()))
        self._h = h

    def params(self, *argv):
        a = (C.c_char_p * (len(argv) + 1))(b"kiwid", *[x.encode() for x in argv])
        return self._L.acq_dropin_params(self._h, len(argv) + 1, a)

    def set_acq(self, navstar=1, qzss=1, galileo=1):
        return self._L.acq_dropin_set_acq(self._h, navstar, qzss, galileo)

    def set_refine(self, on=True):
        """Hand ChanStart the acq_refine values (FS-sample ca_shift, nearest Doppler bin); off by default."""
        return self._L.acq_dropin_set_refine(self._h, int(bool(on)))

    def search_pass(self, mode=LITERAL):
        rc = self._L.acq_dropin_pass(self._h, mode)
        if rc < 0:
            raise RuntimeError("acq_dropin_pass failed: %d %s" % (rc, self._L.acq_last_error().decode()))
        return rc

    def enable(self, sat):
        return self._L.acq_dropin_enable(self._h, sat)

    def is_busy(self, sat):
        return self._L.acq_dropin_is_busy(self._h, sat)

    def close(self):
        if self._h:
            self._L.acq_dropin_destroy(self._h)
            self._h = None
