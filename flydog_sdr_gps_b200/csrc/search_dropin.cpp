// search_dropin.cpp -- see search_dropin.h.  Host C++ above the C ABI; no CUDA types here.
#include "../../include/search_dropin.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

struct acq_dropin {
    acq_engine *eng = nullptr;
    acq_host_iface host{};
    std::vector<acq_sat> sats;
    std::vector<char> busy;
    std::vector<uint8_t> mask;  // empty: every satellite; else mask[sat] == 0 skips it (gps_debug / gps_e1b_only)
    int minimum_sig = 16;  // MIN_SIG (gps/gps.h:60), overridable with -gsig (search.cpp:82-84)
    int test_mode = 0;
    int refine = 0;  // acq_dropin_set_refine: hand ChanStart the interpolated code phase / nearest Doppler bin
    int acq_navstar = 1, acq_qzss = 1, acq_galileo = 1;
    // state that SearchTask keeps across satellites and passes (search.cpp:513-515)
    int last_ch = -1, lo_shift = 0, ca_shift = 0;
    float snr = 0;
    std::vector<uint8_t> capture;
};

namespace {

const int G2_INIT = 0x400;   // gps/gps.h:127
const int E1B_MODE = 0x800;  // kiwi.config:269

bool enabled(const acq_dropin *d, const acq_sat &s)
{
    if (!d->mask.empty() && !d->mask[&s - d->sats.data()]) return false;  // search.cpp:537-539
    if (s.type == ACQ_NAVSTAR && !d->acq_navstar) return false;  // search.cpp:533-535
    if (s.type == ACQ_QZSS && !d->acq_qzss) return false;
    if (s.type == ACQ_E1B && !d->acq_galileo) return false;
    return true;
}

int codegen_init(const acq_sat &s)
{
    switch (s.type) {  // search.cpp:559-563
        case ACQ_QZSS: return G2_INIT | s.t2;
        case ACQ_E1B: return E1B_MODE | (s.prn - 1);
        default: return (s.t1 << 4) + s.t2;
    }
}

int min_sig(const acq_dropin *d, const acq_sat &s) { return s.type == ACQ_E1B ? 16 : d->minimum_sig; }  // search.cpp:549

void yield(acq_dropin *d, const char *where)
{
    if (d->host.yield) d->host.yield(d->host.user, where);
}

}  // namespace

extern "C" {

int acq_dropin_create(acq_dropin **out, const acq_sat *sats, int n_sats, const acq_host_iface *host, int device)
{
    if (!out || !sats || n_sats <= 0 || !host || !host->chan_reset || !host->chan_start || !host->capture ||
        !host->timer_us)
        return ACQ_ERR_ARG;
    *out = nullptr;
    acq_dropin *d = new (std::nothrow) acq_dropin;
    if (!d) return ACQ_ERR_NOMEM;
    acq_params prm;
    acq_params_default(&prm);  // the reference's compile-time search: -20..+20 bins, K = 1, reference wrap
    const int rc = acq_create(&d->eng, &prm, sats, n_sats, device);
    if (rc != ACQ_OK) {
        delete d;
        return rc;
    }
    d->host = *host;
    d->sats.assign(sats, sats + n_sats);
    d->busy.assign(n_sats, 0);
    d->capture.resize(ACQ_BLOCK_BYTES);
    *out = d;
    return ACQ_OK;
}

int acq_dropin_destroy(acq_dropin *d)
{
    if (!d) return ACQ_OK;
    acq_destroy(d->eng);
    delete d;
    return ACQ_OK;
}

int acq_dropin_params(acq_dropin *d, int argc, char *argv[])
{
    if (!d) return ACQ_ERR_ARG;
    for (int i = 1; i < argc;) {  // same scan as search.cpp:75-94
        const char *v = argv[i];
        if (strcmp(v, "-gsig") == 0) {
            i++;
            if (i < argc) d->minimum_sig = (int)strtol(argv[i], 0, 0);
        } else if (strcmp(v, "-gt") == 0) {
            d->test_mode = 1;
        }
        i++;
        while (i < argc && ((argv[i][0] != '+') && (argv[i][0] != '-'))) i++;
    }
    return ACQ_OK;
}

int acq_dropin_set_acq(acq_dropin *d, int navstar, int qzss, int galileo)
{
    if (!d) return ACQ_ERR_ARG;
    d->acq_navstar = navstar;
    d->acq_qzss = qzss;
    d->acq_galileo = galileo;
    return ACQ_OK;
}

int acq_dropin_set_mask(acq_dropin *d, const uint8_t *mask, int n_sats)
{
    if (!d) return ACQ_ERR_ARG;
    if (!mask) {
        d->mask.clear();
        return ACQ_OK;
    }
    if (n_sats != (int)d->sats.size()) return ACQ_ERR_ARG;
    d->mask.assign(mask, mask + n_sats);
    return ACQ_OK;
}

int acq_dropin_min_sig(const acq_dropin *d) { return d ? d->minimum_sig : ACQ_ERR_ARG; }

int acq_dropin_set_refine(acq_dropin *d, int on)
{
    if (!d) return ACQ_ERR_ARG;
    d->refine = on ? 1 : 0;
    return ACQ_OK;
}

int acq_dropin_enable(acq_dropin *d, int sat)
{
    if (!d || sat < 0 || sat >= (int)d->sats.size()) return ACQ_ERR_ARG;
    d->busy[sat] = 0;
    return ACQ_OK;
}

int acq_dropin_is_busy(const acq_dropin *d, int sat)
{
    if (!d || sat < 0 || sat >= (int)d->sats.size()) return ACQ_ERR_ARG;
    return d->busy[sat];
}

acq_engine *acq_dropin_engine(acq_dropin *d) { return d ? d->eng : nullptr; }

static int pass_literal(acq_dropin *d)
{
    int started = 0;
    const acq_host_iface &h = d->host;
    for (int sat = 0; sat < (int)d->sats.size(); sat++) {
        const acq_sat &sp = d->sats[sat];
        if (!enabled(d, sp)) continue;
        const int msig = min_sig(d, sp);
        if (d->busy[sat]) {  // search.cpp:551-554
            yield(d, "busy1");
            continue;
        }
        const int ch = h.chan_reset(h.user, sat, codegen_init(sp));  // search.cpp:565
        if (ch < 0) continue;
        if (d->last_ch != ch && d->snr < msig && h.stat_sat) h.stat_sat(h.user, 0, d->last_ch, -1, 0, 0);  // :569
        const unsigned t0 = h.timer_us(h.user);  // us = t_sample = timer_us()  (:571)
        if (h.capture(h.user, d->capture.data()) != 0) return ACQ_ERR_ARG;  // Sample(), SPI half (:395-406)
        acq_record rec;
        const int32_t one = sat;
        const int rc = acq_search(d->eng, d->capture.data(), 1, &one, 1, &rec);  // Sample() DSP + Correlate()
        if (rc != ACQ_OK) return rc;
        d->snr = rec.snr;
        if (rec.snr > 0) {  // Correlate() leaves the caller's variables untouched otherwise (:455,495)
            d->lo_shift = rec.dop;
            d->ca_shift = rec.lag;
        }
        d->ca_shift *= ACQ_DECIM;  // :575
        const int us = (int)(h.timer_us(h.user) - t0);  // :577
        if (h.stat_sat) h.stat_sat(h.user, d->snr, ch, sat, d->snr < msig, us);  // :580
        d->last_ch = ch;
        if (d->snr < msig) continue;  // :591
        if (d->refine && rec.snr > 0) {  // extension (off by default): FS-sample code phase, nearest Doppler bin
            acq_fine fine;
            const int rr = acq_refine(d->eng, &rec, 1, &fine);
            if (rr != ACQ_OK) return rr;
            d->ca_shift = fine.ca_shift;
            d->lo_shift = (int)lrintf(fine.dop_hz / (float)ACQ_BIN_HZ);
        }
        if (h.stat_dop) h.stat_dop(h.user, ch, (int)(d->lo_shift * (float)ACQ_BIN_HZ), d->ca_shift);  // :595
        d->busy[sat] = 1;  // :597
        h.chan_start(h.user, ch, sat, (int)t0, d->lo_shift, d->ca_shift, (int)d->snr);  // :601
        started++;
    }
    return started;
}

static int pass_batch(acq_dropin *d)
{
    const acq_host_iface &h = d->host;
    std::vector<int32_t> sel;
    for (int sat = 0; sat < (int)d->sats.size(); sat++)
        if (enabled(d, d->sats[sat]) && !d->busy[sat]) sel.push_back(sat);
    if (sel.empty()) {
        yield(d, "busy1");
        return 0;
    }
    const unsigned t0 = h.timer_us(h.user);
    if (h.capture(h.user, d->capture.data()) != 0) return ACQ_ERR_ARG;
    std::vector<acq_record> rec(sel.size());
    const int rc = acq_search(d->eng, d->capture.data(), 1, sel.data(), (int)sel.size(), rec.data());
    if (rc != ACQ_OK) return rc;
    std::vector<acq_fine> fine;
    if (d->refine) {
        fine.resize(sel.size());
        const int rr = acq_refine(d->eng, rec.data(), (int)rec.size(), fine.data());
        if (rr != ACQ_OK) return rr;
    }
    const int us = (int)(h.timer_us(h.user) - t0);
    int started = 0;
    for (size_t i = 0; i < sel.size(); i++) {
        const int sat = sel[i];
        const int msig = min_sig(d, d->sats[sat]);
        if (rec[i].snr < msig) {
            if (h.stat_sat) h.stat_sat(h.user, rec[i].snr, -1, sat, 1, us);
            continue;
        }
        const int ch = h.chan_reset(h.user, sat, codegen_init(d->sats[sat]));
        if (ch < 0) break;  // no free tracking channel left
        const int ca_shift = d->refine ? fine[i].ca_shift : rec[i].lag * ACQ_DECIM;
        const int lo_shift = d->refine ? (int)lrintf(fine[i].dop_hz / (float)ACQ_BIN_HZ) : rec[i].dop;
        if (h.stat_sat) h.stat_sat(h.user, rec[i].snr, ch, sat, 0, us);
        if (h.stat_dop) h.stat_dop(h.user, ch, (int)(lo_shift * (float)ACQ_BIN_HZ), ca_shift);
        d->busy[sat] = 1;
        h.chan_start(h.user, ch, sat, (int)t0, lo_shift, ca_shift, (int)rec[i].snr);
        started++;
    }
    return started;
}

int acq_dropin_pass(acq_dropin *d, int mode)
{
    if (!d) return ACQ_ERR_ARG;
    if (mode == ACQ_DROPIN_LITERAL) return pass_literal(d);
    if (mode == ACQ_DROPIN_BATCH) return pass_batch(d);
    return ACQ_ERR_ARG;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// capture sources
// ---------------------------------------------------------------------------------------------
struct acq_capture_file {
    FILE *fp = nullptr;
    long long size = 0, pos = 0;
};

extern "C" {

int acq_capture_from_packets(const uint8_t *const *packets, int n_packets, int packet_bytes, uint8_t *dst)
{
    if (!packets || !dst || n_packets <= 0 || packet_bytes <= 0) return ACQ_ERR_ARG;
    if ((long long)n_packets * packet_bytes != ACQ_BLOCK_BYTES) return ACQ_ERR_ARG;
    for (int k = 0; k < n_packets; k++) {
        if (!packets[k]) return ACQ_ERR_ARG;
        memcpy(dst + (size_t)k * packet_bytes, packets[k], packet_bytes);  // search.cpp:399-406
    }
    return ACQ_OK;
}

int acq_capture_file_open(acq_capture_file **out, const char *path)
{
    if (!out || !path) return ACQ_ERR_ARG;
    *out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) return ACQ_ERR_ARG;
    acq_capture_file *f = new (std::nothrow) acq_capture_file;
    if (!f || fseek(fp, 0, SEEK_END) != 0) {
        fclose(fp);
        delete f;
        return ACQ_ERR_ARG;
    }
    f->fp = fp;
    f->size = ftell(fp);
    rewind(fp);
    *out = f;
    return ACQ_OK;
}

int acq_capture_file_next(acq_capture_file *f, uint8_t *dst, int n_blocks)
{
    if (!f || !f->fp || !dst || n_blocks <= 0) return ACQ_ERR_ARG;
    const long long want = (long long)n_blocks * ACQ_BLOCK_BYTES;
    if (f->size - f->pos < want) return ACQ_CAPTURE_EOF;  // search.cpp:375-378
    if (fread(dst, 1, (size_t)want, f->fp) != (size_t)want) return ACQ_ERR_ARG;
    f->pos += want;
    return ACQ_OK;
}

long long acq_capture_file_remaining(const acq_capture_file *f)
{
    return (f && f->fp) ? (f->size - f->pos) / ACQ_BLOCK_BYTES : 0;
}

int acq_capture_file_rewind(acq_capture_file *f)
{
    if (!f || !f->fp) return ACQ_ERR_ARG;
    rewind(f->fp);
    f->pos = 0;
    return ACQ_OK;
}

int acq_capture_file_close(acq_capture_file *f)
{
    if (!f) return ACQ_OK;
    if (f->fp) fclose(f->fp);
    delete f;
    return ACQ_OK;
}

int acq_capture_file_iface(void *user, uint8_t *dst)
{
    return acq_capture_file_next(static_cast<acq_capture_file *>(user), dst, 1);
}

}  // extern "C"
