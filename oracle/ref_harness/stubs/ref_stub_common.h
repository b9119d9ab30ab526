/*
 * ref_stub_common.h -- minimal stand-ins for the reference's runtime headers so that the
 * UNMODIFIED /root/reference/gps/search.cpp compiles as a test oracle (oracle/_ref).
 * TEST INFRASTRUCTURE ONLY.  These declare just the names search.cpp / gps.h use
 * (reference gps/search.cpp:21-34, gps/gps.h:23-27); constant values are the ones the
 * reference's build generates from kiwi.config:243-271.
 */
#pragma once
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

/* kiwi.config:243-271 (generated into kiwi.gen.h by the reference's build) */
#define GPS_MAX_CHANS 12
#define GPS_INTEG_BITS 20
#define GPS_RPT 32
#define GPS_SAMPS 256
#define GPS_SAMPS_RPT GPS_RPT
#define GPS_IQ_SAMPS 255
#define GPS_IQ_SAMPS_W (GPS_IQ_SAMPS * 4)
#define L1_CODELEN 1023
#define E1B_CODELEN 4092
#define E1B_MODE 0x800

/* globals search.cpp reads (kiwi.h / rx / update state) */
extern int gps_chans, gps_debug, gps_e1b_only;
extern bool update_in_progress, sd_copy_in_progress, backup_in_progress, is_locked;
void kiwi_exit(int err);

/* rx_util.h */
#define EXTERNAL_ONLY 1
int rx_count_server_conns(int what);

/* clk.h */
struct ref_clk_t { int adc_gps_clk_corrections; };
extern ref_clk_t clk;

/* cfg.h */
#define CFG_REQUIRED 1
bool admcfg_bool(const char *name, bool *error, int flags);

/* misc.h / timer */
unsigned timer_us(void);

/* coroutines.h: cooperative scheduler calls become no-ops (single-threaded harness) */
#define GPS_ACQ_PRIORITY 2
#define CTF_NO_PRIO_INV 0
#define NT_LONG_RUN 1
typedef void (*ref_task_fn)(void *);
void NextTask(const char *where);
void NextTaskP(const char *where, int prio);
void TaskSleepUsec(int us);
void TaskSleepSec(int s);
int TaskID(void);
void TaskSleepID(int id, int us);
void TaskWakeup(int id);
int CreateTaskF(ref_task_fn fn, void *param, int prio, int flags);

/* spi.h / spi_dev.h: the capture arrives as 16 x 512-byte packets (search.cpp:389-406) */
typedef enum { CmdSample = 1, CmdGetGPSSamples = 2 } SPI_CMD;
struct SPI_MISO { char byte[2048]; };
struct ref_spi_shmem_t { SPI_MISO gps_search_miso; };
extern ref_spi_shmem_t *SPI_SHMEM;
void spi_set(SPI_CMD cmd, int a = 0, int b = 0);
void spi_get(SPI_CMD cmd, SPI_MISO *rx, int bytes, int a = 0, int b = 0);
