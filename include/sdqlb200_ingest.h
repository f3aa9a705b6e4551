/* sdqlpy-b200 column ingest -- C ABI of libsdqlb200_ingest.so: reference-layout columns -> resident layout, on the device.
 *
 * What it replaces in the reference (edin-dal/sdqlpy): the generated module borrows the numpy buffers of `db` as they are
 * -- int64 for int / date, float64, UCS4 `<U n` for string(n) -- by casting PyArray_DATA (sdql_compiler.py:644-668).
 * The B200 backend keeps int32 / fp64 / dictionary codes / fixed-width bytes resident instead; these
 * kernels make that conversion part of the upload: the raw buffer crosses PCIe once, unchanged, and is narrowed / encoded
 * where the bandwidth is (HBM), instead of in numpy on the host (astype, np.unique).
 *
 * All pointers are DEVICE pointers unless marked; stream ordered; 0 on success, negative SDQLB200_E_* otherwise.
 */
#ifndef SDQLB200_INGEST_H
#define SDQLB200_INGEST_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* int64 -> int32 (schema types int / date).  d_minmax[0] = min, [1] = max over all calls since it was set to
 * {INT64_MAX, INT64_MIN}: the column statistics (sdqlb200_col.min / max) and the int32 range check come from it. */
int sdqlb200_ingest_i64(const int64_t* d_in, int32_t* d_out, int64_t n, int64_t* d_minmax, void* stream);
/* UCS4 `<U nchar` rows -> `width` bytes per row (string(n) columns that are pattern-matched), zero padded / truncated.
 * *d_bad is set to 1 + the first row with a code point > 255. */
int sdqlb200_ingest_ucs4_bytes(const uint32_t* d_in, uint8_t* d_out, int64_t rows, int32_t nchar, int32_t width,
                               unsigned long long* d_bad, void* stream);
/* Dictionary encoding of a `<U nchar` column, pass 1: every row's string is hashed (64 bit) into an open-addressing set
 * of `cap` slots (power of two; d_keys filled with 0xFF before the first call); d_rep[slot] = the smallest global row id
 * (row_base + i) holding that value.  *d_count = distinct values so far (stops growing at cap / 2: overflow). */
int sdqlb200_ingest_ucs4_distinct(const uint32_t* d_in, int64_t rows, int64_t row_base, int32_t nchar,
                                  unsigned long long* d_keys, long long* d_rep, int64_t cap,
                                  unsigned long long* d_count, void* stream);
/* pass 2: d_out[i] = d_slot_code[slot of row i] as uint8 (out_width 1) or int32 (4).  Every row is compared with the
 * dictionary entry of its code (d_dict: ndict x nchar UCS4) -- a 64-bit hash collision cannot go unnoticed:
 * *d_bad = 1 + first mismatching row. */
int sdqlb200_ingest_ucs4_codes(const uint32_t* d_in, int64_t rows, int32_t nchar, const unsigned long long* d_keys,
                               const int32_t* d_slot_code, int64_t cap, const uint32_t* d_dict, void* d_out,
                               int32_t out_width, unsigned long long* d_bad, void* stream);
/* d_out[i] = d_table[d_in[i]] as uint8 (out_width 1) or int32 (4): provisional (first-seen) codes -> final dictionary order */
int sdqlb200_ingest_remap(const int32_t* d_in, const int32_t* d_table, void* d_out, int32_t out_width, int64_t n, void* stream);
/* d_out[i] = d_table256[d_in[i]]: bytes of a narrowed `<U1` column -> dictionary codes (in place allowed) */
int sdqlb200_ingest_remap_u8(const uint8_t* d_in, const uint8_t* d_table256, uint8_t* d_out, int64_t n, void* stream);

/* HOST side of the upload (round 2).  With the inputs in host memory on every call the step is bound by the PCIe link, and
 * int64 keys / dates and UCS4 `<U1` flags carry 4 resp. 3 dead bytes per value across it.  These passes narrow such columns
 * with `threads` host threads into (page-locked) staging buffers WHILE the fp64 columns -- which need no host work --
 * occupy the link; the narrowed image then crosses instead of the raw one (Q1: 38 instead of 48 bytes per row).
 * h_* are HOST pointers.  Both return 0, or SDQLB200_E_ARG. */
/* int64 -> int32; minmax[0] = min, minmax[1] = max of the column (the caller range-checks: values are truncated) */
int sdqlb200_ingest_host_i64(const int64_t* h_in, int32_t* h_out, int64_t n, int32_t threads, int64_t* minmax);
/* `<U1` (one UCS4 code point per row) -> one byte per row; present[b >> 6] bit (b & 63) set for every byte value seen;
 * *bad_row = first row with a code point > 255, or -1 */
int sdqlb200_ingest_host_ucs4_1(const uint32_t* h_in, uint8_t* h_out, int64_t n, int32_t threads, uint64_t present[4],
                                int64_t* bad_row);
const char* sdqlb200_ingest_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
