/* sdqlpy-b200 column wire formats -- C ABI of libsdqlb200_wire.so.
 *
 * What it replaces in the reference (edin-dal/sdqlpy): nothing is decoded there -- read_csv materialises every
 * column as int64 / float64 / UCS4 (sdql_lib.py:83-97) and the generated module borrows those buffers
 * (sdql_compiler.py:644-668).  On a B200 the host->device link (PCIe) is ~100x slower than HBM, so columns cross it
 * in a LOSSLESS packed form and are expanded on the device into the resident layout the query kernels read
 * (DESIGN.md section 3).  decode(encode(x)) == x bit for bit; the encoder (sdqlpy_b200/wire.py) verifies that per
 * column and keeps the plain representation when no packed form is exact.
 *
 * Conventions as in sdqlb200.h: plain pointers and sizes, 0 on success, negative SDQLB200_E_* otherwise.
 */
#ifndef SDQLB200_WIRE_H
#define SDQLB200_WIRE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    SDQLB200_WIRE_DICT8_F64 = 0,   /* uint8  code -> double table[code]  (<= 256 distinct values)            */
    SDQLB200_WIRE_DICT16_F64 = 1,  /* uint16 code -> double table[code]  (<= 65536 distinct values)          */
    SDQLB200_WIRE_DICT8_I32 = 2,   /* uint8  code -> int32  table[code]                                       */
    SDQLB200_WIRE_DICT16_I32 = 3,  /* uint16 code -> int32  table[code]  (dates: ~2500 distinct YYYYMMDD)     */
    SDQLB200_WIRE_FIXED32_F64 = 4, /* int32 v -> (double)v / scale       (decimal(.,2) money: scale = 100)    */
    SDQLB200_WIRE_KINDS = 5,
    /* bit-packed kinds (sdqlb200_wire_decode_bits): element i is the nbits-wide field at bit i*nbits of a little-endian
     * bit stream (bit b of the stream = bit b%8 of byte b/8); 1 <= nbits <= 32 */
    SDQLB200_WIRE_BITS_DICT_F64 = 5,  /* code -> double table[code]                                           */
    SDQLB200_WIRE_BITS_DICT_I32 = 6,  /* code -> int32  table[code]                                           */
    SDQLB200_WIRE_BITS_FIXED_F64 = 7, /* code -> (double)(base + code) / scale   (frame-of-reference decimals) */
    SDQLB200_WIRE_BITS_I32 = 8,       /* code -> (int32)(base + code)            (frame-of-reference integers) */
    SDQLB200_WIRE_BITS_U8 = 9,        /* code -> uint8 code                      (dictionary-coded strings)    */
    SDQLB200_WIRE_ALL_KINDS = 10
};

/* Expand `rows` packed elements at `src` into the resident column `dst` (both DEVICE pointers, 16-byte aligned;
 * dst holds rows doubles or rows int32).  `table`: DEVICE dictionary for the DICT kinds (NULL otherwise);
 * `scale`: divisor for FIXED32 (ignored otherwise).  Stream ordered on `stream` (cudaStream_t); one launch. */
int sdqlb200_wire_decode(int32_t kind, const void* src, void* dst, int64_t rows, const void* table, double scale,
                         void* stream);
/* Bit-packed kinds: expand `rows` nbits-wide fields of the bit stream at `src` (DEVICE, 16-byte aligned, readable up
 * to 8 bytes past the last field) into `dst` (DEVICE, 16-byte aligned: rows doubles / int32 / uint8).  `table`: DEVICE
 * dictionary for the BITS_DICT kinds; `base`, `scale`: frame of reference for BITS_FIXED_F64 / BITS_I32. */
int sdqlb200_wire_decode_bits(int32_t kind, const void* src, void* dst, int64_t rows, int32_t nbits, const void* table,
                              int64_t base, double scale, void* stream);
/* bytes one packed element / one decoded element of `kind` occupies (0 for an unknown kind; src width 0 for the
 * bit-packed kinds, whose width is nbits / 8) */
int32_t sdqlb200_wire_src_width(int32_t kind);
int32_t sdqlb200_wire_dst_width(int32_t kind);
const char* sdqlb200_wire_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
