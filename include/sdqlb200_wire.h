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
    SDQLB200_WIRE_KINDS = 5
};

/* Expand `rows` packed elements at `src` into the resident column `dst` (both DEVICE pointers, 16-byte aligned;
 * dst holds rows doubles or rows int32).  `table`: DEVICE dictionary for the DICT kinds (NULL otherwise);
 * `scale`: divisor for FIXED32 (ignored otherwise).  Stream ordered on `stream` (cudaStream_t); one launch. */
int sdqlb200_wire_decode(int32_t kind, const void* src, void* dst, int64_t rows, const void* table, double scale,
                         void* stream);
/* bytes one packed element / one decoded element of `kind` occupies (0 for an unknown kind) */
int32_t sdqlb200_wire_src_width(int32_t kind);
int32_t sdqlb200_wire_dst_width(int32_t kind);
const char* sdqlb200_wire_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
