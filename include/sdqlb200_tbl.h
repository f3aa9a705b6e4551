/* sdqlpy-b200 .tbl reader -- C ABI of libsdqlb200_tbl.so: device-side parsing of dbgen's pipe-delimited text.
 *
 * What it replaces in the reference (edin-dal/sdqlpy): read_csv (sdql_lib.py:118-128 -> 69-115), a Python csv.reader
 * row loop that converts every field with int() / float() / int(v.replace("-", "")) and materialises int64 / float64 /
 * UCS4 numpy columns (sdql_lib.py:83-97); test_all.py:35-42 loads the eight TPC-H tables with it.  Here the file's
 * bytes are copied to the device as they are and three kernels do the work:
 *   index  count the newlines of every 4 KB tile, prefix-sum the tile counts, write the byte offset of every row
 *   parse  one thread per row walks the row's fields and writes the wanted columns in the resident device layout:
 *          int / date -> int32, float -> fp64 (correctly rounded, see below), string(n) -> n bytes, zero padded
 * Semantics for well-formed input are those of the reference's conversions: int -> int(v); date "YYYY-MM-DD" ->
 * int(v.replace("-", "")) (sdql_lib.py:83-84); float -> float(v) for plain decimal notation: the digits are
 * accumulated exactly as an integer m with k fraction digits and m / 10^k is ONE IEEE division of two exactly
 * representable doubles (m < 2^53, k <= 22), hence the correctly rounded value of the decimal string -- what Python's
 * float() returns; string(n) -> the first n characters (numpy's "<U n" assignment truncates the same way).
 * Not interpreted (reported as a malformed row instead of guessed): quotes, exponents, more than 15 significant
 * digits, characters >= 0x80, values outside int32.  A trailing '\r' of a row is ignored.
 *
 * Conventions as in sdqlb200.h: plain pointers and sizes, 0 on success, negative SDQLB200_E_* otherwise.
 */
#ifndef SDQLB200_TBL_H
#define SDQLB200_TBL_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SDQLB200_TBL_INT = 0, SDQLB200_TBL_FLOAT = 1, SDQLB200_TBL_DATE = 2, SDQLB200_TBL_STR = 3 };
enum {  /* sdqlb200_tbl_status.error */
    SDQLB200_TBL_OK = 0,
    SDQLB200_TBL_E_FIELDS = 1,   /* the row has fewer fields than the schema                        */
    SDQLB200_TBL_E_NUMBER = 2,   /* not a plain decimal number / date (or empty)                    */
    SDQLB200_TBL_E_RANGE = 3,    /* integer outside int32, or more than 15 significant digits      */
    SDQLB200_TBL_E_CHAR = 4      /* character >= 0x80 in a string field                             */
};
enum { SDQLB200_TBL_MAX_COLS = 32, SDQLB200_TBL_TILE = 4096 };

typedef struct {
    int32_t type;   /* SDQLB200_TBL_*                                                                        */
    int32_t width;  /* STR: bytes per row of the output column                                              */
    void* out;      /* DEVICE: rows int32 (INT, DATE) / rows doubles (FLOAT) / rows * width bytes (STR);    */
                    /* NULL = the column is skipped                                                          */
} sdqlb200_tbl_col;

typedef struct {
    int64_t bad_row;  /* lowest malformed row, -1 if none */
    int64_t error;    /* SDQLB200_TBL_E_* of one malformed row (the lowest one when only one row is bad) */
    int64_t min[SDQLB200_TBL_MAX_COLS], max[SDQLB200_TBL_MAX_COLS];  /* value range of every parsed INT / DATE column */
} sdqlb200_tbl_status;

/* bytes of DEVICE scratch sdqlb200_tbl_index needs for a text of `bytes` bytes */
int64_t sdqlb200_tbl_scratch_bytes(int64_t bytes);
/* Rows of the text at d_text[0, bytes) (DEVICE, 16-byte aligned): number of '\n', plus one if the text does not end
 * with one.  Leaves the per-tile prefix sums in d_scratch for sdqlb200_tbl_row_starts.  Synchronises the stream
 * (the row count sizes the caller's output columns). */
int sdqlb200_tbl_index(const void* d_text, int64_t bytes, void* d_scratch, int64_t* rows_out, void* stream);
/* d_starts[r] = byte offset of row r for r in [0, rows]; row r is d_text[d_starts[r], d_starts[r + 1] - 1).
 * d_starts: DEVICE, rows + 1 int64.  d_scratch as left by sdqlb200_tbl_index for the same text. */
int sdqlb200_tbl_row_starts(const void* d_text, int64_t bytes, const void* d_scratch, int64_t* d_starts, int64_t rows,
                            void* stream);
/* Parse `rows` rows into the columns of `cols` (schema order; ncols <= SDQLB200_TBL_MAX_COLS).  d_status: DEVICE
 * sdqlb200_tbl_status, initialised by this call; read it after the stream has been synchronised. */
int sdqlb200_tbl_parse(const void* d_text, const int64_t* d_starts, int64_t rows, const sdqlb200_tbl_col* cols,
                       int32_t ncols, char delimiter, sdqlb200_tbl_status* d_status, void* stream);
const char* sdqlb200_tbl_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
