/* sdqlpy-b200 device-side TPC-H generator -- C ABI of libsdqlb200_tpchgen.so.
 *
 * What it replaces in the reference (edin-dal/sdqlpy): read_csv of the dbgen .tbl files (sdql_lib.py:69-128,
 * test/test_all.py:35-42) -- a Python csv loop that cannot produce SF100 (600 M lineitem rows) in any useful time,
 * and there is no dbgen / network here anyway.  The two fact tables are generated straight into HBM in the resident
 * columnar layout (int32 keys / dates as YYYYMMDD, float64 measures, uint8 dictionary codes, fixed-width ASCII), with
 * the same counter-based arithmetic as the host generator sdqlpy_b200/tpch/gen.py: every value is a pure function of
 * (seed, stream, row or order index), so host and device columns are bit-identical and any order range can be
 * produced independently (range partitioning across GPUs).
 *
 * All pointers are DEVICE pointers; a NULL column is skipped.  0 on success, negative SDQLB200_E_* otherwise.
 */
#ifndef SDQLB200_TPCHGEN_H
#define SDQLB200_TPCHGEN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int64_t seed;
    int64_t S, P, C, O; /* suppliers, parts, customers, orders of the whole database (gen.TPCH.__init__) */
} sdqlb200_tpch_params;

typedef struct {
    int32_t *l_orderkey, *l_partkey, *l_suppkey, *l_linenumber, *l_shipdate, *l_commitdate, *l_receiptdate;
    double *l_quantity, *l_extendedprice, *l_discount, *l_tax;
    uint8_t *l_returnflag, *l_linestatus, *l_shipinstruct, *l_shipmode; /* dictionary codes (gen.RFLAGS, ...) */
} sdqlb200_lineitem_cols;

typedef struct {
    int32_t *o_orderkey, *o_custkey, *o_orderdate, *o_shippriority;
    double* o_totalprice;
    uint8_t *o_orderstatus, *o_orderpriority; /* dictionary codes (gen.OSTATUS, gen.PRIORITIES) */
    uint8_t* o_comment;                      /* 79 bytes per order, zero padded */
} sdqlb200_orders_cols;

/* lines per order for orders [o0, o1): nl[i - o0] in 1..7 */
int sdqlb200_tpchgen_order_lines(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, int32_t* nl, void* stream);
/* lineitem rows of orders [o0, o1).  off[i - o0] = GLOBAL lineitem row id of the first line of order i
 * (exclusive prefix sum of the lines per order from order 0), o1 - o0 + 1 entries; rows are written at
 * off[i - o0] - off[0] + line. */
int sdqlb200_tpchgen_lineitem(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, const int64_t* off,
                              const sdqlb200_lineitem_cols* cols, void* stream);
/* orders [o0, o1); o_orderstatus / o_totalprice are folded from the order's lines as in gen.py */
int sdqlb200_tpchgen_orders(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, const int64_t* off,
                            const sdqlb200_orders_cols* cols, void* stream);
/* as above, plus o_comment: `vocab` = DEVICE table of `nwords` 12-byte space padded words (gen.WORDS + " ") */
int sdqlb200_tpchgen_orders_text(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, const int64_t* off,
                                 const sdqlb200_orders_cols* cols, const uint8_t* vocab, int32_t nwords, void* stream);
const char* sdqlb200_tpchgen_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
