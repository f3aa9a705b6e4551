// sdqlpy-b200: TPC-H fact tables generated on the device (C ABI: include/sdqlb200_tpchgen.h).
// Integer-for-integer the arithmetic of sdqlpy_b200/tpch/gen.py (mix / rint / days_to_ymd / retail_cents and the
// per-column rules of _gen_lineitem / _gen_orders); fp64 values are produced by the same single IEEE operations
// (int -> double conversion, one division by 100.0; products for o_totalprice are not fused: build with -fmad=false),
// so device and host columns are bit-identical.  One thread per order: it derives the order's lines from the global
// row id of its first line (exclusive prefix sum of lines-per-order, computed by the caller).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "sdqlb200.h"
#include "sdqlb200_tpchgen.h"

namespace {

typedef unsigned long long u64;
typedef long long i64;

thread_local char g_err[256];
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

// stream ids: gen.py, range(1, 37) in declaration order
enum { S_ORD_NL = 1, S_ORD_CUST, S_ORD_DATE, S_ORD_PRIO, S_ORD_CMT, S_ORD_CMTK, S_L_PART, S_L_SUPPJ, S_L_QTY, S_L_DISC,
       S_L_TAX, S_L_SHIP, S_L_COMMIT, S_L_RECEIPT, S_L_RFLAG, S_L_INSTR, S_L_MODE };
constexpr i64 DAY_1992_01_01 = 8035, DATE_SPAN = 2406, DAY_1995_06_17 = 9298;

__host__ __device__ __forceinline__ u64 mix(int stream, u64 idx, u64 seed) {  // gen.mix
    u64 x = idx + (seed * 0x9E3779B97F4A7C15ull + (u64)stream * 0xD1B54A32D192ED03ull);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ i64 rint_(int stream, u64 idx, i64 lo, i64 hi, u64 seed) {  // gen.rint
    return lo + (i64)((mix(stream, idx, seed) >> 11) % (u64)(hi - lo + 1));
}
__host__ __device__ __forceinline__ int days_to_ymd(i64 z) {  // gen.days_to_ymd (all operands positive here)
    z += 719468;
    const i64 era = z / 146097;
    const i64 doe = z - era * 146097;
    const i64 yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    i64 y = yoe + era * 400;
    const i64 doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const i64 mp = (5 * doy + 2) / 153;
    const i64 d = doy - (153 * mp + 2) / 5 + 1;
    const i64 m = mp < 10 ? mp + 3 : mp - 9;
    if (m <= 2) y += 1;
    return (int)(y * 10000 + m * 100 + d);
}
__host__ __device__ __forceinline__ i64 retail_cents(i64 pk) { return 90000 + (pk / 10) % 20001 + 100 * (pk % 1000); }
__host__ __device__ __forceinline__ i64 order_key(i64 i) { return (i / 8) * 32 + i % 8 + 1; }

struct Line {
    i64 pk, qty, ship, receipt, od;
    double ext, disc, tax;
    int linestatus;
};
__device__ __forceinline__ void line_core(const sdqlb200_tpch_params& p, i64 oi, i64 r, Line& l) {
    const u64 seed = (u64)p.seed;
    l.pk = rint_(S_L_PART, (u64)r, 1, p.P, seed);
    l.qty = rint_(S_L_QTY, (u64)r, 1, 50, seed);
    l.ext = (double)(l.qty * retail_cents(l.pk)) / 100.0;
    l.disc = (double)rint_(S_L_DISC, (u64)r, 0, 10, seed) / 100.0;
    l.tax = (double)rint_(S_L_TAX, (u64)r, 0, 8, seed) / 100.0;
    l.od = DAY_1992_01_01 + rint_(S_ORD_DATE, (u64)oi, 0, DATE_SPAN - 1, seed);
    l.ship = l.od + rint_(S_L_SHIP, (u64)r, 1, 121, seed);
    l.receipt = l.ship + rint_(S_L_RECEIPT, (u64)r, 1, 30, seed);
    l.linestatus = l.ship > DAY_1995_06_17 ? 0 : 1;
}

__global__ void k_order_lines(sdqlb200_tpch_params p, i64 o0, i64 o1, int* nl) {
    for (i64 i = o0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < o1; i += (i64)gridDim.x * blockDim.x)
        nl[i - o0] = (int)rint_(S_ORD_NL, (u64)i, 1, 7, (u64)p.seed);
}

__global__ void k_lineitem(sdqlb200_tpch_params p, i64 o0, i64 o1, const i64* off, sdqlb200_lineitem_cols c) {
    const u64 seed = (u64)p.seed;
    const i64 base = off[0];
    for (i64 i = o0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < o1; i += (i64)gridDim.x * blockDim.x) {
        const i64 r0 = off[i - o0], r1 = off[i - o0 + 1];
        const int okey = (int)order_key(i);
        for (i64 r = r0; r < r1; ++r) {
            const i64 w = r - base;
            Line l;
            line_core(p, i, r, l);
            if (c.l_orderkey) c.l_orderkey[w] = okey;
            if (c.l_linenumber) c.l_linenumber[w] = (int)(r - r0 + 1);
            if (c.l_partkey) c.l_partkey[w] = (int)l.pk;
            if (c.l_suppkey) {
                const i64 j = rint_(S_L_SUPPJ, (u64)r, 0, 3, seed);
                c.l_suppkey[w] = (int)((l.pk + j * (p.S / 4 + (l.pk - 1) / p.S)) % p.S + 1);
            }
            if (c.l_quantity) c.l_quantity[w] = (double)l.qty;
            if (c.l_extendedprice) c.l_extendedprice[w] = l.ext;
            if (c.l_discount) c.l_discount[w] = l.disc;
            if (c.l_tax) c.l_tax[w] = l.tax;
            if (c.l_shipdate) c.l_shipdate[w] = days_to_ymd(l.ship);
            if (c.l_commitdate) c.l_commitdate[w] = days_to_ymd(l.od + rint_(S_L_COMMIT, (u64)r, 30, 90, seed));
            if (c.l_receiptdate) c.l_receiptdate[w] = days_to_ymd(l.receipt);
            if (c.l_returnflag)
                c.l_returnflag[w] = l.receipt <= DAY_1995_06_17 ? (uint8_t)(mix(S_L_RFLAG, (u64)r, seed) & 1ull) : (uint8_t)2;
            if (c.l_linestatus) c.l_linestatus[w] = (uint8_t)l.linestatus;
            if (c.l_shipinstruct) c.l_shipinstruct[w] = (uint8_t)rint_(S_L_INSTR, (u64)r, 0, 3, seed);
            if (c.l_shipmode) c.l_shipmode[w] = (uint8_t)rint_(S_L_MODE, (u64)r, 0, 6, seed);
        }
    }
}

__device__ __forceinline__ void put(uint8_t* row, int start, const char* w) {
    for (int k = 0; w[k]; ++k) row[start + k] = (uint8_t)w[k];
}

__global__ void k_orders(sdqlb200_tpch_params p, i64 o0, i64 o1, const i64* off, sdqlb200_orders_cols c,
                         const uint8_t* vocab, int nwords) {
    const u64 seed = (u64)p.seed;
    for (i64 i = o0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < o1; i += (i64)gridDim.x * blockDim.x) {
        const i64 w = i - o0;
        if (c.o_orderkey) c.o_orderkey[w] = (int)order_key(i);
        if (c.o_custkey) {
            const i64 m = p.C - p.C / 3;
            const i64 k = rint_(S_ORD_CUST, (u64)i, 0, m - 1, seed);
            c.o_custkey[w] = (int)(k + k / 2 + 1);
        }
        if (c.o_orderdate) c.o_orderdate[w] = days_to_ymd(DAY_1992_01_01 + rint_(S_ORD_DATE, (u64)i, 0, DATE_SPAN - 1, seed));
        if (c.o_orderpriority) c.o_orderpriority[w] = (uint8_t)rint_(S_ORD_PRIO, (u64)i, 0, 4, seed);
        if (c.o_shippriority) c.o_shippriority[w] = 0;
        if (c.o_orderstatus || c.o_totalprice) {
            const i64 r0 = off[w], r1 = off[w + 1];
            i64 nf = 0, tot = 0;
            for (i64 r = r0; r < r1; ++r) {
                Line l;
                line_core(p, i, r, l);
                nf += l.linestatus == 1;
                // np.round(ext * (1 + tax) * (1 - disc) * 100): left to right, round half to even
                tot += (i64)rint(((l.ext * (1.0 + l.tax)) * (1.0 - l.disc)) * 100.0);
            }
            if (c.o_orderstatus) c.o_orderstatus[w] = (uint8_t)(nf == r1 - r0 ? 0 : (nf == 0 ? 1 : 2));
            if (c.o_totalprice) c.o_totalprice[w] = (double)tot / 100.0;
        }
        if (c.o_comment) {  // gen._text(S_ORD_CMT, i, 6, 79) + the Q13 patterns + gen._cut
            uint8_t row[79];
            for (int b = 0; b < 79; ++b) row[b] = 0;
            for (int s = 0; s < 6; ++s) {
                const u64 code = mix(S_ORD_CMT, (u64)i * 6ull + (u64)s, seed) % (u64)nwords;
                for (int b = 0; b < 12; ++b) row[s * 12 + b] = vocab[code * 12 + b];
            }
            const i64 k = (i64)(mix(S_ORD_CMTK, (u64)i + 7777777ull, seed) % 100ull);
            if (k == 0) { put(row, 10, "special "); put(row, 40, "requests "); }
            else if (k == 1) put(row, 20, "special ");
            else if (k == 2) { put(row, 0, "requests "); put(row, 30, "special "); }
            else if (k == 3) put(row, 10, "specialrequests");
            i64 ln = rint_(S_ORD_CMTK, (u64)i, 19, 70, seed);
            if (k < 4 && ln < 60) ln = 60;
            uint8_t* out = c.o_comment + w * 79;
            for (int b = 0; b < 79; ++b) out[b] = b >= ln ? (uint8_t)0 : row[b];
        }
    }
}

int grid_for(i64 items) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms < 1) sms = 148;
    }
    i64 g = (items + 255) / 256;
    const i64 mx = (i64)sms * 8;
    if (g < 1) g = 1;
    return (int)(g < mx ? g : mx);
}

int check(const sdqlb200_tpch_params* p, i64 o0, i64 o1) {
    if (!p) return fail(SDQLB200_E_ARG, "tpchgen: null parameters");
    if (p->S < 4 || p->P < 1 || p->C < 3 || p->O < 1) return fail(SDQLB200_E_ARG, "tpchgen: bad table sizes");
    if (o0 < 0 || o1 < o0 || o1 > p->O) return fail(SDQLB200_E_ARG, "tpchgen: order range [%lld, %lld) outside [0, %lld)", o0, o1, (i64)p->O);
    return SDQLB200_OK;
}

int launched(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SDQLB200_E_CUDA, "%s:%d: %s launch: %s", __FILE__, __LINE__, what, cudaGetErrorString(e));
    return SDQLB200_OK;
}

}  // namespace

extern "C" {

const char* sdqlb200_tpchgen_last_error(void) { return g_err; }

int sdqlb200_tpchgen_order_lines(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, int32_t* nl, void* stream) {
    if (int rc = check(p, o0, o1)) return rc;
    if (o1 == o0) return SDQLB200_OK;
    if (!nl) return fail(SDQLB200_E_ARG, "tpchgen: null output");
    k_order_lines<<<grid_for(o1 - o0), 256, 0, (cudaStream_t)stream>>>(*p, o0, o1, nl);
    return launched("order_lines");
}

int sdqlb200_tpchgen_lineitem(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, const int64_t* off,
                              const sdqlb200_lineitem_cols* cols, void* stream) {
    if (int rc = check(p, o0, o1)) return rc;
    if (o1 == o0) return SDQLB200_OK;
    if (!off || !cols) return fail(SDQLB200_E_ARG, "tpchgen: null argument");
    k_lineitem<<<grid_for(o1 - o0), 256, 0, (cudaStream_t)stream>>>(*p, o0, o1, (const i64*)off, *cols);
    return launched("lineitem");
}

// vocab: DEVICE table of `nwords` 12-byte, space padded words (gen.WORDS + " "), needed for o_comment only
int sdqlb200_tpchgen_orders_text(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, const int64_t* off,
                                 const sdqlb200_orders_cols* cols, const uint8_t* vocab, int32_t nwords, void* stream) {
    if (int rc = check(p, o0, o1)) return rc;
    if (o1 == o0) return SDQLB200_OK;
    if (!cols || ((cols->o_orderstatus || cols->o_totalprice) && !off)) return fail(SDQLB200_E_ARG, "tpchgen: null argument");
    if (cols->o_comment && (!vocab || nwords < 1)) return fail(SDQLB200_E_ARG, "tpchgen: o_comment needs the vocabulary table");
    k_orders<<<grid_for(o1 - o0), 256, 0, (cudaStream_t)stream>>>(*p, o0, o1, (const i64*)off, *cols, vocab, nwords);
    return launched("orders");
}

int sdqlb200_tpchgen_orders(const sdqlb200_tpch_params* p, int64_t o0, int64_t o1, const int64_t* off,
                            const sdqlb200_orders_cols* cols, void* stream) {
    if (cols && cols->o_comment) return fail(SDQLB200_E_ARG, "tpchgen: use sdqlb200_tpchgen_orders_text for o_comment");
    return sdqlb200_tpchgen_orders_text(p, o0, o1, off, cols, nullptr, 0, stream);
}

}  // extern "C"
