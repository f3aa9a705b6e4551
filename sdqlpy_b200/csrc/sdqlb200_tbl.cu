// sdqlpy-b200: device-side reader of dbgen's pipe-delimited .tbl text (C ABI in include/sdqlb200_tbl.h, host side in
// sdqlpy_b200/tbl.py).  Replaces the reference's read_csv (sdql_lib.py:69-128: Python csv.reader + per-field int() /
// float() calls).  Byte work, HBM/L2 bound:
//   k_tbl_count   every CTA counts the newlines of 4 KB tiles (one 128-bit load per thread, SWAR byte compare)
//   k_tbl_scan    one CTA turns the tile counts into exclusive prefix sums (total = newline count)
//   k_tbl_starts  the tiles are read again; a block-wide exclusive scan of the per-thread counts orders the newlines
//                 of a tile, every thread writes the offsets of the rows that start behind its newlines
//   k_tbl_parse   one thread per row walks the row's fields (the rows of a warp are ~4 KB of consecutive text: the
//                 byte loads hit L1 lines fetched once), converts them and writes the resident column layout
// SDQLB200_EMU (tests only): the same code as single-threaded host C++.
#ifndef SDQLB200_EMU
#include <cuda_runtime.h>
#else
#include "sdqlb200_emu.h"
#endif

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "sdqlb200.h"
#include "sdqlb200_tbl.h"

namespace {

thread_local char g_err[256];

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define TBL_CUDA(x)                                                                                                   \
    do {                                                                                                              \
        cudaError_t e_ = (x);                                                                                         \
        if (e_ != cudaSuccess) return fail(SDQLB200_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)

typedef long long i64;
typedef unsigned long long u64;
constexpr int kBlock = 256;                       // threads per CTA; one 16-byte chunk per thread and tile
constexpr int kTile = SDQLB200_TBL_TILE;          // bytes per tile = kBlock * 16
constexpr int kChunks = kTile / 16;
static_assert(kChunks == kBlock, "one chunk per thread");

#ifndef SDQLB200_EMU
#define TBL_DEV __device__ __forceinline__
#define TBL_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#else
#define TBL_DEV static inline
#define TBL_LAUNCH(kernel, grid, block, stream, ...) kernel(__VA_ARGS__)
#endif

// the 16 bytes at text + off as four little-endian words; bytes at or behind `bytes` read as zero
TBL_DEV void load16(const unsigned char* text, i64 off, i64 bytes, unsigned (&w)[4]) {
    if (off + 16 <= bytes) {
#ifndef SDQLB200_EMU
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(text + off));
#else
        memcpy(w, text + off, 16);
#endif
    } else {
        w[0] = w[1] = w[2] = w[3] = 0u;
        for (int j = 0; j < 16; ++j)
            if (off + j < bytes) w[j >> 2] |= (unsigned)text[off + j] << (8 * (j & 3));
    }
}
// 0x80 in every byte of x that equals '\n' (exact per byte)
TBL_DEV unsigned newline_bytes(unsigned x) {
    x ^= 0x0a0a0a0au;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}
TBL_DEV int popc(unsigned v) {
#ifndef SDQLB200_EMU
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

#ifndef SDQLB200_EMU
// all threads of the CTA call; returns the sum of v over the threads with a lower index, *total = sum over all
TBL_DEV int block_excl_scan(int v, int* total) {
    __shared__ int wsum[kBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // wsum of a previous call has been read
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; ++k) {
        const int s = wsum[k];
        if (k < wid) base += s;
        tot += s;
    }
    *total = tot;
    return base + inc - v;
}
TBL_DEV i64 block_excl_scan64(i64 v, i64* total) {
    __shared__ i64 wsum64[kBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    i64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const i64 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) wsum64[wid] = inc;
    __syncthreads();
    i64 base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; ++k) {
        const i64 s = wsum64[k];
        if (k < wid) base += s;
        tot += s;
    }
    *total = tot;
    return base + inc - v;
}
#else
TBL_DEV int block_excl_scan(int v, int* total) { *total = v; return 0; }
TBL_DEV i64 block_excl_scan64(i64 v, i64* total) { *total = v; return 0; }
#endif

// tile_counts[t] = newlines in text[t * kTile, (t + 1) * kTile)
__global__ void k_tbl_count(const unsigned char* text, i64 bytes, i64 ntiles, int* tile_counts) {
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int cnt = 0;
        for (int s = threadIdx.x; s < kChunks; s += blockDim.x) {  // one iteration on the GPU, all chunks in the emulation
            unsigned w[4];
            load16(text, t * kTile + (i64)s * 16, bytes, w);
            cnt += popc(newline_bytes(w[0])) + popc(newline_bytes(w[1])) + popc(newline_bytes(w[2])) + popc(newline_bytes(w[3]));
        }
        int total;
        block_excl_scan(cnt, &total);
        if (threadIdx.x == 0) tile_counts[t] = total;
    }
}

// one CTA: tile_offsets[t] = sum of tile_counts[0 .. t), *newlines = sum of all
__global__ void k_tbl_scan(const int* tile_counts, i64 ntiles, i64* tile_offsets, i64* newlines) {
    const i64 per = (ntiles + blockDim.x - 1) / blockDim.x;
    const i64 lo = (i64)threadIdx.x * per, hi = lo + per < ntiles ? lo + per : ntiles;
    i64 sum = 0;
    for (i64 t = lo; t < hi; ++t) sum += tile_counts[t];
    i64 total;
    i64 run = block_excl_scan64(sum, &total);
    for (i64 t = lo; t < hi; ++t) { tile_offsets[t] = run; run += tile_counts[t]; }
    if (threadIdx.x == 0) *newlines = total;
}

// starts[0] = 0; starts[i + 1] = offset behind the i-th newline; starts[rows] = end of the last row + 1
__global__ void k_tbl_starts(const unsigned char* text, i64 bytes, i64 ntiles, const i64* tile_offsets, i64* starts, i64 rows) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        starts[0] = 0;
        if (bytes > 0 && text[bytes - 1] != '\n') starts[rows] = bytes + 1;  // unterminated last row: a virtual newline at `bytes`
    }
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const i64 base = tile_offsets[t];
        int run = 0;  // emulation: newlines of this thread's earlier chunks of the tile
        for (int s = threadIdx.x; s < kChunks; s += blockDim.x) {
            const i64 off = t * kTile + (i64)s * 16;
            unsigned w[4];
            load16(text, off, bytes, w);
            unsigned m[4];
            int cnt = 0;
            for (int j = 0; j < 4; ++j) { m[j] = newline_bytes(w[j]); cnt += popc(m[j]); }
            int total;
            i64 pos = base + run + block_excl_scan(cnt, &total);
            if (cnt) {
                for (int j = 0; j < 16; ++j)
                    if ((m[j >> 2] >> (8 * (j & 3) + 7)) & 1u) starts[++pos] = off + j + 1;
            }
            run += cnt;
        }
    }
}

struct ParseCols {
    int n;
    sdqlb200_tbl_col c[SDQLB200_TBL_MAX_COLS];
};

TBL_DEV void status_bad(sdqlb200_tbl_status* st, i64 row, int err) {
    const u64 old = atomicMin((u64*)&st->bad_row, (u64)row);
    if ((u64)row < old) st->error = err;  // racy between different bad rows (documented: "of one malformed row")
}

// powers of ten that are exact doubles
__device__ const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                      1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__global__ void k_tbl_parse(const unsigned char* text, const i64* starts, i64 rows, ParseCols pc, unsigned char delim,
                            sdqlb200_tbl_status* st) {
    i64 mn[SDQLB200_TBL_MAX_COLS], mx[SDQLB200_TBL_MAX_COLS];
    bool any = false;
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (i64)gridDim.x * blockDim.x) {
        i64 p = starts[r], e = starts[r + 1] - 1;  // the row is text[p, e)
        if (e > p && text[e - 1] == '\r') --e;
        if (!any) {
            any = true;
            for (int c = 0; c < pc.n; ++c) { mn[c] = 0x7fffffffffffffffll; mx[c] = -0x7fffffffffffffffll - 1; }
        }
        for (int c = 0; c < pc.n; ++c) {
            if (p > e) { status_bad(st, r, SDQLB200_TBL_E_FIELDS); break; }
            i64 f = p;  // the field is text[p, f)
            const sdqlb200_tbl_col col = pc.c[c];
            if (!col.out) {
                while (f < e && text[f] != delim) ++f;
            } else if (col.type == SDQLB200_TBL_STR) {
                unsigned char* out = (unsigned char*)col.out + r * (i64)col.width;
                int k = 0;
                bool hi = false;
                while (f < e && text[f] != delim) {
                    const unsigned char ch = text[f++];
                    hi |= ch >= 0x80;
                    if (k < col.width) out[k++] = ch;
                }
                while (k < col.width) out[k++] = 0;
                if (hi) status_bad(st, r, SDQLB200_TBL_E_CHAR);
            } else {
                bool neg = false, bad = false, point = false;
                int digits = 0, frac = 0, sig = 0;
                u64 m = 0;
                if (f < e && (text[f] == '-' || text[f] == '+') && col.type != SDQLB200_TBL_DATE) { neg = text[f] == '-'; ++f; }
                while (f < e && text[f] != delim) {
                    const unsigned char ch = text[f++];
                    if (ch >= '0' && ch <= '9') {
                        ++digits;
                        if (m || ch != '0') ++sig;
                        if (sig <= 18) m = m * 10 + (ch - '0'); else bad = true;
                        if (point) ++frac;
                    } else if (ch == '-' && col.type == SDQLB200_TBL_DATE && digits) {
                        // int(v.replace("-", "")): separators are dropped
                    } else if (ch == '.' && col.type == SDQLB200_TBL_FLOAT && !point) {
                        point = true;
                    } else {
                        bad = true;
                    }
                }
                if (!digits || bad) {
                    status_bad(st, r, (bad && sig > 18) ? SDQLB200_TBL_E_RANGE : SDQLB200_TBL_E_NUMBER);
                } else if (col.type == SDQLB200_TBL_FLOAT) {
                    if (sig > 15 || frac > 22) status_bad(st, r, SDQLB200_TBL_E_RANGE);
                    const double v = (double)m / kPow10[frac > 22 ? 22 : frac];  // one correctly rounded division
                    ((double*)col.out)[r] = neg ? -v : v;
                } else {
                    const i64 v = neg ? -(i64)m : (i64)m;
                    if (v < -2147483648ll || v > 2147483647ll) status_bad(st, r, SDQLB200_TBL_E_RANGE);
                    ((int*)col.out)[r] = (int)v;
                    if (v < mn[c]) mn[c] = v;
                    if (v > mx[c]) mx[c] = v;
                }
            }
            p = f + 1;  // behind the delimiter (or behind the row's end: p == e + 1 after the last field)
        }
    }
    if (any)
        for (int c = 0; c < pc.n; ++c)
            if (pc.c[c].out && (pc.c[c].type == SDQLB200_TBL_INT || pc.c[c].type == SDQLB200_TBL_DATE) && mn[c] <= mx[c]) {
                atomicMin((i64*)&st->min[c], mn[c]);
                atomicMax((i64*)&st->max[c], mx[c]);
            }
}

int sms() {
#ifndef SDQLB200_EMU
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n < 1) n = 148; }
    return n;
#else
    return 1;
#endif
}
int grid_for(i64 items, int per_sm) {
    const i64 mx = (i64)sms() * per_sm;
    if (items < 1) items = 1;
    return (int)(items < mx ? items : mx);
}
i64 tiles_of(i64 bytes) { return (bytes + kTile - 1) / kTile; }
i64 up256(i64 b) { return (b + 255) & ~255ll; }

}  // namespace

extern "C" {

const char* sdqlb200_tbl_last_error(void) { return g_err; }

// scratch layout: [newlines: 256 B] [tile_counts: int32 x ntiles] [tile_offsets: int64 x ntiles]
int64_t sdqlb200_tbl_scratch_bytes(int64_t bytes) {
    const i64 nt = tiles_of(bytes < 0 ? 0 : bytes);
    return 256 + up256(nt * 4 + 4) + up256(nt * 8 + 8);
}

int sdqlb200_tbl_index(const void* d_text, int64_t bytes, void* d_scratch, int64_t* rows_out, void* stream) {
    if (bytes < 0 || !d_scratch || !rows_out || (bytes > 0 && !d_text)) return fail(SDQLB200_E_ARG, "tbl_index: bad arguments");
    if ((uintptr_t)d_text & 15) return fail(SDQLB200_E_ARG, "tbl_index: the text must be 16-byte aligned");
    const i64 nt = tiles_of(bytes);
    i64* d_newlines = (i64*)d_scratch;
    int* d_counts = (int*)((char*)d_scratch + 256);
    i64* d_offsets = (i64*)((char*)d_scratch + 256 + up256(nt * 4 + 4));
    TBL_CUDA(cudaMemsetAsync(d_newlines, 0, 8, (cudaStream_t)stream));
    if (nt > 0) {
        TBL_LAUNCH(k_tbl_count, grid_for(nt, 8), kBlock, stream, (const unsigned char*)d_text, (i64)bytes, nt, d_counts);
        TBL_LAUNCH(k_tbl_scan, 1, kBlock, stream, (const int*)d_counts, nt, d_offsets, d_newlines);
        TBL_CUDA(cudaGetLastError());
    }
    i64 newlines = 0;
    unsigned char last = '\n';
    TBL_CUDA(cudaMemcpyAsync(&newlines, d_newlines, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (bytes > 0) TBL_CUDA(cudaMemcpyAsync(&last, (const char*)d_text + bytes - 1, 1, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TBL_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *rows_out = newlines + ((bytes > 0 && last != '\n') ? 1 : 0);
    return SDQLB200_OK;
}

int sdqlb200_tbl_row_starts(const void* d_text, int64_t bytes, const void* d_scratch, int64_t* d_starts, int64_t rows,
                            void* stream) {
    if (bytes < 0 || rows < 0 || !d_scratch || !d_starts || (bytes > 0 && !d_text)) return fail(SDQLB200_E_ARG, "tbl_row_starts: bad arguments");
    const i64 nt = tiles_of(bytes);
    const i64* d_offsets = (const i64*)((const char*)d_scratch + 256 + up256(nt * 4 + 4));
    TBL_LAUNCH(k_tbl_starts, grid_for(nt, 8), kBlock, stream, (const unsigned char*)d_text, (i64)bytes, nt, d_offsets, (i64*)d_starts, (i64)rows);
    TBL_CUDA(cudaGetLastError());
    return SDQLB200_OK;
}

int sdqlb200_tbl_parse(const void* d_text, const int64_t* d_starts, int64_t rows, const sdqlb200_tbl_col* cols,
                       int32_t ncols, char delimiter, sdqlb200_tbl_status* d_status, void* stream) {
    if (rows < 0 || !d_starts || !cols || ncols < 1 || ncols > SDQLB200_TBL_MAX_COLS || !d_status || (rows > 0 && !d_text))
        return fail(SDQLB200_E_ARG, "tbl_parse: bad arguments");
    ParseCols pc;
    memset(&pc, 0, sizeof pc);
    pc.n = ncols;
    for (int c = 0; c < ncols; ++c) {
        if (cols[c].type < SDQLB200_TBL_INT || cols[c].type > SDQLB200_TBL_STR) return fail(SDQLB200_E_ARG, "tbl_parse: column %d: unknown type %d", c, cols[c].type);
        if (cols[c].out && cols[c].type == SDQLB200_TBL_STR && cols[c].width < 1) return fail(SDQLB200_E_ARG, "tbl_parse: column %d: string width %d", c, cols[c].width);
        pc.c[c] = cols[c];
    }
    sdqlb200_tbl_status init;
    init.bad_row = -1;  // == UINT64_MAX for the atomicMin
    init.error = 0;
    for (int c = 0; c < SDQLB200_TBL_MAX_COLS; ++c) { init.min[c] = 0x7fffffffffffffffll; init.max[c] = -0x7fffffffffffffffll - 1; }
    static thread_local sdqlb200_tbl_status h_init;  // outlives the asynchronous copy
    h_init = init;
    TBL_CUDA(cudaMemcpyAsync(d_status, &h_init, sizeof h_init, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (rows > 0) {
        TBL_LAUNCH(k_tbl_parse, grid_for((rows + kBlock - 1) / kBlock, 8), kBlock, stream, (const unsigned char*)d_text,
                   (const i64*)d_starts, (i64)rows, pc, (unsigned char)delimiter, d_status);
        TBL_CUDA(cudaGetLastError());
    }
    return SDQLB200_OK;
}

}  // extern "C"
