// libsdqlb200_ingest.so -- reference-layout columns (int64 / UCS4, as sdqlpy's read_csv makes them, sdql_lib.py:83-97)
// converted to the resident layout ON THE DEVICE (include/sdqlb200_ingest.h).  Replaces the PyArray_DATA casts of
// sdql_compiler.py:644-668 plus what used to be numpy work on the host (astype(int32), np.unique).
// HBM-bound streaming kernels: 128-bit loads, persistent grid of 8 CTAs per SM.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <thread>
#include <vector>

#include "sdqlb200.h"
#include "sdqlb200_ingest.h"

namespace {

thread_local char g_err[256];
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define ING_CUDA(x)                                                                                                  \
    do {                                                                                                             \
        cudaError_t e_ = (x);                                                                                        \
        if (e_ != cudaSuccess) return fail(SDQLB200_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)

constexpr int kBlock = 256;
typedef unsigned long long u64;
typedef long long i64;

int grid_for(i64 items) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms < 1) sms = 148; }
    i64 g = (items + kBlock - 1) / kBlock, mx = (i64)sms * 8;
    return (int)(g < 1 ? 1 : g < mx ? g : mx);
}

__device__ __forceinline__ u64 hash64(u64 x) { x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x; }

// 4 values per thread per iteration: two 128-bit loads, one 128-bit store
__global__ void __launch_bounds__(kBlock) k_i64(const i64* __restrict__ in, int* __restrict__ out, i64 n, i64* minmax) {
    i64 mn = 0x7fffffffffffffffll, mx = -0x7fffffffffffffffll - 1;
    const i64 ngrp = n >> 2, stride = (i64)gridDim.x * blockDim.x;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += stride) {
        longlong2 a, b;
        asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(in + 4 * g));
        asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(in + 4 * g + 2));
        mn = min(min(mn, a.x), min(a.y, min(b.x, b.y)));
        mx = max(max(mx, a.x), max(a.y, max(b.x, b.y)));
        asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * g), "r"((int)a.x), "r"((int)a.y), "r"((int)b.x), "r"((int)b.y) : "memory");
    }
    if (blockIdx.x == 0)
        for (i64 i = (ngrp << 2) + threadIdx.x; i < n; i += blockDim.x) { const i64 v = in[i]; out[i] = (int)v; mn = min(mn, v); mx = max(mx, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(minmax, mn); atomicMax(minmax + 1, mx); }
}

// one thread per output 32-bit word (4 characters); rows are `nchar` code points in, `width` bytes out
__global__ void __launch_bounds__(kBlock) k_ucs4_bytes(const unsigned* __restrict__ in, unsigned char* __restrict__ out, i64 rows, int nchar,
                                                       int width, u64* bad) {
    const i64 total = rows * width, stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const i64 r = e / width;
        const int c = (int)(e - r * width);
        unsigned v = c < nchar ? __ldg(in + r * nchar + c) : 0u;
        if (v > 255u) { atomicMin(bad, (u64)r + 1); v = 0; }
        out[e] = (unsigned char)v;
    }
}

__device__ __forceinline__ u64 row_hash(const unsigned* s, int nchar) {
    u64 h = 0x9e3779b97f4a7c15ull;
    for (int k = 0; k < nchar; ++k) {
        const unsigned v = __ldg(s + k);
        if (!v) break;  // numpy pads with NULs: the value ends here
        h = hash64(h ^ (u64)v);
    }
    return h == ~0ull ? 0ull : h;  // ~0 marks a free slot
}

__global__ void __launch_bounds__(kBlock) k_ucs4_distinct(const unsigned* __restrict__ in, i64 rows, i64 row_base, int nchar, u64* keys,
                                                          i64* rep, i64 cap, u64* count) {
    const u64 m = (u64)cap - 1;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += stride) {
        const u64 h = row_hash(in + i * nchar, nchar);
        u64 s = hash64(h) & m;
        for (int probe = 0; probe < cap; ++probe) {
            u64 k = keys[s];
            if (k == ~0ull) {
                if (*(volatile u64*)count >= (u64)cap / 2) return;  // overflow: the caller sees count >= cap / 2
                k = atomicCAS(keys + s, ~0ull, h);
                if (k == ~0ull) { atomicAdd(count, 1ull); k = h; }
            }
            // the representative only ever decreases: a plain read screens the atomic out for all but the first few rows of a
            // value (600 M rows of a 3-value column otherwise serialise on 3 addresses: 245 ms per 2.4 GB column at SF100)
            if (k == h) { if (row_base + i < *(volatile i64*)(rep + s)) atomicMin(rep + s, row_base + i); break; }
            s = (s + 1) & m;
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_ucs4_codes(const unsigned* __restrict__ in, i64 rows, int nchar, const u64* __restrict__ keys,
                                                       const int* __restrict__ slot_code, i64 cap, const unsigned* __restrict__ dict,
                                                       void* out, int out_width, u64* bad) {
    const u64 m = (u64)cap - 1;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += stride) {
        const unsigned* s = in + i * nchar;
        const u64 h = row_hash(s, nchar);
        u64 sl = hash64(h) & m;
        int code = -1;
        for (int probe = 0; probe < cap; ++probe) {
            const u64 k = __ldg(keys + sl);
            if (k == h) { code = __ldg(slot_code + sl); break; }
            if (k == ~0ull) break;
            sl = (sl + 1) & m;
        }
        bool ok = code >= 0;
        if (ok) {  // the row must BE the dictionary entry (a 64-bit hash collision is reported, never swallowed)
            const unsigned* d = dict + (i64)code * nchar;
            for (int k = 0; k < nchar; ++k) {
                const unsigned a = __ldg(s + k), b = __ldg(d + k);
                if (a != b) { ok = false; break; }
                if (!a) break;
            }
        }
        if (!ok) { atomicMin(bad, (u64)i + 1); code = 0; }
        if (out_width == 1) ((unsigned char*)out)[i] = (unsigned char)code;
        else ((int*)out)[i] = code;
    }
}

__global__ void __launch_bounds__(kBlock) k_remap(const int* __restrict__ in, const int* __restrict__ table, void* out, int out_width, i64 n) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = __ldg(table + __ldg(in + i));
        if (out_width == 1) ((unsigned char*)out)[i] = (unsigned char)c;
        else ((int*)out)[i] = c;
    }
}

__global__ void __launch_bounds__(kBlock) k_remap_u8(const unsigned char* __restrict__ in, const unsigned char* __restrict__ table,
                                                     unsigned char* out, i64 n) {
    __shared__ unsigned char lut[256];
    lut[threadIdx.x] = table[threadIdx.x];  // kBlock == 256
    __syncthreads();
    const i64 nq = n >> 2, stride = (i64)gridDim.x * blockDim.x;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        const unsigned w = __ldg((const unsigned*)in + q);
        ((unsigned*)out)[q] = (unsigned)lut[w & 0xff] | ((unsigned)lut[(w >> 8) & 0xff] << 8) | ((unsigned)lut[(w >> 16) & 0xff] << 16) |
                              ((unsigned)lut[w >> 24] << 24);
    }
    if (blockIdx.x == 0)
        for (i64 i = (nq << 2) + threadIdx.x; i < n; i += blockDim.x) out[i] = lut[in[i]];
}

// [lo, hi) of worker w of nw over n items, in units of 64 (cache line of the narrow output)
inline void host_slice(i64 n, int w, int nw, i64* lo, i64* hi) {
    const i64 per = ((n + nw - 1) / nw + 63) & ~63ll;
    *lo = per * w < n ? per * w : n;
    *hi = per * (w + 1) < n ? per * (w + 1) : n;
}

}  // namespace

extern "C" {

int sdqlb200_ingest_remap_u8(const uint8_t* d_in, const uint8_t* d_table256, uint8_t* d_out, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!d_in || !d_out || !d_table256))) return fail(SDQLB200_E_ARG, "ingest_remap_u8: bad arguments");
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 3) return fail(SDQLB200_E_ARG, "ingest_remap_u8: pointers must be 4-byte aligned");
    if (n == 0) return 0;
    k_remap_u8<<<grid_for((n + 3) / 4), kBlock, 0, (cudaStream_t)stream>>>(d_in, d_table256, d_out, n);
    ING_CUDA(cudaGetLastError());
    return 0;
}

int sdqlb200_ingest_host_i64(const int64_t* h_in, int32_t* h_out, int64_t n, int32_t threads, int64_t* minmax) {
    if (n < 0 || (n > 0 && (!h_in || !h_out)) || !minmax) return fail(SDQLB200_E_ARG, "ingest_host_i64: bad arguments");
    const int nw = threads < 1 ? 1 : threads > 256 ? 256 : threads;
    std::vector<i64> mn(nw, 0x7fffffffffffffffll), mx(nw, -0x7fffffffffffffffll - 1);
    auto work = [&](int w) {
        i64 lo, hi;
        host_slice(n, w, nw, &lo, &hi);
        i64 a = 0x7fffffffffffffffll, b = -0x7fffffffffffffffll - 1;
        for (i64 i = lo; i < hi; ++i) {  // vectorises: 64-bit min / max + truncating store
            const i64 v = h_in[i];
            a = v < a ? v : a;
            b = v > b ? v : b;
            h_out[i] = (int32_t)v;
        }
        mn[w] = a; mx[w] = b;
    };
    std::vector<std::thread> th;
    for (int w = 1; w < nw; ++w) th.emplace_back(work, w);
    work(0);
    for (auto& t : th) t.join();
    i64 a = mn[0], b = mx[0];
    for (int w = 1; w < nw; ++w) { a = mn[w] < a ? mn[w] : a; b = mx[w] > b ? mx[w] : b; }
    minmax[0] = a; minmax[1] = b;
    return 0;
}

int sdqlb200_ingest_host_ucs4_1(const uint32_t* h_in, uint8_t* h_out, int64_t n, int32_t threads, uint64_t present[4],
                                int64_t* bad_row) {
    if (n < 0 || (n > 0 && (!h_in || !h_out)) || !present || !bad_row) return fail(SDQLB200_E_ARG, "ingest_host_ucs4_1: bad arguments");
    const int nw = threads < 1 ? 1 : threads > 256 ? 256 : threads;
    std::vector<uint64_t> pr((size_t)nw * 4, 0);
    std::vector<i64> bad(nw, -1);
    auto work = [&](int w) {
        i64 lo, hi;
        host_slice(n, w, nw, &lo, &hi);
        bool seen[256] = {false};
        unsigned big = 0;
        for (i64 i = lo; i < hi; ++i) {
            const unsigned v = h_in[i];
            big |= v;
            h_out[i] = (uint8_t)v;
            seen[v & 0xff] = true;
        }
        if (big > 255u)
            for (i64 i = lo; i < hi; ++i)
                if (h_in[i] > 255u) { bad[w] = i; break; }
        for (int b = 0; b < 256; ++b)
            if (seen[b]) pr[(size_t)w * 4 + (b >> 6)] |= 1ull << (b & 63);
    };
    std::vector<std::thread> th;
    for (int w = 1; w < nw; ++w) th.emplace_back(work, w);
    work(0);
    for (auto& t : th) t.join();
    present[0] = present[1] = present[2] = present[3] = 0;
    *bad_row = -1;
    for (int w = 0; w < nw; ++w) {
        for (int k = 0; k < 4; ++k) present[k] |= pr[(size_t)w * 4 + k];
        if (bad[w] >= 0 && (*bad_row < 0 || bad[w] < *bad_row)) *bad_row = bad[w];
    }
    return 0;
}

const char* sdqlb200_ingest_last_error(void) { return g_err; }

int sdqlb200_ingest_i64(const int64_t* d_in, int32_t* d_out, int64_t n, int64_t* d_minmax, void* stream) {
    if (n < 0 || (n > 0 && (!d_in || !d_out)) || !d_minmax) return fail(SDQLB200_E_ARG, "ingest_i64: bad arguments");
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return fail(SDQLB200_E_ARG, "ingest_i64: pointers must be 16-byte aligned");
    if (n == 0) return 0;
    k_i64<<<grid_for((n + 3) / 4), kBlock, 0, (cudaStream_t)stream>>>((const i64*)d_in, d_out, n, (i64*)d_minmax);
    ING_CUDA(cudaGetLastError());
    return 0;
}

int sdqlb200_ingest_ucs4_bytes(const uint32_t* d_in, uint8_t* d_out, int64_t rows, int32_t nchar, int32_t width,
                               unsigned long long* d_bad, void* stream) {
    if (rows < 0 || nchar < 1 || width < 1 || !d_bad) return fail(SDQLB200_E_ARG, "ingest_ucs4_bytes: bad arguments");
    if (rows == 0) return 0;
    k_ucs4_bytes<<<grid_for(rows * width), kBlock, 0, (cudaStream_t)stream>>>(d_in, d_out, rows, nchar, width, d_bad);
    ING_CUDA(cudaGetLastError());
    return 0;
}

int sdqlb200_ingest_ucs4_distinct(const uint32_t* d_in, int64_t rows, int64_t row_base, int32_t nchar,
                                  unsigned long long* d_keys, long long* d_rep, int64_t cap,
                                  unsigned long long* d_count, void* stream) {
    if (rows < 0 || nchar < 1 || cap < 2 || (cap & (cap - 1)) || !d_keys || !d_rep || !d_count)
        return fail(SDQLB200_E_ARG, "ingest_ucs4_distinct: bad arguments");
    if (rows == 0) return 0;
    k_ucs4_distinct<<<grid_for(rows), kBlock, 0, (cudaStream_t)stream>>>(d_in, rows, row_base, nchar, d_keys, d_rep, cap, d_count);
    ING_CUDA(cudaGetLastError());
    return 0;
}

int sdqlb200_ingest_ucs4_codes(const uint32_t* d_in, int64_t rows, int32_t nchar, const unsigned long long* d_keys,
                               const int32_t* d_slot_code, int64_t cap, const uint32_t* d_dict, void* d_out,
                               int32_t out_width, unsigned long long* d_bad, void* stream) {
    if (rows < 0 || nchar < 1 || cap < 2 || (cap & (cap - 1)) || (out_width != 1 && out_width != 4) || !d_bad)
        return fail(SDQLB200_E_ARG, "ingest_ucs4_codes: bad arguments");
    if (rows == 0) return 0;
    k_ucs4_codes<<<grid_for(rows), kBlock, 0, (cudaStream_t)stream>>>(d_in, rows, nchar, d_keys, d_slot_code, cap, d_dict, d_out, out_width, d_bad);
    ING_CUDA(cudaGetLastError());
    return 0;
}

int sdqlb200_ingest_remap(const int32_t* d_in, const int32_t* d_table, void* d_out, int32_t out_width, int64_t n, void* stream) {
    if (n < 0 || (out_width != 1 && out_width != 4)) return fail(SDQLB200_E_ARG, "ingest_remap: bad arguments");
    if (n == 0) return 0;
    k_remap<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(d_in, d_table, d_out, out_width, n);
    ING_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
