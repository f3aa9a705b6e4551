// sdqlpy-b200: device-side expansion of packed (wire-format) columns into the resident columnar layout.
// C ABI in include/sdqlb200_wire.h; host encoders in sdqlpy_b200/wire.py.  HBM-bound streaming kernels: every thread
// expands 4 consecutive elements per iteration (one 32/64/128-bit packed load, 128/256 bits of stores), grid-stride
// over a persistent grid of 8 CTAs per SM.  Dictionary tables (<= 512 KB) are read through the L1/L2 caches.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "sdqlb200.h"
#include "sdqlb200_wire.h"

namespace {

thread_local char g_err[256];

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

constexpr int kBlock = 256;

__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p + 2), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void st4(int* p, int a, int b, int c, int d) {
    asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ unsigned ld_u32(const void* p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_u64(const void* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_i128(const void* p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// CODE = uint8_t | uint16_t, T = double | int
template <class CODE, class T>
__global__ void __launch_bounds__(kBlock) k_dict(const CODE* __restrict__ src, T* __restrict__ dst, long long n,
                                                const T* __restrict__ table) {
    const long long ngrp = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += stride) {
        unsigned c0, c1, c2, c3;
        if (sizeof(CODE) == 1) {
            const unsigned w = ld_u32(src + (g << 2));
            c0 = w & 0xff; c1 = (w >> 8) & 0xff; c2 = (w >> 16) & 0xff; c3 = w >> 24;
        } else {
            const uint2 w = ld_u64(src + (g << 2));
            c0 = w.x & 0xffff; c1 = w.x >> 16; c2 = w.y & 0xffff; c3 = w.y >> 16;
        }
        st4(dst + (g << 2), __ldg(table + c0), __ldg(table + c1), __ldg(table + c2), __ldg(table + c3));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail (< 4 elements)
        const long long i = (ngrp << 2) + threadIdx.x;
        dst[i] = __ldg(table + (unsigned)src[i]);
    }
}

__global__ void __launch_bounds__(kBlock) k_fixed32(const int* __restrict__ src, double* __restrict__ dst, long long n,
                                                   double scale) {
    const long long ngrp = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += stride) {
        const int4 v = ld_i128(src + (g << 2));
        // IEEE-754 double division (nvcc default -prec-div=true): the correctly rounded quotient, i.e. exactly the
        // double a decimal parser produces for "v / scale" -- the encoder verifies this per element
        st4(dst + (g << 2), (double)v.x / scale, (double)v.y / scale, (double)v.z / scale, (double)v.w / scale);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (ngrp << 2) + threadIdx.x;
        dst[i] = (double)src[i] / scale;
    }
}

// bit-packed columns: MODE 0 = table[code], 1 = (double)(base + code) / scale, 2 = (int)(base + code), 3 = (uint8)code
template <class OUT, int MODE>
__global__ void __launch_bounds__(kBlock) k_bits(const unsigned* __restrict__ src, OUT* __restrict__ dst, long long n,
                                                 int nbits, const OUT* __restrict__ table, long long base, double scale) {
    const unsigned mask = nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u);
    auto code_at = [&](long long i) -> unsigned {
        const long long bit = i * nbits;
        const long long w = bit >> 5;
        const unsigned lo = __ldg(src + w), hi = __ldg(src + w + 1);  // neighbours share words: L1-cached loads
        return __funnelshift_r(lo, hi, (unsigned)(bit & 31)) & mask;
    };
    auto value = [&](unsigned c) -> OUT {
        if (MODE == 0) return __ldg(table + c);
        if (MODE == 1) return (OUT)((double)(base + (long long)c) / scale);
        if (MODE == 2) return (OUT)(base + (long long)c);
        return (OUT)c;
    };
    const long long ngrp = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngrp; g += stride) {
        const long long i = g << 2;
        const OUT a = value(code_at(i)), b = value(code_at(i + 1)), c = value(code_at(i + 2)), d = value(code_at(i + 3));
        if constexpr (sizeof(OUT) == 1) {
            const unsigned w = (unsigned)a | ((unsigned)b << 8) | ((unsigned)c << 16) | ((unsigned)d << 24);
            asm volatile("st.global.u32 [%0], %1;" ::"l"(dst + i), "r"(w) : "memory");
        } else {
            st4(dst + i, a, b, c, d);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (ngrp << 2) + threadIdx.x;
        dst[i] = value(code_at(i));
    }
}

int grid_for(long long groups) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms < 1) sms = 148;
    }
    long long g = (groups + kBlock - 1) / kBlock;
    const long long mx = (long long)sms * 8;
    if (g < 1) g = 1;
    return (int)(g < mx ? g : mx);
}

}  // namespace

extern "C" {

int32_t sdqlb200_wire_src_width(int32_t kind) {
    switch (kind) {
        case SDQLB200_WIRE_DICT8_F64: case SDQLB200_WIRE_DICT8_I32: return 1;
        case SDQLB200_WIRE_DICT16_F64: case SDQLB200_WIRE_DICT16_I32: return 2;
        case SDQLB200_WIRE_FIXED32_F64: return 4;
    }
    return 0;  // unknown, or bit-packed (nbits / 8 bytes per element)
}

int32_t sdqlb200_wire_dst_width(int32_t kind) {
    switch (kind) {
        case SDQLB200_WIRE_DICT8_F64: case SDQLB200_WIRE_DICT16_F64: case SDQLB200_WIRE_FIXED32_F64: return 8;
        case SDQLB200_WIRE_DICT8_I32: case SDQLB200_WIRE_DICT16_I32: return 4;
        case SDQLB200_WIRE_BITS_DICT_F64: case SDQLB200_WIRE_BITS_FIXED_F64: return 8;
        case SDQLB200_WIRE_BITS_DICT_I32: case SDQLB200_WIRE_BITS_I32: return 4;
        case SDQLB200_WIRE_BITS_U8: return 1;
    }
    return 0;
}

const char* sdqlb200_wire_last_error(void) { return g_err; }

int sdqlb200_wire_decode(int32_t kind, const void* src, void* dst, int64_t rows, const void* table, double scale,
                         void* stream) {
    if (rows < 0 || (rows > 0 && (!src || !dst))) return fail(SDQLB200_E_ARG, "wire_decode: null buffer");
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return fail(SDQLB200_E_ARG, "wire_decode: buffers must be 16-byte aligned");
    const bool dict = kind != SDQLB200_WIRE_FIXED32_F64;
    if (kind < 0 || kind >= SDQLB200_WIRE_KINDS) return fail(SDQLB200_E_ARG, "wire_decode: unknown kind %d", kind);
    if (dict && !table) return fail(SDQLB200_E_ARG, "wire_decode: dictionary kinds need a device table");
    if (!dict && !(scale > 0)) return fail(SDQLB200_E_ARG, "wire_decode: fixed-point scale must be positive");
    if (rows == 0) return SDQLB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for((rows + 3) / 4);
    switch (kind) {
        case SDQLB200_WIRE_DICT8_F64:
            k_dict<uint8_t, double><<<grid, kBlock, 0, st>>>((const uint8_t*)src, (double*)dst, rows, (const double*)table);
            break;
        case SDQLB200_WIRE_DICT16_F64:
            k_dict<uint16_t, double><<<grid, kBlock, 0, st>>>((const uint16_t*)src, (double*)dst, rows, (const double*)table);
            break;
        case SDQLB200_WIRE_DICT8_I32:
            k_dict<uint8_t, int><<<grid, kBlock, 0, st>>>((const uint8_t*)src, (int*)dst, rows, (const int*)table);
            break;
        case SDQLB200_WIRE_DICT16_I32:
            k_dict<uint16_t, int><<<grid, kBlock, 0, st>>>((const uint16_t*)src, (int*)dst, rows, (const int*)table);
            break;
        case SDQLB200_WIRE_FIXED32_F64:
            k_fixed32<<<grid, kBlock, 0, st>>>((const int*)src, (double*)dst, rows, scale);
            break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SDQLB200_E_CUDA, "%s:%d: wire_decode launch: %s", __FILE__, __LINE__, cudaGetErrorString(e));
    return SDQLB200_OK;
}

int sdqlb200_wire_decode_bits(int32_t kind, const void* src, void* dst, int64_t rows, int32_t nbits, const void* table,
                              int64_t base, double scale, void* stream) {
    if (kind < SDQLB200_WIRE_BITS_DICT_F64 || kind >= SDQLB200_WIRE_ALL_KINDS)
        return fail(SDQLB200_E_ARG, "wire_decode_bits: kind %d is not bit-packed", kind);
    if (nbits < 1 || nbits > 32) return fail(SDQLB200_E_ARG, "wire_decode_bits: nbits %d outside 1..32", nbits);
    if (rows < 0 || (rows > 0 && (!src || !dst))) return fail(SDQLB200_E_ARG, "wire_decode_bits: null buffer");
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return fail(SDQLB200_E_ARG, "wire_decode_bits: buffers must be 16-byte aligned");
    const bool dict = kind == SDQLB200_WIRE_BITS_DICT_F64 || kind == SDQLB200_WIRE_BITS_DICT_I32;
    if (dict && !table) return fail(SDQLB200_E_ARG, "wire_decode_bits: dictionary kinds need a device table");
    if (kind == SDQLB200_WIRE_BITS_FIXED_F64 && !(scale > 0)) return fail(SDQLB200_E_ARG, "wire_decode_bits: scale must be positive");
    if (kind == SDQLB200_WIRE_BITS_U8 && nbits > 8) return fail(SDQLB200_E_ARG, "wire_decode_bits: uint8 codes have at most 8 bits");
    if (rows == 0) return SDQLB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for((rows + 3) / 4);
    const unsigned* s32 = (const unsigned*)src;
    switch (kind) {
        case SDQLB200_WIRE_BITS_DICT_F64:
            k_bits<double, 0><<<grid, kBlock, 0, st>>>(s32, (double*)dst, rows, nbits, (const double*)table, 0, 1.0);
            break;
        case SDQLB200_WIRE_BITS_DICT_I32:
            k_bits<int, 0><<<grid, kBlock, 0, st>>>(s32, (int*)dst, rows, nbits, (const int*)table, 0, 1.0);
            break;
        case SDQLB200_WIRE_BITS_FIXED_F64:
            k_bits<double, 1><<<grid, kBlock, 0, st>>>(s32, (double*)dst, rows, nbits, nullptr, base, scale);
            break;
        case SDQLB200_WIRE_BITS_I32:
            k_bits<int, 2><<<grid, kBlock, 0, st>>>(s32, (int*)dst, rows, nbits, nullptr, base, 1.0);
            break;
        case SDQLB200_WIRE_BITS_U8:
            k_bits<unsigned char, 3><<<grid, kBlock, 0, st>>>(s32, (unsigned char*)dst, rows, nbits, nullptr, 0, 1.0);
            break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SDQLB200_E_CUDA, "%s:%d: wire_decode_bits launch: %s", __FILE__, __LINE__, cudaGetErrorString(e));
    return SDQLB200_OK;
}

}  // extern "C"
