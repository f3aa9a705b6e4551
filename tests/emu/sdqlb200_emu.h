// TEST-ONLY single-thread host emulation of the handful of CUDA constructs that generated query modules use.
// Built by tests/emu/build_emu.py with -DSDQLB200_EMU so the GPU-less dev container can check the *logic* of
// generated kernels (key packing, tables, FD-minimised group keys, result materialisation) against the oracle.
// It runs every kernel as <<<1,1>>>; it is never compiled into, shipped with, or loaded by the package.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __shared__ static thread_local  // per host thread: the N-rank engine tests run one emulated "GPU" per thread

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static emu_dim3 threadIdx_emu0;
static const emu_dim3 blockDim, gridDim;
struct emu_idx { unsigned x = 0, y = 0, z = 0; };
static const emu_idx threadIdx, blockIdx;

typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
static inline const char* cudaGetErrorString(int) { return "emu"; }
static inline int cudaGetLastError() { return 0; }
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline int cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
#define cudaMemcpyHostToDevice 1
#define cudaMemcpyDeviceToHost 2
#define cudaMemcpyDeviceToDevice 3

static inline void __syncthreads() {}
static inline void __threadfence() {}
static inline void __syncwarp() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { return s >= 32 ? hi : (s ? (lo >> s) | (hi << (32 - s)) : lo); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
using std::max;
using std::min;

static thread_local unsigned long long emu_dyn_smem[1 << 16];
#define SDQL_EXTERN_SMEM(name) unsigned long long* name = emu_dyn_smem
#define SDQL_LAUNCH(kernel, grid, block, smem, stream, ...) kernel(__VA_ARGS__)
#define SDQL_UNUSED(x) (void)(x)

// host-driver stubs
typedef int cudaEvent_t;
static inline cudaEvent_t sdqlhost_ev(int) { return 0; }
static inline cudaEvent_t sdqlhost_kev(int) { return 0; }
static inline int cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline int cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
static inline int sdqlhost_sms() { return 1; }
static inline int sdqlhost_occupancy(const void*, size_t) { return 1; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
template <class F> static inline int cudaFuncSetAttribute(F, int, int) { return 0; }
