// sdqlpy-b200 device runtime: the hand-written sm_100a building blocks that generated query kernels are
// assembled from (replaces the reference's native runtime L5: varchar.h, tuple_helper.h, map_helper.h and the
// vendored phmap -- SURVEY.md section 2 rows 11/12).
//
//   * streaming column loads    128-bit / 256-bit ld.global.nc.L1::no_allocate, 4 rows per thread per column
//   * column tile ring          bulk-async (TMA engine) global -> shared copies of whole column tiles into a
//                               multi-stage shared-memory ring, completion tracked by mbarriers (full / empty per stage)
//   * Tbl                       one device dictionary: direct-indexed (dense key domain) or open-addressing hash
//                               (atomicCAS claim, linear probing); every slot keeps a representative source index
//   * reductions                warp-shuffle + shared-memory block reductions (fp64 / int64)
//   * fixed-width byte strings  ==, startsWith, endsWith, firstIndex/contains with the reference's VarChar
//                               semantics (varchar.h:61-135) but bounded to the row (no over-read)
//
// SDQLB200_EMU: compile the same code as plain single-threaded C++ (tests/ only: lets the GPU-less dev container
// check generated query logic; never built into or loaded by the package).
#pragma once
#include <cstdint>
#include <cstring>

#ifndef SDQLB200_EMU
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define SDQL_DEV __device__ __forceinline__
#define SDQL_EXTERN_SMEM(name) extern __shared__ __align__(128) unsigned long long name[]
#define SDQL_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#else
#include "sdqlb200_emu.h"
#define SDQL_DEV static inline
#endif

namespace sdqlrt {

typedef unsigned long long u64;
typedef long long i64;

constexpr int kBlock = 256;   // threads per CTA for every generated kernel
constexpr int kVec = 4;       // rows per thread per iteration
#ifndef SDQLB200_EMU
constexpr int kLanes = 32;    // threads that share one byte-row staging buffer (a warp)
#else
constexpr int kLanes = 1;
#endif
constexpr int kStageRows = kLanes * kVec;  // rows one warp stages per iteration
constexpr int kQueueSlots = kLanes * kVec + kLanes;  // hit compaction: row ids one warp can hold (leftover < kLanes + one iteration)
constexpr u64 kEmpty = ~0ull;

// ---------------------------------------------------------------------------------------------
// SDQLB200_STATS (diagnostic builds only, never timed): counts of the data-dependent memory operations of a query --
// the inputs of the "bytes-moved" roofline of join-heavy queries (scan bytes + tables written + 32 B per
// random access).  Read and reset through sdqlb200_stats().  A regular build compiles every stat() call to nothing.
// ---------------------------------------------------------------------------------------------
enum {
    kStBitTests = 0,    // presence-bit tests (one 4-byte load each; the bitmap is L2 resident)
    kStFinds = 1,       // probes that went past the presence bits into the table
    kStFindSlots = 2,   // table slots those probes touched (1 for a direct table, the probe-chain length for a hashed one)
    kStUpserts = 3,     // insert-or-find operations of table builds / group-bys
    kStUpsertSlots = 4, // table slots they touched
    kStAtomics = 5,     // global red.add on aggregate arrays
    kStGathers = 6,     // data-dependent loads of column values / representative rows (late materialisation)
    kStCount = 8
};
#ifdef SDQLB200_STATS
#ifndef SDQLB200_EMU
__device__ unsigned long long g_stats[kStCount];
SDQL_DEV void stat(int k, unsigned long long n = 1) {
    const unsigned m = __activemask();  // one atomic per converged group of lanes
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&g_stats[k], n * (unsigned long long)__popc(m));
}
#else
static unsigned long long g_stats[kStCount];
SDQL_DEV void stat(int k, unsigned long long n = 1) { g_stats[k] += n; }
#endif
#else
SDQL_DEV void stat(int, unsigned long long = 1) {}
#endif

// ---------------------------------------------------------------------------------------------
// streaming loads
// ---------------------------------------------------------------------------------------------
#ifndef SDQLB200_EMU
SDQL_DEV void ld4(const int* p, int (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p));
}
SDQL_DEV void ld4(const double* p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "l"(p + 2));
}
SDQL_DEV void ld4(const unsigned char* p, int (&v)[4]) {
    unsigned w;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w) : "l"(p));
    v[0] = w & 0xff; v[1] = (w >> 8) & 0xff; v[2] = (w >> 16) & 0xff; v[3] = w >> 24;
}
template <class T> SDQL_DEV T ld1(const T* p) { return __ldg(p); }
SDQL_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p)); }
#else
SDQL_DEV void prefetch_l2(const void*) {}
template <class T, class V> SDQL_DEV void ld4(const T* p, V (&v)[4]) { for (int k = 0; k < 4; ++k) v[k] = (V)p[k]; }
template <class T> SDQL_DEV T ld1(const T* p) { return *p; }
#endif

// ---------------------------------------------------------------------------------------------
// column tile ring: cp.async.bulk (TMA engine, SASS UBLKCP) + mbarrier transaction counting.
// One elected thread arms full[s] with the stage's byte count and issues one bulk copy per scanned column; every
// thread waits on full[s], reads its rows from shared memory, and each warp arrives on empty[s] so the stage can be
// refilled.  No registers are tied up by loads in flight: bytes in flight = stages x tile, independent of occupancy.
// ---------------------------------------------------------------------------------------------
#ifndef SDQLB200_EMU
SDQL_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SDQL_DEV void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
SDQL_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SDQL_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
SDQL_DEV void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
SDQL_DEV bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
SDQL_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar` (complete_tx)
SDQL_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------
// byte-row staging: the rows of a fixed-width string column that a warp is about to examine (kLanes x kVec consecutive
// rows = one contiguous run of bytes) are copied into the warp's shared-memory buffer with coalesced 128-bit loads, so
// the per-row character loops read shared memory (row strides are odd multiples of 4 bytes for odd widths: conflict
// free) instead of issuing one uncoalesced global byte load per character.  All lanes of the warp must call.
// ---------------------------------------------------------------------------------------------
SDQL_DEV void stage_rows(unsigned char* dst, const unsigned char* col, i64 row0, i64 n, int W) {
    const i64 r1 = row0 + kStageRows < n ? row0 + kStageRows : n;
#ifndef SDQLB200_EMU
    __syncwarp();  // the previous iteration's readers are done with the buffer
    if (r1 > row0) {
        const size_t bytes = (size_t)(r1 - row0) * (size_t)W;
        const unsigned char* src = col + row0 * W;  // 16-byte aligned: row0 is a multiple of 128, the column base of 256
        const int lane = threadIdx.x & 31;
        const size_t nv = bytes >> 4;
        for (size_t k = lane; k < nv; k += 32) {
            uint4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"((const uint4*)src + k));
            ((uint4*)dst)[k] = v;
        }
        for (size_t k = (nv << 4) + lane; k < bytes; k += 32) dst[k] = ld1(src + k);
    }
    __syncwarp();
#else
    if (r1 > row0) memcpy(dst, col + row0 * W, (size_t)(r1 - row0) * (size_t)W);
#endif
}

#ifndef SDQLB200_EMU
// primitives of the warp text scan (sdqlb200_textscan.cuh)
#define TX_NOINLINE __device__ __noinline__
SDQL_DEV int tx_lane() { return (int)(threadIdx.x & 31u); }
SDQL_DEV void tx_syncwarp() { __syncwarp(); }
SDQL_DEV unsigned tx_shfl_down(unsigned v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
SDQL_DEV unsigned tx_shfl(unsigned v, int l) { return __shfl_sync(0xffffffffu, v, l); }
SDQL_DEV unsigned tx_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
SDQL_DEV int tx_ffs(unsigned v) { return __ffs((int)v); }
SDQL_DEV void tx_atomic_or(unsigned* p, unsigned v) { atomicOr(p, v); }
SDQL_DEV void tx_ldnc16(const unsigned char* p, unsigned (&w)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
}
#endif
#include "sdqlb200_textscan.cuh"

#ifndef SDQLB200_EMU
SDQL_DEV unsigned warp_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
SDQL_DEV void warp_sync() { __syncwarp(); }
#else
SDQL_DEV unsigned warp_ballot(bool p) { return p ? 1u : 0u; }
SDQL_DEV void warp_sync() {}
#endif

// dictionary-code columns are uint8 when the dictionary has <= 256 entries, int32 otherwise (uniform branch)
SDQL_DEV void ld4_code(const void* p, i64 i0, int width, int (&v)[4]) {
    if (width == 1) ld4((const unsigned char*)p + i0, v);
    else ld4((const int*)p + i0, v);
}
SDQL_DEV int ld1_code(const void* p, i64 i, int width) {
    return width == 1 ? (int)ld1((const unsigned char*)p + i) : ld1((const int*)p + i);
}
// data-dependent (gather) loads of a column value at a row found through a table: late materialisation
template <class T> SDQL_DEV T ldg1(const T* p) { stat(kStGathers); return ld1(p); }
SDQL_DEV int ldg1_code(const void* p, i64 i, int width) { stat(kStGathers); return ld1_code(p, i, width); }

// ---------------------------------------------------------------------------------------------
// device dictionary
// ---------------------------------------------------------------------------------------------
struct Tbl {
    u64* keys;   // hash mode: packed key per slot (kEmpty = free); unused in direct mode
    int* rep;    // representative source index (row of the scanned relation / slot of the scanned table); -1 = free;
                 // after a cross-GPU merge: -2 = present but owned (iterated, late-materialised) by another rank
    i64 cap;     // slots (direct: key domain size; hash: power of two)
    int direct;  // 1: slot == packed key
    // presence filter in front of selective tables: one bit per value of the packed key (bmod == 0) or of its first part
    // (bmod = that part's range: mixed-radix packing puts it in the low digits).  nullptr = no filter.
    unsigned* bits;
    u64 bmod;
};

SDQL_DEV bool tbl_maybe(const Tbl& t, u64 key) {
    if (!t.bits) return true;
    stat(kStBitTests);
    const u64 b = t.bmod ? key % t.bmod : key;
    return (ld1(t.bits + (b >> 5)) >> (unsigned)(b & 31)) & 1u;
}
// the same with the packed first key part at hand (p0 == key % bmod): no 64-bit modulo
SDQL_DEV bool tbl_maybe(const Tbl& t, u64 key, u64 p0) {
    if (!t.bits) return true;
    stat(kStBitTests);
    const u64 b = t.bmod ? p0 : key;
    return (ld1(t.bits + (b >> 5)) >> (unsigned)(b & 31)) & 1u;
}
// dictionary-valued entry d[k] of a table keyed by (k, x): may any (k, x) be present?  outer = packed k, x in [0, rng)
SDQL_DEV bool tbl_maybe_outer(const Tbl& t, u64 outer, i64 rng, i64 mul, u64 p0) {
    if (!t.bits) return true;
    if (t.bmod) return tbl_maybe(t, outer, p0);
    for (i64 x = 0; x < rng; ++x)
        if (tbl_maybe(t, outer + (u64)x * (u64)mul)) return true;
    return false;
}
SDQL_DEV u64 hash64(u64 x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}

// Probe of a hashed table (open addressing, linear probing).  The probe examines the whole 32-byte sector the home slot
// lies in -- four 8-byte keys, two 128-bit loads -- per memory round trip: the sector is what a random 8-byte read moves
// anyway, and a probe chain of up to four slots inside it resolves without a second dependent load (what a 4-lane
// cooperative group per key would achieve, without giving up 3 of 4 lanes to it).  Tables have >= 1024 slots (power of
// two), so a group never wraps.  kProbeKeys = 4 is the whole sector; -DSDQLB200_PROBE_KEYS=2 examines an aligned pair (one
// 128-bit load, 4 registers less: inlined into every probe site the full group cost q19_k1 -- whose tables are direct or
// tiny -- an occupancy step, 48 -> 60 registers, 2.12 -> 2.52 ms at SF100; an out-of-line probe cost every caller the ABI).  B200, SF100 (profiles/r02_visit10/r02_v10_ab_probe_sf100.json): q9_k5 6.85 ->
// 6.46 ms, q20_k3 3.47 -> 3.17 ms, Q2 0.746 -> 0.717 ms against one slot per step (-DSDQLB200_PROBE_SLOT keeps that).
#ifndef SDQLB200_PROBE_KEYS
#define SDQLB200_PROBE_KEYS 2
#endif
constexpr int kProbeKeys = SDQLB200_PROBE_KEYS;
SDQL_DEV int tbl_probe(const Tbl& t, u64 key) {
    const u64 m = (u64)t.cap - 1;
    u64 h = hash64(key) & m;
#if !defined(SDQLB200_PROBE_SLOT) && !defined(SDQLB200_EMU)
    for (;;) {
        stat(kStFindSlots);
        const u64 base = h & ~(u64)(kProbeKeys - 1);
        u64 k4[kProbeKeys];
#pragma unroll
        for (int j = 0; j < kProbeKeys; j += 2) {
            const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(t.keys + base + j));
            k4[j] = a.x; k4[j + 1] = a.y;
        }
        const int j0 = (int)(h - base);
#pragma unroll
        for (int j = 0; j < kProbeKeys; ++j) {
            if (j < j0) continue;
            if (k4[j] == key) return (int)(base + j);
            if (k4[j] == kEmpty) return -1;
        }
        h = (base + kProbeKeys) & m;
    }
#else
    for (;;) {
        stat(kStFindSlots);
        u64 k = ld1(t.keys + h);
        if (k == key) return (int)h;
        if (k == kEmpty) return -1;
        h = (h + 1) & m;
    }
#endif
}

// -> slot or -1.  `ok` = every key part was inside the table's packing range.  p0: the packed first key part.
SDQL_DEV int tbl_find(const Tbl& t, u64 key, bool ok, u64 p0) {
    if (!ok || !tbl_maybe(t, key, p0)) return -1;
    stat(kStFinds);
    if (t.direct) { stat(kStFindSlots); return ld1(t.rep + key) != -1 ? (int)key : -1; }  // -2 = present, owned by another rank
    return tbl_probe(t, key);
}
SDQL_DEV int tbl_find(const Tbl& t, u64 key, bool ok) {
    if (!ok || !tbl_maybe(t, key)) return -1;
    stat(kStFinds);
    if (t.direct) { stat(kStFindSlots); return ld1(t.rep + key) != -1 ? (int)key : -1; }  // -2 = present, owned by another rank
    return tbl_probe(t, key);
}

// Single-part keys whose part is an int32 column value (the common probe: a foreign key looked up in a table keyed by the
// primary key): the packed key is the 32-bit offset from the build column's minimum, and the probe needs no 64-bit
// multiply / modulo.  Same results as pack_part + tbl_find (a single-part table never has a first-part modulus).
// Strided-dense parts (sdqlb200_col.stride): the column only uses the first `sk` residues of every block of 2^sb values, so
// the holes are packed away: d -> (d >> sb) * sk + (d & (2^sb - 1)).  sb == 0: plain dense packing.
SDQL_DEV bool pack_key1(int x, i64 mn, i64 rng, int sb, int sk, u64& key) {
    const int m = (int)mn;  // statistics of an int32 / dictionary-code column
    unsigned d = (unsigned)x - (unsigned)m;
    bool ok = x >= m;
    if (sb) { const unsigned lo = d & ((1u << sb) - 1u); ok &= lo < (unsigned)sk; d = (d >> sb) * (unsigned)sk + lo; }
    key = d;
    return ok && (u64)d < (u64)rng;
}
SDQL_DEV int tbl_find1(const Tbl& t, unsigned d, bool ok) {
    if (!ok) return -1;
    if (t.bits) {
        stat(kStBitTests);
        if (!((ld1(t.bits + (d >> 5)) >> (d & 31u)) & 1u)) return -1;
    }
    stat(kStFinds);
    if (t.direct) { stat(kStFindSlots); return ld1(t.rep + d) != -1 ? (int)d : -1; }
    return tbl_probe(t, (u64)d);
}

// insert-or-find; `src` becomes the slot's representative if the slot is new.  -> slot
SDQL_DEV int tbl_upsert(const Tbl& t, u64 key, int src, bool& is_new) {
    stat(kStUpserts);
    if (t.direct) {
        stat(kStUpsertSlots);
        int old = t.rep[key];
        if (old < 0) old = atomicCAS(t.rep + key, -1, src);
        is_new = old < 0;
        return (int)key;
    }
    u64 m = (u64)t.cap - 1, h = hash64(key) & m;
    for (;;) {
        stat(kStUpsertSlots);
        u64 k = t.keys[h];
        if (k == key) { is_new = false; return (int)h; }
        if (k == kEmpty) {
            u64 prev = atomicCAS(t.keys + h, kEmpty, key);
            if (prev == kEmpty) { t.rep[h] = src; is_new = true; return (int)h; }
            if (prev == key) { is_new = false; return (int)h; }
        }
        h = (h + 1) & m;
    }
}

SDQL_DEV int rep_of(const Tbl& t, int slot) {  // safe for slot == -1 (returns a valid index >= 0)
    if (slot >= 0) stat(kStGathers);
    int r = ld1(t.rep + (slot < 0 ? 0 : slot));
    return r < 0 ? 0 : r;
}

// mixed-radix key packing: part p in [mn, mn + rng) contributes (p - mn) * mul
SDQL_DEV bool pack_part(i64 p, i64 mn, i64 rng, i64 mul, int sb, int sk, u64& key) {
    u64 d = (u64)(p - mn);
    bool ok = true;
    if (sb) { const u64 lo = d & ((1ull << sb) - 1ull); ok = lo < (u64)sk; d = (d >> sb) * (u64)sk + lo; }
    key += d * (u64)mul;
    return ok && d < (u64)rng;
}

SDQL_DEV i64 unpack_part(u64 key, i64 mn, i64 rng, i64 mul, int sb, int sk) {
    if (rng < 0) return (i64)key;  // single unbounded part
    u64 d = (key / (u64)mul) % (u64)rng;
    if (sb) d = ((d / (u64)sk) << sb) + d % (u64)sk;
    return (i64)d + mn;
}
SDQL_DEV u64 tbl_key(const Tbl& t, i64 slot) { return t.direct ? (u64)slot : ld1(t.keys + slot); }

// value range of an int64 aggregate array over ALL of its slots (free slots hold 0: whatever slot a lookup reads, its value is
// inside the range) joined with 0: mm[0] = min(0, values), mm[1] = max(0, values); mm is zeroed by the caller.  Group-bys keyed
// by a value read out of another dictionary (Q13: customers per order count) get their key domain from it at run time.
__global__ void k_minmax_i64(const i64* a, i64 n, i64* mm) {
    i64 lo = 0, hi = 0;
#ifndef SDQLB200_EMU
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 v = ld1(a + i);
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const i64 l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        if (lo < 0) atomicMin((long long*)mm, (long long)lo);
        if (hi > 0) atomicMax((long long*)mm + 1, (long long)hi);
    }
#else
    for (i64 i = 0; i < n; ++i) { lo = a[i] < lo ? a[i] : lo; hi = a[i] > hi ? a[i] : hi; }
    mm[0] = lo < mm[0] ? lo : mm[0];
    mm[1] = hi > mm[1] ? hi : mm[1];
#endif
}

// presence filter of a finished table: one pass over its slots right after the build kernel (no per-insert atomics in
// the build).  A warp examines 32 consecutive slots; when their bits fall into one word (direct tables: always, unless
// the first-part modulus wraps inside the group) the warp issues a single atomicOr.
__global__ void k_tbl_bits(Tbl t) {
#ifndef SDQLB200_EMU
    const int lane = threadIdx.x & 31;
    const i64 nwarp = ((i64)gridDim.x * blockDim.x) >> 5;
    if (t.direct && t.bmod == 0) {
        // dense array, one bit per slot: every lane loads 4 x int4 (16 slots), a warp step covers 512 slots = 16 whole
        // words of the bitmap, each assembled from the nibbles of 8 neighbouring lanes and written without atomics
        // (the rep array is padded to 256 bytes: the last partial int4 is readable, its slots >= cap are masked)
        const i64 nquad = (t.cap + 3) >> 2;
        for (i64 q0 = (((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 128; q0 < nquad; q0 += nwarp * 128) {
            int4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const i64 q = q0 + k * 32 + lane;
                if (q < nquad) {
                    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(t.rep + 4 * q));
                } else {
                    v[k] = make_int4(-1, -1, -1, -1);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const i64 s0 = 4 * (q0 + k * 32 + lane);
                unsigned nib = (v[k].x != -1 && s0 < t.cap ? 1u : 0u) | (v[k].y != -1 && s0 + 1 < t.cap ? 2u : 0u) |
                               (v[k].z != -1 && s0 + 2 < t.cap ? 4u : 0u) | (v[k].w != -1 && s0 + 3 < t.cap ? 8u : 0u);
                unsigned w = nib << (4 * (lane & 7));
                w |= __shfl_xor_sync(0xffffffffu, w, 1);
                w |= __shfl_xor_sync(0xffffffffu, w, 2);
                w |= __shfl_xor_sync(0xffffffffu, w, 4);
                if ((lane & 7) == 0 && w) t.bits[((q0 + k * 32) >> 3) + (lane >> 3)] = w;
            }
        }
        return;
    }
    for (i64 base = ((((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5); base < t.cap; base += nwarp << 5) {
        const i64 i = base + lane;
        bool occ = false;
        u64 key = 0;
        if (i < t.cap) {
            if (t.direct) { occ = t.rep[i] != -1; key = (u64)i; }
            else { key = t.keys[i]; occ = key != kEmpty; }
        }
        const unsigned any = __ballot_sync(0xffffffffu, occ);
        if (!any) continue;
        const u64 b = occ ? (t.bmod ? key % t.bmod : key) : 0;
        const unsigned long long word = b >> 5;
        const unsigned m = occ ? 1u << (unsigned)(b & 31) : 0u;
        const int leader = __ffs(any) - 1;
        const unsigned long long w0 = __shfl_sync(0xffffffffu, word, leader);
        if (__all_sync(0xffffffffu, !occ || word == w0)) {
            const unsigned r = __reduce_or_sync(0xffffffffu, m);
            if (lane == leader) atomicOr(t.bits + w0, r);
        } else if (occ) {
            atomicOr(t.bits + word, m);
        }
    }
#else
    for (i64 i = 0; i < t.cap; ++i) {
        const bool occ = t.direct ? t.rep[i] != -1 : t.keys[i] != kEmpty;
        if (!occ) continue;
        const u64 key = t.direct ? (u64)i : t.keys[i];
        const u64 b = t.bmod ? key % t.bmod : key;
        t.bits[b >> 5] |= 1u << (unsigned)(b & 31);
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// atomics / reductions
// ---------------------------------------------------------------------------------------------
SDQL_DEV void red_add(double* p, double v) { stat(kStAtomics); atomicAdd(p, v); }
SDQL_DEV void red_add(i64* p, i64 v) { stat(kStAtomics); atomicAdd((u64*)p, (u64)v); }

#ifndef SDQLB200_EMU
template <class T> SDQL_DEV T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SDQL_DEV int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// all threads must call; result valid in thread 0.  Sum order is fixed (lane tree, then warp 0 tree) so the
// result is run-to-run deterministic for a fixed grid.
template <class T> SDQL_DEV T block_sum(T v) {
    __shared__ T sh[kBlock / 32];
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : (T)0;
        r = warp_sum(r);
    }
    return r;
}
SDQL_DEV int block_max(int v) {
    __shared__ int shm[kBlock / 32];
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = -1;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? shm[threadIdx.x] : -1;
        r = warp_max(r);
    }
    return r;
}
// warp-aggregated append: one atomicAdd per converged group of appending lanes
SDQL_DEV u64 append_slot(u64* counter) {
    auto g = cooperative_groups::coalesced_threads();
    u64 base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, (u64)g.size());
    return g.shfl(base, 0) + g.thread_rank();
}
#else
template <class T> SDQL_DEV T block_sum(T v) { return v; }
SDQL_DEV int block_max(int v) { return v; }
SDQL_DEV u64 append_slot(u64* counter) { return (*counter)++; }
#endif

// "last block done" election for two-level reductions; `counter` must be zero before the launch and is reset.
SDQL_DEV bool last_block(unsigned* counter) {
#ifndef SDQLB200_EMU
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
        if (last) *counter = 0;
    }
    __syncthreads();
    if (last) __threadfence();
    return last;
#else
    return true;
#endif
}

// ---------------------------------------------------------------------------------------------
// cross-GPU merge of a direct-indexed table: every occupied slot gets exactly one owner rank (the lowest rank that
// saw the key); aggregate arrays are all-reduced separately.  encode -> all-reduce(min) -> decode.
// ---------------------------------------------------------------------------------------------
__global__ void k_owner_encode(const int* rep, int* own, long long cap, int rank) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (long long)gridDim.x * blockDim.x)
        own[i] = rep[i] >= 0 ? rank : 0x7fffffff;
}
__global__ void k_owner_decode(int* rep, const int* own, long long cap, int rank) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (long long)gridDim.x * blockDim.x) {
        const int o = own[i];
        if (o != rank) rep[i] = (o == 0x7fffffff) ? -1 : -2;
    }
}

// ---------------------------------------------------------------------------------------------
// cross-GPU merge of a HASHED table ("all-reduce of a dictionary" = hash all-to-all + all-gather):
//   count/pack   every occupied slot becomes one record [key, owner rank, field 0 .. field nf-1] (8-byte words),
//                grouped by destination rank = hash(key) mod world                         -> NCCL all-to-all
//   absorb(0)    the destination combines the records of all ranks in a scratch table (sum of the fields, owner = the
//                lowest rank that saw the key); its entries are packed again                -> NCCL all-gather
//   absorb(1)    every rank writes the combined entries back into its own table: fields := global sums, slots owned by
//                another rank get rep = -2 (present for probes, skipped when the table is iterated)
// ---------------------------------------------------------------------------------------------
struct TblIO {
    u64* keys;
    int* rep;
    i64 cap;
    int nf;
    unsigned f64_mask;  // bit j: field j is fp64 (else int64)
    u64* agg[16];
};
// small direct-indexed tables (low-cardinality group-by: the common case) are merged with ONE all-reduce: presence,
// fp64 fields and int64 fields travel together as fp64 words of one staging buffer [1 + nf64 + 2 * ni64][cap]:
//   presence   2^rank if this rank has the key (sum = bit mask of the ranks that saw it; owner = lowest set bit)
//   fp64 field the value itself (exactly the in-place all-reduce)
//   int64      high and low 32-bit halves as two exact fp64 words (sums stay below 2^53); re-assembled modulo 2^64
__global__ void k_merge_pack(TblIO t, int rank, double* mg) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (i64)gridDim.x * blockDim.x) {
        mg[i] = t.rep[i] >= 0 ? (double)(1ull << rank) : 0.0;
        i64 col = 1;
        for (int j = 0; j < t.nf; ++j) {
            const u64 v = t.agg[j][i];
            if ((t.f64_mask >> j) & 1u) { mg[col * t.cap + i] = __longlong_as_double((i64)v); col += 1; }
            else { mg[col * t.cap + i] = (double)(v >> 32); mg[(col + 1) * t.cap + i] = (double)(v & 0xffffffffull); col += 2; }
        }
    }
}
__global__ void k_merge_unpack(TblIO t, int rank, const double* mg) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (i64)gridDim.x * blockDim.x) {
        const u64 mask = (u64)mg[i];
        if (mask) {
            int owner = 0;
            while (!((mask >> owner) & 1ull)) ++owner;
            if (owner != rank) t.rep[i] = -2;  // present, iterated (late-materialised) by its owner rank
        }
        i64 col = 1;
        for (int j = 0; j < t.nf; ++j) {
            if ((t.f64_mask >> j) & 1u) { t.agg[j][i] = (u64)__double_as_longlong(mg[col * t.cap + i]); col += 1; }
            else { t.agg[j][i] = ((u64)mg[col * t.cap + i] << 32) + (u64)mg[(col + 1) * t.cap + i]; col += 2; }
        }
    }
}
constexpr i64 kFusedMergeMaxSlots = 65536;

SDQL_DEV int shuffle_dest(u64 key, int world) { return (int)((hash64(key ^ 0x9e3779b97f4a7c15ull) >> 33) % (u64)world); }

__global__ void k_tbl_count(TblIO t, int world, u64* counts) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (i64)gridDim.x * blockDim.x) {
        const u64 k = t.keys[i];
        if (k != kEmpty) atomicAdd(counts + shuffle_dest(k, world), 1ull);
    }
}
// own == nullptr: owner word = rank (first hop); else owner word = own[slot] (second hop, scratch table)
__global__ void k_tbl_pack(TblIO t, int world, int rank, const int* own, const u64* offsets, u64* cursor, i64* rec) {
    const int W = 2 + t.nf;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (i64)gridDim.x * blockDim.x) {
        const u64 k = t.keys[i];
        if (k == kEmpty) continue;
        const int d = shuffle_dest(k, world);
        const u64 pos = offsets[d] + atomicAdd(cursor + d, 1ull);
        i64* r = rec + pos * W;
        r[0] = (i64)k;
        r[1] = own ? (i64)own[i] : (i64)rank;
        for (int j = 0; j < t.nf; ++j) r[2 + j] = (i64)t.agg[j][i];
    }
}
__global__ void k_tbl_absorb(TblIO t, const i64* rec, i64 n, int mode, int rank, int* own) {
    const int W = 2 + t.nf;
    const u64 m = (u64)t.cap - 1;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64* r = rec + i * W;
        const u64 key = (u64)r[0];
        u64 h = hash64(key) & m;
        bool claimed = false;
        for (;;) {  // find or claim (the table is sized for the union of all ranks' keys)
            u64 k = t.keys[h];
            if (k == kEmpty) {
                k = atomicCAS(t.keys + h, kEmpty, key);
                if (k == kEmpty) { claimed = true; break; }
            }
            if (k == key) break;
            h = (h + 1) & m;
        }
        if (mode == 0) {
            for (int j = 0; j < t.nf; ++j) {
                if ((t.f64_mask >> j) & 1u) red_add((double*)t.agg[j] + h, __longlong_as_double(r[2 + j]));
                else red_add((i64*)t.agg[j] + h, r[2 + j]);
            }
            atomicMin(own + h, (int)r[1]);
        } else {
            for (int j = 0; j < t.nf; ++j) t.agg[j][h] = (u64)r[2 + j];
            if ((int)r[1] != rank || claimed) t.rep[h] = -2;
        }
    }
}

// sparse merge of a DIRECT-indexed partial dictionary (SDQLB200_MERGE_DIRECT): records are (slot, owner rank, fields)
__global__ void k_dtbl_count(TblIO t, u64* count) {
    u64 c = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (i64)gridDim.x * blockDim.x) c += t.rep[i] >= 0;
#ifndef SDQLB200_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
#else
    *count += c;
#endif
}
__global__ void k_dtbl_pack(TblIO t, int rank, u64* cursor, i64* rec) {
    const int W = 2 + t.nf;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < t.cap; i += (i64)gridDim.x * blockDim.x) {
        if (t.rep[i] < 0) continue;
        i64* r = rec + atomicAdd(cursor, 1ull) * W;
        r[0] = i;
        r[1] = (i64)rank;
        for (int j = 0; j < t.nf; ++j) r[2 + j] = (i64)t.agg[j][i];
    }
}
// the OTHER ranks' records (lo <= index < hi skipped: this rank's own run): fields added in place, ownership settled --
// an entry a lower rank also holds, or that this rank does not hold, becomes rep = -2 (present, iterated by its owner)
__global__ void k_dtbl_absorb(TblIO t, const i64* rec, i64 n, i64 lo, i64 hi, int rank) {
    const int W = 2 + t.nf;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        if (i >= lo && i < hi) continue;
        const i64* r = rec + i * W;
        const i64 slot = r[0];
        for (int j = 0; j < t.nf; ++j) {
            if ((t.f64_mask >> j) & 1u) red_add((double*)t.agg[j] + slot, __longlong_as_double(r[2 + j]));
            else red_add((i64*)t.agg[j] + slot, r[2 + j]);
        }
        if ((int)r[1] < rank) t.rep[slot] = -2;       // (racing writers all store -2)
        else atomicCAS(t.rep + slot, -1, -2);
    }
}

// ---------------------------------------------------------------------------------------------
// fixed-width, zero-padded byte strings (device form of VarChar<N>, 1 byte per char)
// ---------------------------------------------------------------------------------------------
SDQL_DEV int str_len(const unsigned char* s, int w) {
    int n = 0;
    while (n < w && s[n]) ++n;
    return n;
}
// first index of pat in s, -1 if absent (varchar.h:91-97 firstIndex -> wcsstr, which stops at the first NUL)
// single pass: a match cannot start at or span a NUL (patterns contain none), so the scan simply stops at the first NUL.
// Four characters are loaded per step, one step ahead of their use: the early-exit branches do not wait on a load each.
SDQL_DEV int str_find_bytes(const unsigned char* s, int w, const char* pat, int plen) {
    const int nst = w - plen + 1;  // start positions
    if (nst <= 0) return -1;
    const unsigned char p0 = (unsigned char)pat[0];
    unsigned char c0 = s[0], c1 = 1 < w ? s[1] : 0, c2 = 2 < w ? s[2] : 0, c3 = 3 < w ? s[3] : 0;
    for (int i = 0; i < nst; i += 4) {
        const unsigned char d0 = i + 4 < w ? s[i + 4] : 0, d1 = i + 5 < w ? s[i + 5] : 0, d2 = i + 6 < w ? s[i + 6] : 0,
                            d3 = i + 7 < w ? s[i + 7] : 0;
        const unsigned char cc[4] = {c0, c1, c2, c3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j >= nst || cc[j] == 0) return -1;
            if (cc[j] == p0) {
                int k = 1;
                while (k < plen && s[i + j + k] == (unsigned char)pat[k]) ++k;
                if (k == plen) return i + j;
            }
        }
        c0 = d0; c1 = d1; c2 = d2; c3 = d3;
    }
    return -1;
}

// 0x80 in every byte of x that is zero (exact per byte)
SDQL_DEV unsigned zero_bytes(unsigned x) { return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu); }

// Word-wise search: the row is read as aligned 32-bit words (the row itself may start at any byte; the bytes in front of
// it belong to the previous row of the column), four start positions per step.  A step is decided by three SWAR tests
// -- a NUL among the four characters, the pattern's FIRST character at position j and its SECOND at j + 1 -- so the
// per-character path runs only for a NUL (once per row) or a two-character prefix match (rare); a first-character-only
// filter would send every warp down that path at nearly every step (some lane always sees the letter).  The next word
// is loaded one step ahead of its use.  The last < 12 characters are searched by the byte loop, so no word beyond the
// row's end is touched.
SDQL_DEV int str_find(const unsigned char* s, int w, const char* pat, int plen) {
    if (plen < 2 || w < 16) return str_find_bytes(s, w, pat, plen);
    const int last = w - plen;  // last possible start
    if (last < 0) return -1;
    const unsigned a = (unsigned)(size_t)s & 3u;
    const unsigned* wp = (const unsigned*)(s - a);
    const unsigned p0x4 = (unsigned char)pat[0] * 0x01010101u, p1x4 = (unsigned char)pat[1] * 0x01010101u;
    unsigned w0 = wp[0], w1 = wp[1];
    int i = 0;
    for (; i + 12 <= w; i += 4) {  // reads up to the aligned word holding s[i + 11 - a]: inside the row
        const unsigned w2 = wp[(i >> 2) + 2];
        const unsigned v = __funnelshift_rc(w0, w1, a * 8u);        // characters i .. i+3
        const unsigned v1 = __funnelshift_rc(w0, w1, a * 8u + 8u);  // characters i+1 .. i+4
        w0 = w1;
        w1 = w2;
        const unsigned z = zero_bytes(v);
        const unsigned m = zero_bytes(v ^ p0x4) & zero_bytes(v1 ^ p1x4);
        if (!(z | m)) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((z >> (8 * j + 7)) & 1u) return -1;
            if (((m >> (8 * j + 7)) & 1u) && i + j <= last) {
                int k = 2;
                while (k < plen && s[i + j + k] == (unsigned char)pat[k]) ++k;
                if (k == plen) return i + j;
            }
        }
    }
    const int r = str_find_bytes(s + i, w - i, pat, plen);
    return r < 0 ? -1 : r + i;
}
// out-of-line copy for call sites that run rarely (rows that passed the warp text scan): keeps the kernels small
#ifndef SDQLB200_EMU
__device__ __noinline__ int str_find_rare(const unsigned char* s, int w, const char* pat, int plen) { return str_find(s, w, pat, plen); }
#else
static inline int str_find_rare(const unsigned char* s, int w, const char* pat, int plen) { return str_find(s, w, pat, plen); }
#endif
SDQL_DEV bool str_starts(const unsigned char* s, int w, const char* pat, int plen) {  // varchar.h:99-111
    if (plen > w) return false;
    for (int j = 0; j < plen; ++j)
        if (s[j] != (unsigned char)pat[j]) return false;
    return true;
}
// varchar.h:113-125: endsWith == (firstIndex(pat) == len - plen), i.e. the FIRST occurrence must be the suffix
SDQL_DEV bool str_ends(const unsigned char* s, int w, const char* pat, int plen) {
    return str_find(s, w, pat, plen) == str_len(s, w) - plen;
}
SDQL_DEV bool str_eq(const unsigned char* s, int w, const char* pat, int plen) {  // varchar.h:61-77
    if (plen > w) return false;
    for (int j = 0; j < plen; ++j)
        if (s[j] != (unsigned char)pat[j]) return false;
    for (int j = plen; j < w; ++j)
        if (s[j]) return false;
    return true;
}
SDQL_DEV i64 str_pack(const unsigned char* s, int from, int n) {  // substr<n>(from, from+n-1) as an integer key
    u64 v = 0;
    for (int j = 0; j < n; ++j) v = (v << 8) | s[from + j];
    return (i64)v;
}

}  // namespace sdqlrt
