// Host-side helpers for generated query modules: workspace arena, run-time table sizing (direct-indexed vs
// hash, from column statistics), error plumbing.  Replaces the container declarations the reference emits up
// front (outputDictionariesInitializationCode, sdql_ir_cpp_generator_par.py:205, 267, 309, 347, 418).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sdqlb200.h"
#include "sdqlb200_rt.cuh"

namespace sdqlhost {

static thread_local char g_err[512];

static inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define SDQL_CUDA(x)                                                                                     \
    do {                                                                                                 \
        cudaError_t e_ = (x);                                                                            \
        if (e_ != cudaSuccess)                                                                           \
            return sdqlhost::fail(SDQLB200_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)

// bump allocator over the caller's device workspace; keeps counting past the end so a dry run yields the size
struct Arena {
    char* base;
    unsigned long long cap, used;
    Arena(void* p, unsigned long long n) : base((char*)p), cap(p ? n : 0), used(0) {}
    template <class T> T* alloc(long long n) {
        unsigned long long bytes = ((unsigned long long)(n > 0 ? n : 1) * sizeof(T) + 255) & ~255ull;
        unsigned long long off = used;
        used += bytes;
        return (base && used <= cap) ? (T*)(base + off) : nullptr;
    }
    bool ok() const { return base && used <= cap; }
};

struct TblPlan {
    long long ff_off, ff_bytes;  // region to fill with 0xFF (rep, keys)
};

// Decide direct vs hash and allocate.  parts: value ranges [mn, mn+rng) of the by-value key parts.
// src_rows bounds the number of distinct keys.  Returns false if the key domain cannot be packed in 63 bits.
static inline bool size_table(sdqlrt::Tbl* t, int nparts, const long long* mn, const long long* rng,
                              long long src_rows, long long* o_mn, long long* o_rng, long long* o_mul, Arena& ar) {
    long double dom = 1;
    long long mul = 1;
    for (int j = 0; j < nparts; ++j) {
        o_mn[j] = mn[j];
        o_mul[j] = mul;
        if (rng[j] < 0) {  // unbounded part (raw 64-bit value): hash mode, must be the only part
            o_rng[j] = -1;
            if (nparts != 1) return false;
            dom = 9.0e18L;
            break;
        }
        o_rng[j] = rng[j] < 1 ? 1 : rng[j];
        dom *= (long double)o_rng[j];
        if (dom > 9.0e18L) {
            if (nparts == 1) { dom = 9.0e18L; break; }
            return false;
        }
        mul *= o_rng[j];
    }
    if (src_rows < 1) src_rows = 1;
    long long domain = (long long)dom;
    // direct when the dense array is not (much) bigger than what a hash table for src_rows keys would need.
    // SDQLB200_FORCE_HASH=1 (debug knob): never direct -- exercises the hash build / probe / merge paths at small scale
    static const bool force_hash = getenv("SDQLB200_FORCE_HASH") && getenv("SDQLB200_FORCE_HASH")[0] == '1';
    bool direct = !force_hash && dom <= (long double)(8 * src_rows + 65536) && domain < (1ll << 31);
    t->direct = direct ? 1 : 0;
    if (direct) {
        t->cap = domain;
        t->keys = nullptr;
    } else {
        long long need = src_rows < domain ? src_rows : domain;
        long long cap = 1024;
        while (cap < 2 * need) cap <<= 1;
        t->cap = cap;
        t->keys = ar.alloc<sdqlrt::u64>(cap);
    }
    t->rep = ar.alloc<int>(t->cap);
    return true;
}

// Column tile ring (generated rel-scan kernels, "tma" pipeline): pick the number of stages and the CTAs per SM the
// shared memory allows.  More resident CTAs (latency hiding for probes) win over deeper rings as long as every CTA
// keeps at least two tiles in flight.  tier_smem = bytes the kernel's shared aggregation table needs in front.
struct RingPlan {
    int stages, ctas_per_sm;
    unsigned ring_off;
    size_t smem;
};
static inline bool plan_ring(size_t tier_smem, size_t stage_bytes, RingPlan* p) {
    const size_t kSmemPerSM = 232448;  // 227 KB
    const size_t off = (tier_smem + 127) & ~(size_t)127;
    const size_t stride = (stage_bytes + 127) & ~(size_t)127;
    if (!stride) return false;
    for (int cps = 4; cps >= 1; --cps) {
        const size_t per = kSmemPerSM / cps - 1536;  // 1 KB per CTA is reserved by the system, static smem < 512 B
        if (per < off + 128 + stride) continue;
        long long S = (long long)((per - off - 128) / stride);
        if (S < 2 && cps > 1) continue;
        if (S > 4) S = 4;
        p->stages = (int)S;
        p->ctas_per_sm = cps;
        p->ring_off = (unsigned)off;
        p->smem = off + 128 + (size_t)S * stride;
        return true;
    }
    return false;
}

// persistent grid for tile loops: every resident CTA slot gets a CTA, never more CTAs than tiles
static inline int grid_for_tiles(long long tiles, int ctas_per_sm, int sms) {
    long long mx = (long long)sms * (ctas_per_sm < 1 ? 1 : ctas_per_sm);
    if (tiles < 1) tiles = 1;
    return (int)(tiles < mx ? tiles : mx);
}

static inline int grid_for(long long work_items, int ctas_per_sm, int sms) {
    long long g = (work_items + sdqlrt::kBlock - 1) / sdqlrt::kBlock;
    long long mx = (long long)sms * ctas_per_sm;
    if (g < 1) g = 1;
    return (int)(g < mx ? g : mx);
}

}  // namespace sdqlhost
