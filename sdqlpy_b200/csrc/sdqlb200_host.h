// Host-side helpers for generated query modules: workspace arena, run-time table sizing (direct-indexed vs
// hash, from column statistics), error plumbing.  Replaces the container declarations the reference emits up
// front (outputDictionariesInitializationCode, sdql_ir_cpp_generator_par.py:205, 267, 309, 347, 418).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

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
        if (e_ != cudaSuccess) {                                                                         \
            cudaGetLastError(); /* reported here: a later cudaGetLastError() must not see it again */    \
            return sdqlhost::fail(SDQLB200_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
        }                                                                                                \
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

// One dictionary's place in the workspace: the region reserved for it by the dry run (worst case: every source row
// produces a distinct key) and the arrays currently laid out at the start of that region.
struct TblRegion {
    unsigned long long off, len;        // reserved bytes in the arena
    unsigned long long ff_len;          // [off, off + ff_len): keys + rep, to be filled with 0xFF before the build
    unsigned long long z_len;           // [off + ff_len, off + ff_len + z_len): aggregate arrays, to be zeroed
    long double dom;                    // size of the packed key domain (9e18 = unbounded)
    int nf;                             // aggregate fields (8 bytes per slot each)
    unsigned long long bdom, bmod;      // presence filter: bits (0 = none) and modulus of the packed key (0 = whole key)
};

// a table smaller than this needs no presence filter in front (it is cache resident itself)
static inline unsigned long long bits_min_bytes() {
    static long long v = -1;
    if (v < 0) { const char* e = getenv("SDQLB200_BITS_MIN_BYTES"); v = e ? atoll(e) : (4ll << 20); if (v < 0) v = 0; }
    return (unsigned long long)v;
}

static inline unsigned long long up256(unsigned long long b) { return (b + 255) & ~255ull; }

// strided-dense key columns (sdqlb200_col.stride): block size 2^sb, residues used sk, and the number of packed values
static inline int col_sb(const sdqlb200_col& c) { return (c.stride >> 16) & 0xff; }
static inline int col_sk(const sdqlb200_col& c) { return c.stride & 0xffff; }
static inline long long col_range(const sdqlb200_col& c) {
    const int sb = col_sb(c);
    if (!sb || col_sk(c) < 1) return c.max - c.min + 1;
    return (((c.max - c.min) >> sb) + 1) * (long long)col_sk(c);
}

// direct vs hash for `rows` distinct keys at most.  Direct when the dense array is not (much) bigger than what a hash
// table for `rows` keys would need.  SDQLB200_FORCE_HASH=1 (debug knob): never direct -- exercises the hash build /
// probe / merge paths at small scale.
static inline void plan_table(long double dom, long long rows, int* direct, long long* cap, long long sparse = 8) {
    static const bool force_hash = getenv("SDQLB200_FORCE_HASH") && getenv("SDQLB200_FORCE_HASH")[0] == '1';
    if (rows < 1) rows = 1;
    const long long domain = (long long)dom;
    *direct = (!force_hash && dom <= (long double)(sparse * rows + 65536) && domain < (1ll << 31)) ? 1 : 0;
    if (*direct) { *cap = domain; return; }
    const long long need = rows < domain ? rows : domain;
    long long c = 1024;
    while (c < 2 * need) c <<= 1;
    *cap = c;
}
static inline unsigned long long table_bytes(int direct, long long cap, int nf) {
    const unsigned long long n = (unsigned long long)(cap > 0 ? cap : 1);
    return (direct ? 0 : up256(n * 8)) + up256(n * 4) + (unsigned long long)nf * up256(n * 8);
}
static inline unsigned long long bits_bytes(unsigned long long bdom) { return bdom ? up256(((bdom + 31) / 32) * 4) : 0; }
// lay the table's arrays out at the start of its region: [keys (hash only)] [rep] [agg 0] .. [agg nf-1]
static inline void place_table(sdqlrt::Tbl* t, char* base, TblRegion* r, void** aggs) {
    const unsigned long long n = (unsigned long long)(t->cap > 0 ? t->cap : 1);
    unsigned long long o = r->off;
    t->keys = nullptr;
    if (!t->direct) { t->keys = base ? (sdqlrt::u64*)(base + o) : nullptr; o += up256(n * 8); }
    t->rep = base ? (int*)(base + o) : nullptr;
    o += up256(n * 4);
    r->ff_len = o - r->off;
    for (int j = 0; j < r->nf; ++j) { aggs[j] = base ? (void*)(base + o) : nullptr; o += up256(n * 8); }
    t->bits = nullptr;
    t->bmod = r->bmod;
    // presence bits in front of big tables (the bitmap stays cache resident) and of every hashed table (a bit test
    // is far cheaper than hashing + probing, however small the table)
    if (r->bdom && (!t->direct || o - r->off >= bits_min_bytes())) {
        t->bits = base ? (unsigned*)(base + o) : nullptr;
        o += bits_bytes(r->bdom);
    }
    r->z_len = o - r->off - r->ff_len;
}

// Decide direct vs hash and reserve the table's region.  parts: value ranges [mn, mn+rng) of the by-value key parts.
// src_rows bounds the number of distinct keys.  Returns false if the key domain cannot be packed in 63 bits.
static inline bool size_table(sdqlrt::Tbl* t, int nparts, const long long* mn, const long long* rng,
                              long long src_rows, long long* o_mn, long long* o_rng, long long* o_mul, Arena& ar,
                              TblRegion* r, int nf, void** aggs, bool want_bits, bool prefix_bits = false) {
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
    plan_table(dom, src_rows, &t->direct, &t->cap);
    r->dom = dom;
    r->nf = nf;
    r->bdom = r->bmod = 0;
    if (want_bits) {  // one bit per packed key if that is at most 2^31 bits (256 MB), else one per value of the first part
        const long double lim = 2147483648.0L;
        // SDQLB200_BITS_PREFIX=1 (debug knob): composite keys always get the first-part filter (exercised at small scale)
        static const bool force_prefix = getenv("SDQLB200_BITS_PREFIX") && getenv("SDQLB200_BITS_PREFIX")[0] == '1';
        // prefix_bits: the table is a dictionary of dictionaries probed by its outer key part (one test rejects the whole inner
        // dictionary, and the per-part bitmap is `inner range` times smaller -- cache resident where the full one is not)
        if (dom <= lim && !((force_prefix || prefix_bits) && nparts > 1)) r->bdom = (unsigned long long)dom;
        else if (nparts > 1 && o_rng[0] > 0 && (long double)o_rng[0] <= lim) { r->bdom = (unsigned long long)o_rng[0]; r->bmod = r->bdom; }
    }
    r->off = ar.used;
    r->len = table_bytes(t->direct, t->cap, nf) + bits_bytes(r->bdom);
    ar.alloc<char>((long long)r->len);
    place_table(t, ar.ok() ? ar.base : nullptr, r, aggs);
    return true;
}

// Right-size a table once the number of rows that reach its build is known (cardinality pass): same region, smaller
// (cache-resident) arrays.  Keeps the worst-case plan when the new one would not fit the reservation.
static inline bool replan_table(sdqlrt::Tbl* t, char* base, TblRegion* r, long long rows, void** aggs, long long sparse = 64) {
    int direct;
    long long cap;
    // with the true cardinality known, a dense array stays the better table down to one key per 64 slots: sorted
    // builds and probes stream through it, while a hash table of millions of keys is built with random CAS traffic.
    // (A table that is merged across ranks passes sparse = 4: a dense table is merged by all-reducing the WHOLE array --
    // Q17's 20 M-slot table for 20 K parts moved 400 MB per merge -- a hashed one by shuffling its entries.)
    plan_table(r->dom, rows, &direct, &cap, sparse);
    if (table_bytes(direct, cap, r->nf) + bits_bytes(r->bdom) > r->len) return false;
    if (direct == t->direct && cap == t->cap) return false;
    t->direct = direct;
    t->cap = cap;
    place_table(t, base, r, aggs);
    return true;
}

// A single-part table whose key is a value read out of another dictionary has no statistics when the workspace is laid out
// (planned as a hash table of raw 64-bit keys).  Once that dictionary is complete its value range [lo, lo + rng) is known:
// a small range turns the table into a direct-indexed one in the same region (and lets a group-by take the register /
// shared-memory tiers).  Keeps the hash plan when the range is large or the dense layout would not fit the reservation.
static inline bool redomain_table(sdqlrt::Tbl* t, char* base, TblRegion* r, long long lo, long long rng, long long* o_mn,
                                  long long* o_rng, long long* o_mul, void** aggs) {
    if (rng < 1 || rng > 65536 || table_bytes(1, rng, r->nf) > r->len) return false;
    o_mn[0] = lo; o_rng[0] = rng; o_mul[0] = 1;
    r->dom = (long double)rng;
    r->bdom = r->bmod = 0;
    t->direct = 1;
    t->cap = rng;
    place_table(t, base, r, aggs);
    return true;
}

// tables whose reservation is at least this big get a cardinality pass before they are built
static inline unsigned long long count_min_bytes() {
    static long long v = -1;
    if (v < 0) { const char* e = getenv("SDQLB200_COUNT_MIN_BYTES"); v = e ? atoll(e) : (8ll << 20); if (v < 0) v = 0; }
    return (unsigned long long)v;
}

// ... and whose reservation is at least this many times the bytes the cardinality pass has to scan
static inline unsigned long long count_min_ratio() {
    static long long v = -1;
    if (v < 0) { const char* e = getenv("SDQLB200_COUNT_MIN_RATIO"); v = e ? atoll(e) : 4; if (v < 0) v = 0; }
    return (unsigned long long)v;
}
// ... or is simply huge: initialising and then randomly touching several GB costs more than a second pass over the
// predicate columns whatever the ratio (Q12 at SF100: 12.6 GB direct table for 3 M entries, 8.3 ms; counted: 4.7 ms;
// lowering the ratio for everybody instead cost Q7 / Q9 / Q20 0.3-0.9 ms each, profiles/r02_visit8)
static inline unsigned long long count_big_bytes() {
    static long long v = -1;
    if (v < 0) { const char* e = getenv("SDQLB200_COUNT_BIG_BYTES"); v = e ? atoll(e) : (4ll << 30); if (v < 0) v = 0; }
    return (unsigned long long)v;
}
// direct tables of at most this many slots are merged across ranks with ONE fused all-reduce (presence + all fields);
// SDQLB200_FUSED_MERGE_MAX lowers it so that small-scale tests reach the large-table merges (sparse exchange / dense)
static inline long long fused_merge_max() {
    static long long v = -1;
    if (v < 0) { const char* e = getenv("SDQLB200_FUSED_MERGE_MAX"); v = e ? atoll(e) : sdqlrt::kFusedMergeMaxSlots; if (v < 0 || v > sdqlrt::kFusedMergeMaxSlots) v = sdqlrt::kFusedMergeMaxSlots; }
    return v;
}
static inline bool debug() {
    static const bool d = getenv("SDQLB200_DEBUG") && getenv("SDQLB200_DEBUG")[0] == '1';
    return d;
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
