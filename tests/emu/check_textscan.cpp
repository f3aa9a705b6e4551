// TEST-ONLY: runs the 32-lane code path of sdqlrt::warp_text_scan (sdqlb200_textscan.cuh) on the CPU -- one
// std::thread per lane, barrier-based shuffles and warp syncs -- and compares the candidate masks bit for bit with
// the scalar definition (row r of the run is a candidate for pattern p iff some byte position inside row r starts
// with the pattern's first four characters; bytes behind the column's end read as zero), and checks that every row
// that really contains the pattern is a candidate.  The buffer ends exactly at the column's end, in a page followed
// by an inaccessible one: any read past the column faults.
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <thread>
#include <unistd.h>
#include <vector>

#define SDQL_DEV static inline
#define TX_NOINLINE static inline
namespace sdqlrt {
typedef long long i64;
constexpr int kLanes = 32, kVec = 4, kStageRows = kLanes * kVec;
template <class T> SDQL_DEV T ld1(const T* p) { return *p; }
static std::barrier<> g_bar(32);
static thread_local int t_lane;
static unsigned g_xchg[32];
SDQL_DEV int tx_lane() { return t_lane; }
SDQL_DEV void tx_syncwarp() { g_bar.arrive_and_wait(); }
SDQL_DEV unsigned tx_shfl_down(unsigned v, int d) {
    g_xchg[t_lane] = v;
    g_bar.arrive_and_wait();
    const unsigned r = t_lane + d < 32 ? g_xchg[t_lane + d] : v;
    g_bar.arrive_and_wait();
    return r;
}
SDQL_DEV unsigned tx_shfl(unsigned v, int l) {
    g_xchg[t_lane] = v;
    g_bar.arrive_and_wait();
    const unsigned r = g_xchg[l & 31];
    g_bar.arrive_and_wait();
    return r;
}
SDQL_DEV unsigned tx_ballot(bool p) {
    g_xchg[t_lane] = p ? 1u : 0u;
    g_bar.arrive_and_wait();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= g_xchg[l] << l;
    g_bar.arrive_and_wait();
    return r;
}
SDQL_DEV int tx_ffs(unsigned v) { return __builtin_ffs((int)v); }
SDQL_DEV void tx_atomic_or(unsigned* p, unsigned v) { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
SDQL_DEV void tx_ldnc16(const unsigned char* p, unsigned (&w)[4]) {
    if ((uintptr_t)p & 15) { fprintf(stderr, "unaligned 16-byte load\n"); abort(); }
    memcpy(w, p, 16);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
#include "sdqlb200_textscan.cuh"
}  // namespace sdqlrt

using namespace sdqlrt;

static int naive_find(const unsigned char* s, int w, const char* pat, int plen) {
    int n = 0;
    while (n < w && s[n]) ++n;
    for (int i = 0; i + plen <= n; ++i)
        if (!memcmp(s + i, pat, plen)) return i;
    return -1;
}

template <int NP>
static long run_case(int W, long n, const char* const (&pats)[NP], unsigned seed, bool full_rows) {
    const size_t bytes = (size_t)n * W;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t span = ((bytes + page - 1) / page + 1) * page;
    unsigned char* map = (unsigned char*)mmap(nullptr, span + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    mprotect(map + span, page, PROT_NONE);
    // the column must start 256-byte aligned AND end right in front of the guard page: possible when bytes % 256 == 0,
    // otherwise keep the alignment (the scan's own bound checks are then verified by the scalar comparison only)
    unsigned char* col = (bytes % 256 == 0) ? map + span - bytes : map;
    srand(seed);
    const char* alpha = "specrqugCtomaln ";
    for (long r = 0; r < n; ++r) {
        unsigned char* s = col + (size_t)r * W;
        const int len = full_rows ? W : rand() % (W + 1);
        for (int i = 0; i < W; ++i) s[i] = i < len ? alpha[rand() % 16] : 0;
        if (rand() % 4 == 0) {
            const char* p = pats[rand() % NP];
            const int plen = (int)strlen(p);
            if (len >= plen) memcpy(s + rand() % (len - plen + 1), p, plen);
            else if (full_rows == false && rand() % 2) memcpy(s + W - 2 > s ? s + W - 2 : s, p, 2);  // a prefix cut by the row end
        }
        if (rand() % 8 == 0 && len > 2) s[rand() % len] = 0;  // an embedded NUL: the search ends there (a pattern behind it does not count)
    }
    unsigned pat4[NP];
    for (int p = 0; p < NP; ++p) pat4[p] = pat4_of(pats[p]);
    long bad = 0;
    std::vector<unsigned> mask(NP * kTextWords), want(NP * kTextWords);
    TextPat tps[NP];
    for (int p = 0; p < NP; ++p) {
        const int plen = (int)strlen(pats[p]);
        tps[p].plen = plen;
        for (int k = 0; k < 4; ++k) {
            tps[p].w[k] = tps[p].m[k] = 0;
            for (int j = 0; j < 4; ++j)
                if (4 * k + j < plen && 4 * k + j < 16) { tps[p].w[k] |= (unsigned)(unsigned char)pats[p][4 * k + j] << (8 * j); tps[p].m[k] |= 0xffu << (8 * j); }
        }
    }
    alignas(8) short pos[NP * kStageRows];
    for (long row0 = 0; row0 < n + kStageRows; row0 += kStageRows) {  // one extra run behind the end: r1 <= row0
        for (auto& m : mask) m = 0xdeadbeefu;
        std::vector<std::thread> th;
        for (int l = 0; l < 32; ++l)
            th.emplace_back([&, l] {
                t_lane = l;
                warp_text_scan<NP>(col, row0, n, W, pat4, mask.data());
                warp_text_resolve<NP>(col, row0, n, W, tps, mask.data(), pos);
            });
        for (auto& t : th) t.join();
        for (auto& m : want) m = 0;
        const long r1 = row0 + kStageRows < n ? row0 + kStageRows : n;
        const size_t run0 = (size_t)row0 * W, runb = r1 > row0 ? (size_t)(r1 - row0) * W : 0;
        // scalar definition: a 16-byte chunk of the run with a hit position marks every row of the run it overlaps
        for (size_t off = 0; off < runb; off += 16)
            for (int p = 0; p < NP; ++p) {
                bool hit = false;
                for (size_t q = off; q < off + 16; ++q) {
                    unsigned win = 0;
                    for (int j = 0; j < 4; ++j)
                        if (run0 + q + j < bytes) win |= (unsigned)col[run0 + q + j] << (8 * j);
                    hit = hit || win == pat4[p];
                }
                if (!hit) continue;
                const size_t last = off + 15 < runb ? off + 15 : runb - 1;
                for (size_t r = off / W; r <= last / W; ++r) want[p * kTextWords + (r >> 5)] |= 1u << (r & 31);
            }
        for (int k = 0; k < NP * kTextWords; ++k)
            if (mask[k] != want[k]) { if (bad < 5) printf("MASK MISMATCH W=%d n=%ld row0=%ld word %d: got %08x want %08x\n", W, n, row0, k, mask[k], want[k]); ++bad; }
        // the aligned-word variant (patterns of >= 7 characters): same bits as its scalar definition, no matching row missed
        bool all7 = true;
        for (int p = 0; p < NP; ++p) all7 = all7 && strlen(pats[p]) >= 7;
        if (all7) {
            unsigned pw[NP][4];
            for (int p = 0; p < NP; ++p)
                for (int o = 0; o < 4; ++o) pw[p][o] = pat4_of(pats[p] + o);
            std::vector<unsigned> am(NP * kTextWords, 0xdeadbeefu), aw(NP * kTextWords, 0u);
            std::vector<std::thread> th2;
            for (int l = 0; l < 32; ++l)
                th2.emplace_back([&, l] { t_lane = l; warp_text_scan_aligned<NP>(col, row0, n, W, pw, am.data()); });
            for (auto& t : th2) t.join();
            for (size_t off = 0; off < runb; off += 16)  // row0 * W is a multiple of 16: chunks and words are aligned in the column
                for (int p = 0; p < NP; ++p) {
                    bool hit = false;
                    for (size_t q = off; q < off + 16; q += 4) {
                        unsigned x = 0;
                        for (int j = 0; j < 4; ++j)
                            if (run0 + q + j < bytes) x |= (unsigned)col[run0 + q + j] << (8 * j);
                        for (int o = 0; o < 4; ++o) hit = hit || x == pw[p][o];
                    }
                    if (!hit) continue;
                    const size_t last = off + 15 < runb ? off + 15 : runb - 1;
                    for (size_t r = off / W; r <= last / W; ++r) aw[p * kTextWords + (r >> 5)] |= 1u << (r & 31);
                }
            for (int k = 0; k < NP * kTextWords; ++k)
                if (am[k] != aw[k]) { if (bad < 5) printf("ALIGNED MASK MISMATCH W=%d n=%ld row0=%ld word %d: got %08x want %08x\n", W, n, row0, k, am[k], aw[k]); ++bad; }
            for (long r = row0; r < r1; ++r)
                for (int p = 0; p < NP; ++p)
                    if (naive_find(col + (size_t)r * W, W, pats[p], (int)strlen(pats[p])) >= 0 &&
                        !((am[p * kTextWords + ((r - row0) >> 5)] >> ((r - row0) & 31)) & 1u)) {
                        if (bad < 5) printf("ALIGNED MISSED MATCH W=%d row %ld pat %s\n", W, r, pats[p]);
                        ++bad;
                    }
        }
        for (long r = row0; r < r1; ++r)
            for (int p = 0; p < NP; ++p) {
                const bool found = naive_find(col + (size_t)r * W, W, pats[p], (int)strlen(pats[p])) >= 0;
                const bool cand = (mask[p * kTextWords + ((r - row0) >> 5)] >> ((r - row0) & 31)) & 1u;
                if (found && !cand) { if (bad < 5) printf("MISSED MATCH W=%d row %ld pat %s\n", W, r, pats[p]); ++bad; }
                // the warp-cooperative exact search: the reference's firstIndex for every row of the run
                const int wantpos = naive_find(col + (size_t)r * W, W, pats[p], (int)strlen(pats[p]));
                if (pos[p * kStageRows + (r - row0)] != wantpos) {
                    if (bad < 5) printf("POSITION MISMATCH W=%d row %ld pat %s: got %d want %d\n", W, r, pats[p], pos[p * kStageRows + (r - row0)], wantpos);
                    ++bad;
                }
            }
    }
    munmap(map, span + page);
    return bad;
}

int main() {
    long bad = 0, cases = 0;
    const char* const two[2] = {"special", "requests"};
    const char* const one[1] = {"green"};
    const char* const three[3] = {"Customer", "Complaints", "spec"};
    const int widths[] = {79, 55, 101, 16, 5, 25, 4, 7, 128, 117};
    const long rows[] = {0, 1, 3, 127, 128, 129, 256, 300, 1000, 2048};
    for (int W : widths)
        for (long n : rows)
            for (int full = 0; full < 2; ++full) {
                bad += run_case<2>(W, n, two, 1000u + W * 31 + (unsigned)n, full);
                bad += run_case<1>(W, n, one, 2000u + W * 17 + (unsigned)n, full);
                bad += run_case<3>(W, n, three, 3000u + W * 13 + (unsigned)n, full);
                cases += 3;
            }
    printf("%ld cases, %ld mismatches\n", cases, bad);
    return bad != 0;
}
