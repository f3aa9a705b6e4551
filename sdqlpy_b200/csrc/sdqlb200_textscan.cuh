// sdqlpy-b200 device runtime, part: warp text scan (included by sdqlb200_rt.cuh inside namespace sdqlrt).
// Kept in a file of its own so that tests/emu/check_textscan.cpp can compile the 32-lane code path on the CPU (one
// std::thread per lane, barrier-based shuffles) against the scalar definition.  Needs from the includer: SDQL_DEV, i64,
// kStageRows, ld1<T>(), and -- unless SDQLB200_EMU -- TX_NOINLINE, tx_lane(), tx_syncwarp(), tx_shfl_down(), tx_shfl(),
// tx_ballot(), tx_ffs(), tx_atomic_or(), tx_ldnc16().
// ---------------------------------------------------------------------------------------------
// warp text scan: candidate rows for firstIndex / contains on a scanned string column.
// The kStageRows rows a warp examines per iteration are one contiguous run of bytes.  All lanes stream the run with
// coalesced 128-bit loads and compare the four bytes at EVERY byte position with the first four characters of each of
// the NP patterns (one shift + one 32-bit compare per position and pattern), without regard to row boundaries or NULs.
// A row that contains a pattern has those four characters at some position inside the row, so the rows that can match
// are a subset of the rows hit; the exact per-row search (str_find, reference semantics) then runs for those rows only
// and every other row's firstIndex is -1.  mask[p * kTextWords + w] bit b: row 32 * w + b of the run may contain
// pattern p.  All lanes of the warp must call; the masks may be read after the call returns.
// The instruction count per byte is ~2 per lane (per pattern ~1 more), against ~30 for the per-row search, so the scan
// runs at the speed the bytes arrive from HBM.
// ---------------------------------------------------------------------------------------------
constexpr int kTextWords = (kStageRows + 31) / 32;  // mask words per pattern
SDQL_DEV unsigned pat4_of(const char* pat) {  // first four characters as one little-endian word (patterns have >= 4)
    return (unsigned)(unsigned char)pat[0] | ((unsigned)(unsigned char)pat[1] << 8) |
           ((unsigned)(unsigned char)pat[2] << 16) | ((unsigned)(unsigned char)pat[3] << 24);
}
#ifndef SDQLB200_EMU
// 16 bytes at src + off as four little-endian words; bytes at or behind `total` read as zero (no access past the column).
// Offsets are run-local (a run is at most kStageRows rows: 32-bit arithmetic in the scan loop).
SDQL_DEV void text_load16(const unsigned char* src, unsigned off, unsigned total, unsigned (&w)[4]) {
    if (off + 16u <= total) {
        tx_ldnc16(src + off, w);
    } else {
        w[0] = w[1] = w[2] = w[3] = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (off + j < total) w[j >> 2] |= (unsigned)ld1(src + off + j) << (8 * (j & 3));
    }
}
SDQL_DEV unsigned text_load4(const unsigned char* src, unsigned off, unsigned total) {
    if (off + 4u <= total) return ld1((const unsigned*)(src + off));
    unsigned v = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (off + j < total) v |= (unsigned)ld1(src + off + j) << (8 * j);
    return v;
}
// readable bytes from the run's start, clamped to what the scan can touch (its own bytes + the 3 chunks it requests ahead)
SDQL_DEV unsigned text_total(i64 n, i64 row0, int W, unsigned bytes) {
    const size_t t = (size_t)(n - row0) * (size_t)W;
    return t > (size_t)bytes + 4096u ? bytes + 4096u : (unsigned)t;
}
#endif
#ifndef SDQLB200_EMU
// marks the rows of the run that the 16-byte chunk at byte `off` overlaps (two at most when W >= 16)
SDQL_DEV void text_mark_chunk(unsigned* mask, unsigned off, unsigned bytes, int W) {
    const unsigned r0 = off / (unsigned)W;
    const unsigned last = off + 15u < bytes ? off + 15u : bytes - 1u;
    const unsigned r1 = last / (unsigned)W;
    for (unsigned r = r0; r <= r1; ++r) tx_atomic_or(mask + (r >> 5), 1u << (r & 31u));
}
#endif
template <int NP>
SDQL_DEV void warp_text_scan(const unsigned char* col, i64 row0, i64 n, int W, const unsigned (&pat4)[NP], unsigned* mask) {
    static_assert(NP >= 1 && NP * kTextWords <= 32, "at most 8 patterns per column");
    const i64 r1 = row0 + kStageRows < n ? row0 + kStageRows : n;
#ifndef SDQLB200_EMU
    const int lane = tx_lane();
    tx_syncwarp();  // the previous iteration's readers are done with the masks
    if (lane < NP * kTextWords) mask[lane] = 0u;
    tx_syncwarp();
    if (r1 <= row0) return;  // the same decision in every lane
    const unsigned bytes = (unsigned)(r1 - row0) * (unsigned)W;  // this run
    const unsigned total = text_total(n, row0, W, bytes);        // readable bytes from the run's start (clamped)
    const unsigned char* src = col + row0 * W;                   // 16-byte aligned: row0 is a multiple of 128, the base of 256
    const unsigned nv = (bytes + 15u) >> 4;
    // software pipeline: two steps' chunks are in flight while one is examined (one was not enough to cover the HBM latency:
    // 23 % of q13_k0's stall samples sat on the chunk's first use).  The three bytes behind a chunk are the start of the next
    // lane's chunk (one shuffle); the last lane fetches its own (a 4-byte load out of the line the next step's first lane
    // reads anyway).
    unsigned nx[4], ny[4], nx4 = 0u, ny4 = 0u;
    text_load16(src, (unsigned)lane << 4, total, nx);
    text_load16(src, ((unsigned)lane << 4) + 512u, total, ny);
    if (lane == 31) { nx4 = text_load4(src, ((unsigned)lane << 4) + 16u, total); ny4 = text_load4(src, ((unsigned)lane << 4) + 528u, total); }
    for (unsigned kb = 0; kb < nv; kb += 32) {
        const unsigned off = (kb + lane) << 4;
        unsigned w[5] = {nx[0], nx[1], nx[2], nx[3], nx4};
#pragma unroll
        for (int j = 0; j < 4; ++j) nx[j] = ny[j];
        nx4 = ny4;
        if (kb + 64 < nv) {
            text_load16(src, off + 1024u, total, ny);
            if (lane == 31) ny4 = text_load4(src, off + 1040u, total);
        }
        const unsigned dn = tx_shfl_down(w[0], 1);
        if (lane != 31) w[4] = dn;
        if (off >= bytes) continue;  // behind the shuffle: every lane takes part in it
        bool hit[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) hit[p] = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const unsigned win = a ? __funnelshift_r(w[j], w[j + 1], 8 * a) : w[j];
#pragma unroll
                for (int p = 0; p < NP; ++p) hit[p] |= (win == pat4[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (!hit[p]) continue;
            // a hit (a few percent of the chunks; ~a quarter of the warp steps in Q13): mark the rows this 16-byte chunk overlaps
            // -- a superset of the rows hit, the resolve behind the scan is exact.  (Finding the hit position again, word by
            // word, was a third of q13_k0's instructions: one lane in a rolled loop, 31 waiting.)
            text_mark_chunk(mask + p * kTextWords, off, bytes, W);
        }
    }
    tx_syncwarp();
#else
    for (int k = 0; k < NP * kTextWords; ++k) mask[k] = 0u;
    if (r1 <= row0) return;
    const size_t bytes = (size_t)(r1 - row0) * (size_t)W, total = (size_t)(n - row0) * (size_t)W;
    const unsigned char* src = col + row0 * W;
    for (size_t q = 0; q < bytes; ++q) {
        unsigned win = 0;
        for (int j = 0; j < 4; ++j)
            if (q + j < total) win |= (unsigned)src[q + j] << (8 * j);
        for (int p = 0; p < NP; ++p)
            if (win == pat4[p]) { const unsigned r = (unsigned)(q / (size_t)W); mask[p * kTextWords + (r >> 5)] |= 1u << (r & 31u); }
    }
#endif
}
// ---------------------------------------------------------------------------------------------
// aligned-word variant of the scan for patterns of >= 7 characters: an occurrence at byte q covers the 4-byte aligned word
// at ceil(q / 4) * 4 entirely, and that word then equals the pattern's characters [o, o + 4) for o = 0 .. 3.  So the words of
// the run are compared as they are loaded (4 constants per pattern, 16 compares per 16 bytes and pattern) -- no byte
// windows (12 funnel shifts per 16 bytes), no neighbour word, no shuffle.  A word belongs to the row of its first byte; an
// occurrence inside a row has its aligned word inside that row, so no matching row is missed.  pw[p][o]: characters
// o .. o+3 of pattern p.  The run starts at a multiple of 4 bytes of the column (row0 is a multiple of kStageRows).
// ---------------------------------------------------------------------------------------------
template <int NP>
SDQL_DEV void warp_text_scan_aligned(const unsigned char* col, i64 row0, i64 n, int W, const unsigned (&pw)[NP][4], unsigned* mask) {
    static_assert(NP >= 1 && NP * kTextWords <= 32, "at most 8 patterns per column");
    const i64 r1 = row0 + kStageRows < n ? row0 + kStageRows : n;
#ifndef SDQLB200_EMU
    const int lane = tx_lane();
    tx_syncwarp();
    if (lane < NP * kTextWords) mask[lane] = 0u;
    tx_syncwarp();
    if (r1 <= row0) return;
    const unsigned bytes = (unsigned)(r1 - row0) * (unsigned)W;
    const unsigned total = text_total(n, row0, W, bytes);
    const unsigned char* src = col + row0 * W;
    const unsigned nv = (bytes + 15u) >> 4;
    unsigned nx[4], ny[4];  // two steps' chunks in flight
    text_load16(src, (unsigned)lane << 4, total, nx);
    text_load16(src, ((unsigned)lane << 4) + 512u, total, ny);
    for (unsigned kb = 0; kb < nv; kb += 32) {
        const unsigned off = (kb + lane) << 4;
        const unsigned w[4] = {nx[0], nx[1], nx[2], nx[3]};
#pragma unroll
        for (int j = 0; j < 4; ++j) nx[j] = ny[j];
        if (kb + 64 < nv) text_load16(src, off + 1024u, total, ny);
        if (off >= bytes) continue;
        bool hit[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) hit[p] = false;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int p = 0; p < NP; ++p) hit[p] |= (w[j] == pw[p][o]);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (hit[p]) text_mark_chunk(mask + p * kTextWords, off, bytes, W);
        }
    }
    tx_syncwarp();
#else
    for (int k = 0; k < NP * kTextWords; ++k) mask[k] = 0u;
    if (r1 <= row0) return;
    const size_t bytes = (size_t)(r1 - row0) * (size_t)W, total = (size_t)(n - row0) * (size_t)W;
    const unsigned char* src = col + row0 * W;
    const size_t lead = (4 - (size_t)(row0 * W) % 4) % 4;  // the emulated run may start anywhere: words aligned in the COLUMN
    for (size_t q = lead; q < bytes; q += 4) {
        unsigned x = 0;
        for (int j = 0; j < 4; ++j)
            if (q + j < total) x |= (unsigned)src[q + j] << (8 * j);
        for (int p = 0; p < NP; ++p)
            if (x == pw[p][0] || x == pw[p][1] || x == pw[p][2] || x == pw[p][3]) {
                const unsigned r = (unsigned)(q / (size_t)W);
                mask[p * kTextWords + (r >> 5)] |= 1u << (r & 31u);
            }
    }
#endif
}
// ---------------------------------------------------------------------------------------------
// warp text resolve: the exact firstIndex (varchar.h:91-97: wcsstr, the search ends at the row's first NUL) of every
// pattern in every candidate row of the run, computed by the WHOLE warp, one candidate row at a time, instead of by the one
// lane that owns the row.  The per-lane search was 60 % of q13_k0's instructions: ~5 of a warp's 128 rows are candidates,
// so nearly every iteration sent a few lanes through a ~300-instruction branchy search while the other lanes waited.
// Lane l examines the start positions 4l .. 4l+3 of the row: its own four characters (one aligned 32-bit load + the next
// lane's word, shifted into place; zero behind the row's end) plus the following words of the next lanes (shuffles),
// compared with each pattern that marked the row as 1 .. 4 masked words; ballots pick the first match per pattern and the
// first NUL (a match counts only in front of it).  No divergence; rows longer than 4 * (32 - 4) characters take several
// steps.  Patterns are at most 16 characters (longer ones keep the per-lane search).
// pos[p * kStageRows + r]: firstIndex of pattern p in row r of the run, -1 = absent.  All lanes must call.
// ---------------------------------------------------------------------------------------------
struct TextPat { unsigned w[4], m[4]; int plen; };
SDQL_DEV unsigned tx_zero_bytes(unsigned x) { return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu); }  // 0x80 per zero byte
#ifndef SDQLB200_EMU
// pm: bit p set = pattern p marked this row.  Writes out[p * kStageRows] for those patterns (lane 0).
template <int NP>
SDQL_DEV void warp_row_find(const unsigned char* s, int W, const TextPat (&pats)[NP], unsigned pm, short* out) {
    const int lane = tx_lane();
    constexpr int kJudge = 32 - 4;  // lanes whose four start positions have all their (<= 16 + 3) characters inside a step
    const unsigned a = (unsigned)(size_t)s & 3u;
    const unsigned* wp = reinterpret_cast<const unsigned*>(s - a);
    int res[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) res[p] = -1;
    for (int base = 0; base < W && pm; base += 4 * kJudge) {
        // characters o .. o+3 of the row: aligned word wi covers characters 4 wi - a .. 4 wi - a + 3
        const int wi = (base >> 2) + lane, o = base + 4 * lane;
        unsigned x = (4 * wi - (int)a < W) ? ld1(wp + wi) : 0u;
        unsigned xn = tx_shfl_down(x, 1);
        if (lane == 31) xn = (a && 4 * (wi + 1) - (int)a < W) ? ld1(wp + wi + 1) : 0u;
        unsigned v[5];
        v[0] = a ? __funnelshift_r(x, xn, 8u * a) : x;
        if (o + 4 > W) v[0] = o < W ? (v[0] & ((1u << (8 * (W - o))) - 1u)) : 0u;  // behind the row's end: NUL
#pragma unroll
        for (int k = 1; k <= 4; ++k) v[k] = tx_shfl_down(v[0], k);
        const unsigned z = tx_zero_bytes(v[0]);
        const unsigned mz = tx_ballot(z != 0u);
        int nul = 0x7fffffff;
        if (mz) { const int L = tx_ffs(mz) - 1; nul = base + 4 * L + ((tx_ffs(tx_shfl(z, L)) - 1) >> 3); }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (!((pm >> p) & 1u)) continue;  // warp-uniform
            const int nw = (pats[p].plen + 3) >> 2;
            unsigned hit = 0u;  // bit j: the pattern starts at o + j
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k >= nw) continue;
                    const unsigned win = j ? __funnelshift_r(v[k], v[k + 1], 8 * j) : v[k];
                    ok = ok && ((win & pats[p].m[k]) == pats[p].w[k]);
                }
                hit |= (ok ? 1u : 0u) << j;
            }
            if (lane >= kJudge) hit = 0u;
            const unsigned mh = tx_ballot(hit != 0u);
            int at = 0x7fffffff;
            if (mh) { const int L = tx_ffs(mh) - 1; at = base + 4 * L + tx_ffs(tx_shfl(hit, L)) - 1; }
            if (at < nul) { res[p] = at; pm &= ~(1u << p); }                 // a match cannot contain a NUL: it lies in front of it
            else if (nul < base + 4 * kJudge) pm &= ~(1u << p);              // the string ends inside the judged positions: absent
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int p = 0; p < NP; ++p) out[p * kStageRows] = (short)res[p];
    }
}
#endif
template <int NP>
SDQL_DEV void warp_text_resolve(const unsigned char* col, i64 row0, i64 n, int W, const TextPat (&pats)[NP], const unsigned* mask,
                                short* pos) {
#ifndef SDQLB200_EMU
    const int lane = tx_lane();
#pragma unroll
    for (int p = 0; p < NP; ++p) reinterpret_cast<unsigned long long*>(pos + p * kStageRows)[lane] = ~0ull;  // this lane's 4 rows: -1
    tx_syncwarp();
#pragma unroll 1
    for (int w = 0; w < kTextWords; ++w) {
        unsigned mp[NP], many = 0u;  // the same words in every lane
#pragma unroll
        for (int p = 0; p < NP; ++p) { mp[p] = pats[p].plen <= 16 ? mask[p * kTextWords + w] : 0u; many |= mp[p]; }
        while (many) {
            const int b = tx_ffs(many) - 1;
            many &= many - 1u;
            unsigned pm = 0u;
#pragma unroll
            for (int p = 0; p < NP; ++p) pm |= ((mp[p] >> b) & 1u) << p;
            const int r = w * 32 + b;
            warp_row_find<NP>(col + (row0 + r) * (i64)W, W, pats, pm, pos + r);
        }
    }
    tx_syncwarp();
#else
    const i64 r1 = row0 + kStageRows < n ? row0 + kStageRows : n;
    for (int p = 0; p < NP; ++p)
        for (int r = 0; r < kVec; ++r) {  // the emulated "lane" owns the first kVec rows of its run (text_pos below)
            pos[p * kStageRows + r] = -1;
            if (pats[p].plen > 16 || row0 + r >= r1 || !((mask[p * kTextWords + (r >> 5)] >> (r & 31)) & 1u)) continue;
            const unsigned char* s = col + (row0 + r) * (i64)W;
            int len = 0;
            while (len < W && s[len]) ++len;
            char pat[17];
            for (int j = 0; j < pats[p].plen; ++j) pat[j] = (char)(pats[p].w[j >> 2] >> (8 * (j & 3)));
            for (int i = 0; i + pats[p].plen <= len; ++i) {
                int k = 0;
                while (k < pats[p].plen && s[i + k] == (unsigned char)pat[k]) ++k;
                if (k == pats[p].plen) { pos[p * kStageRows + r] = (short)i; break; }
            }
        }
#endif
}
// exact firstIndex of pattern p in row u (0 .. kVec-1) of this lane
SDQL_DEV int text_pos(const short* pos, int p, int u) {
#ifndef SDQLB200_EMU
    return pos[p * kStageRows + (tx_lane() << 2) + u];
#else
    return pos[p * kStageRows + u];
#endif
}
// candidate bit of row u (0 .. kVec-1) of this lane for pattern p
SDQL_DEV bool text_cand(const unsigned* mask, int p, int u) {
#ifndef SDQLB200_EMU
    const unsigned r = ((unsigned)tx_lane() << 2) + (unsigned)u;
#else
    const unsigned r = (unsigned)u;
#endif
    return (mask[p * kTextWords + (r >> 5)] >> (r & 31u)) & 1u;
}

