// sdqlpy-b200 device runtime, part: warp text scan (included by sdqlb200_rt.cuh inside namespace sdqlrt).
// Kept in a file of its own so that tests/emu/check_textscan.cpp can compile the 32-lane code path on the CPU (one
// std::thread per lane, barrier-based shuffles) against the scalar definition.  Needs from the includer: SDQL_DEV, i64,
// kStageRows, ld1<T>(), and -- unless SDQLB200_EMU -- TX_NOINLINE, tx_lane(), tx_syncwarp(), tx_shfl_down(),
// tx_atomic_or(), tx_ldnc16().
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
// 16 bytes at src + off as four little-endian words; bytes at or behind `total` read as zero (no access past the column)
SDQL_DEV void text_load16(const unsigned char* src, size_t off, size_t total, unsigned (&w)[4]) {
    if (off + 16 <= total) {
        tx_ldnc16(src + off, w);
    } else {
        w[0] = w[1] = w[2] = w[3] = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (off + j < total) w[j >> 2] |= (unsigned)ld1(src + off + j) << (8 * (j & 3));
    }
}
SDQL_DEV unsigned text_load4(const unsigned char* src, size_t off, size_t total) {
    if (off + 4 <= total) return ld1((const unsigned*)(src + off));
    unsigned v = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (off + j < total) v |= (unsigned)ld1(src + off + j) << (8 * j);
    return v;
}
#endif
#ifndef SDQLB200_EMU
TX_NOINLINE void text_mark(unsigned* mask, size_t row) { tx_atomic_or(mask + (row >> 5), 1u << (unsigned)(row & 31u)); }
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
    const size_t bytes = (size_t)(r1 - row0) * (size_t)W;       // this run
    const size_t total = (size_t)(n - row0) * (size_t)W;        // readable bytes from the run's start to the column's end
    const unsigned char* src = col + row0 * W;                   // 16-byte aligned: row0 is a multiple of 128, the base of 256
    const size_t nv = (bytes + 15) >> 4;
    // software pipeline: the next step's chunk is requested before this one is examined.  The three bytes behind a
    // chunk are the start of the next lane's chunk (one shuffle); the last lane fetches its own (a 4-byte load out of
    // the line the next step's first lane reads anyway).
    unsigned nx[4], nx4 = 0u;
    text_load16(src, (size_t)lane << 4, total, nx);
    if (lane == 31) nx4 = text_load4(src, ((size_t)lane << 4) + 16, total);
    for (size_t kb = 0; kb < nv; kb += 32) {
        const size_t off = (kb + lane) << 4;
        unsigned w[5] = {nx[0], nx[1], nx[2], nx[3], nx4};
        if (kb + 32 < nv) {
            text_load16(src, off + 512, total, nx);
            if (lane == 31) nx4 = text_load4(src, off + 528, total);
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
            // rare (a few percent of the chunks): find the positions again, one word at a time in a rolled loop -- the
            // code stays small (the fully unrolled form made the kernel miss the instruction cache)
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const unsigned lo = j == 0 ? w[0] : j == 1 ? w[1] : j == 2 ? w[2] : w[3];
                const unsigned hi = j == 0 ? w[1] : j == 1 ? w[2] : j == 2 ? w[3] : w[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const unsigned win = a ? __funnelshift_r(lo, hi, 8 * a) : lo;
                    if (win == pat4[p] && off + 4 * j + a < bytes) text_mark(mask + p * kTextWords, (off + 4 * j + a) / (size_t)W);
                }
            }
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
// candidate bit of row u (0 .. kVec-1) of this lane for pattern p
SDQL_DEV bool text_cand(const unsigned* mask, int p, int u) {
#ifndef SDQLB200_EMU
    const unsigned r = ((unsigned)tx_lane() << 2) + (unsigned)u;
#else
    const unsigned r = (unsigned)u;
#endif
    return (mask[p * kTextWords + (r >> 5)] >> (r & 31u)) & 1u;
}

