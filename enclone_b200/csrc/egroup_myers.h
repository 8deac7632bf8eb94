// egroup_myers.h — the bit-parallel Levenshtein distance of egroup.cu, host- and device-compilable so that the CPU test
// suite can fuzz the very arithmetic the kernel runs (tests/myers_fuzz.cpp) against the textbook DP.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define EG_HD __host__ __device__
#else
#define EG_HD
#endif
#define MAXW 8                     // 64-row blocks per pattern: patterns up to 512 symbols

// symbol of a byte: nucleotides ACGT N - -> 0..5 (anything else 6); amino acids and '*' -> byte & 31
EG_HD inline uint32_t sym_nt(uint8_t b) {
    switch (b) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; case '-': return 5; default: return 6; }
}
EG_HD inline uint32_t sym_aa(uint8_t b) { return b & 31u; }

// Match vectors of a pattern: peq[s] bit i = (pat[i] has symbol s); one spare zero word for the banded variant's window.
// Only the rows of symbols the pattern contains are initialised (a short CDR3 touches ~10 of the 32 amino-acid rows, and
// clearing the table would cost more than the distance itself); the returned mask says which rows are valid.
template <int NSYM, bool AA>
EG_HD inline uint32_t build_peq(unsigned long long (*peq)[MAXW + 1], const uint8_t *pat, uint32_t m) {
    const uint32_t W = (m + 63) / 64;
    uint32_t present = 0;
    for (uint32_t i = 0; i < m; i++) {
        const uint32_t s = AA ? sym_aa(pat[i]) : sym_nt(pat[i]);
        if (!((present >> s) & 1u)) { for (uint32_t b = 0; b <= W; b++) peq[s][b] = 0ull; present |= 1u << s; }
        peq[s][i >> 6] |= 1ull << (i & 63);
    }
    return present;
}

// Levenshtein distance (unit costs, global) of pat[0..m) and txt[0..n), 0 < m <= 64 * MAXW: Myers' bit-vector algorithm in
// Hyyro's block form.  Block b holds rows 64b .. 64b+63 of the DP matrix as vertical deltas (Pv: +1, Mv: -1); a column
// is processed block by block, the horizontal delta leaving a block (hout) entering the next one (hin); the first block
// sees +1 (D[0][j] = j).  `score` follows the bottom row of the last block; the rows below the pattern's last one are
// then walked back with their vertical deltas.
template <int NSYM, bool AA>
EG_HD inline uint32_t myers_full(const unsigned long long (*peq)[MAXW + 1], uint32_t present, uint32_t m, const uint8_t *txt, uint32_t n) {
    const uint32_t W = (m + 63) / 64;
    unsigned long long Pv[MAXW], Mv[MAXW];
    for (uint32_t b = 0; b < W; b++) { Pv[b] = ~0ull; Mv[b] = 0ull; }
    int score = (int)(64 * W);
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t s = AA ? sym_aa(txt[j]) : sym_nt(txt[j]);
        const bool has = (present >> s) & 1u;
        int hin = 1;
        for (uint32_t b = 0; b < W; b++) {
            unsigned long long Eq = has ? peq[s][b] : 0ull;
            const unsigned long long pv = Pv[b], mv = Mv[b];
            const unsigned long long hneg = hin < 0 ? 1ull : 0ull;
            const unsigned long long Xv = Eq | mv;
            Eq |= hneg;
            const unsigned long long Xh = (((Eq & pv) + pv) ^ pv) | Eq;
            unsigned long long Ph = mv | ~(Xh | pv), Mh = pv & Xh;
            int hout = 0;
            if (Ph >> 63) hout = 1;
            if (Mh >> 63) hout = -1;
            Ph <<= 1; Mh <<= 1;
            Mh |= hneg;
            if (hin > 0) Ph |= 1ull;
            Pv[b] = Mh | ~(Xv | Ph);
            Mv[b] = Ph & Xv;
            hin = hout;
        }
        score += hin;
    }
    for (uint32_t r = 64 * W; r-- > m;) {          // rows m .. 64W-1: undo their vertical deltas
        const unsigned long long bit = 1ull << (r & 63);
        if (Pv[r >> 6] & bit) score--;
        if (Mv[r >> 6] & bit) score++;
    }
    return (uint32_t)score;
}

// Is the distance <= k?  For k <= 31 and |n - m| <= k only the diagonal band |row - column| <= k can hold a cell of an
// alignment that cheap (Ukkonen), and a band of 64 rows sliding down one row per column covers it (Hyyro's diagonal
// tiling): at column j bit b is row j + k - 63 + b.  The score is followed down the band's lowest diagonal (row j + k) until it reaches the last row,
// then along the last row; a score above 2k + n - m cannot come back under k.  Exact whenever the answer is yes.
template <int NSYM, bool AA>
EG_HD inline bool myers_band_within(const unsigned long long (*peq)[MAXW + 1], uint32_t present, uint32_t m, const uint8_t *txt, uint32_t n, uint32_t k) {
    // column 0 in the coordinates of column 1 (bit b = row k - 62 + b): rows 1 .. k+1 step +1 (bits 63-k .. 63); the rows
    // <= 0 are virtual, all equal to j in column j (vertical delta 0, no matches), which leaves D[0][j] = j
    unsigned long long VP = ~0ull << (63 - k), VN = 0ull;
    int curr = (int)(m >= k ? k : m);
    const int cut = (int)(2 * k + n - m);
    const uint32_t diag_cols = m > k ? m - k : 0;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t s = AA ? sym_aa(txt[i]) : sym_nt(txt[i]);
        const int start = (int)i + (int)k - 63;                    // pattern index of bit 0
        unsigned long long X;
        if (!((present >> s) & 1u)) X = 0ull;
        else if (start < 0) X = peq[s][0] << (-start);
        else {
            const uint32_t w = (uint32_t)start >> 6, o = (uint32_t)start & 63;
            X = peq[s][w] >> o;
            if (o) X |= peq[s][w + 1] << (64 - o);
        }
        const unsigned long long D0 = (((X & VP) + VP) ^ VP) | X | VN;
        const unsigned long long HP = VN | ~(D0 | VP), HN = D0 & VP;
        if (i < diag_cols) curr += (D0 >> 63) ? 0 : 1;
        else {
            const unsigned long long bit = 1ull << (63 - (i + 1 + k - m));
            if (HP & bit) curr++;
            if (HN & bit) curr--;
        }
        if (curr > cut) return false;
        VP = HN | ~((D0 >> 1) | HP);
        VN = (D0 >> 1) & HP;
    }
    return curr <= (int)k;
}

// distance(a, b) <= k; the shorter string is the pattern
template <int NSYM, bool AA>
EG_HD inline bool lev_within(const uint8_t *a, uint32_t n, const uint8_t *b, uint32_t m, uint32_t k) {
    if (n > m) { const uint8_t *t = a; a = b; b = t; const uint32_t u = n; n = m; m = u; }
    if (m - n > k) return false;
    if (n == 0) return true;                                       // distance m <= k
    unsigned long long peq[NSYM][MAXW + 1];
    const uint32_t present = build_peq<NSYM, AA>(peq, a, n);
    if (k <= 31) return myers_band_within<NSYM, AA>(peq, present, n, b, m, k);
    return myers_full<NSYM, AA>(peq, present, n, b, m) <= k;
}
template <int NSYM, bool AA>
EG_HD inline uint32_t myers(const uint8_t *pat, uint32_t m, const uint8_t *txt, uint32_t n) {
    if (m == 0) return n;
    if (n == 0) return m;
    unsigned long long peq[NSYM][MAXW + 1];
    const uint32_t present = build_peq<NSYM, AA>(peq, pat, m);
    return myers_full<NSYM, AA>(peq, present, m, txt, n);
}

// Amino-acid strings against a pattern of at most 64 symbols whose match vectors are already at hand (peq32[s], one word
// per symbol, valid where `present` says so): the classic one-word form, the score read at the pattern's last row.
EG_HD inline uint32_t myers64_aa(const unsigned long long *peq32, uint32_t present, uint32_t m, const uint8_t *txt, uint32_t n) {
    if (m == 0) return n;
    unsigned long long Pv = ~0ull, Mv = 0ull;
    const unsigned long long top = 1ull << (m - 1);
    uint32_t score = m;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t s = sym_aa(txt[j]);
        const unsigned long long Eq = ((present >> s) & 1u) ? peq32[s] : 0ull;
        const unsigned long long Xv = Eq | Mv;
        const unsigned long long Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        unsigned long long Ph = Mv | ~(Xh | Pv), Mh = Pv & Xh;
        if (Ph & top) score++;
        if (Mh & top) score--;
        Ph = (Ph << 1) | 1ull; Mh <<= 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
    }
    return score;
}

// largest d with (n1 - d) / n1 >= min_r || (n2 - d) / n2 >= min_r, the acceptance test of grouper.rs:427-431 in its own
// f64 expressions (monotone in d); -1 when even d = 0 fails
EG_HD inline bool ratio_ok(uint32_t d, uint32_t n1, uint32_t n2, double min_r) {
    const double r1 = (double)(d <= n1 ? n1 - d : 0u) / (double)n1, r2 = (double)(d <= n2 ? n2 - d : 0u) / (double)n2;
    return r1 >= min_r || r2 >= min_r;
}
EG_HD inline int dmax_ratio(uint32_t n1, uint32_t n2, double min_r) {
    const uint32_t top = n1 > n2 ? n1 : n2;
    if (!ratio_ok(0, n1, n2, min_r)) return -1;
    double e = (1.0 - min_r) * (double)top;
    uint32_t d = e <= 0.0 ? 0u : (e >= (double)top ? top : (uint32_t)e);
    while (d < top && ratio_ok(d + 1, n1, n2, min_r)) d++;
    while (d > 0 && !ratio_ok(d, n1, n2, min_r)) d--;
    return (int)d;
}

EG_HD inline uint32_t lev_nt(const uint8_t *a, uint32_t n, const uint8_t *b, uint32_t m) {      // the shorter string is the pattern
    return n <= m ? myers<7, false>(a, n, b, m) : myers<7, false>(b, m, a, n);
}
EG_HD inline uint32_t lev_aa(const uint8_t *a, uint32_t n, const uint8_t *b, uint32_t m) {
    return n <= m ? myers<32, true>(a, n, b, m) : myers<32, true>(b, m, a, n);
}

