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

// Levenshtein distance (unit costs, global) of pat[0..m) and txt[0..n), m <= 64 * MAXW: Myers' bit-vector algorithm in
// Hyyro's block form.  Block b holds rows 64b .. 64b+63 of the DP matrix as vertical deltas (Pv: +1, Mv: -1); a column
// is processed block by block, the horizontal delta leaving a block (hout) entering the next one (hin); the first block
// sees +1 (D[0][j] = j).  `score` follows the bottom row of the last block; the rows below the pattern's last one are
// then walked back with their vertical deltas.
template <int NSYM, bool AA>
EG_HD inline uint32_t myers(const uint8_t *pat, uint32_t m, const uint8_t *txt, uint32_t n) {
    if (m == 0) return n;
    if (n == 0) return m;
    const uint32_t W = (m + 63) / 64;
    unsigned long long peq[NSYM][MAXW];
    for (int s = 0; s < NSYM; s++)
        for (uint32_t b = 0; b < W; b++) peq[s][b] = 0ull;
    for (uint32_t i = 0; i < m; i++) { const uint32_t s = AA ? sym_aa(pat[i]) : sym_nt(pat[i]); peq[s][i >> 6] |= 1ull << (i & 63); }
    unsigned long long Pv[MAXW], Mv[MAXW];
    for (uint32_t b = 0; b < W; b++) { Pv[b] = ~0ull; Mv[b] = 0ull; }
    int score = (int)(64 * W);
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t s = AA ? sym_aa(txt[j]) : sym_nt(txt[j]);
        int hin = 1;
        for (uint32_t b = 0; b < W; b++) {
            unsigned long long Eq = peq[s][b];
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

EG_HD inline uint32_t lev_nt(const uint8_t *a, uint32_t n, const uint8_t *b, uint32_t m) {      // the shorter string is the pattern
    return n <= m ? myers<7, false>(a, n, b, m) : myers<7, false>(b, m, a, n);
}
EG_HD inline uint32_t lev_aa(const uint8_t *a, uint32_t n, const uint8_t *b, uint32_t m) {
    return n <= m ? myers<32, true>(a, n, b, m) : myers<32, true>(b, m, a, n);
}

