// Fuzz of enclone_b200/csrc/egroup_myers.h (the kernel's Levenshtein arithmetic, compiled for the host) against the DP.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../enclone_b200/csrc/egroup_myers.h"
static uint32_t dp(const uint8_t *a, uint32_t n, const uint8_t *b, uint32_t m) {
    std::vector<uint32_t> row(m + 1);
    for (uint32_t j = 0; j <= m; j++) row[j] = j;
    for (uint32_t i = 1; i <= n; i++) {
        uint32_t prev = row[0];
        row[0] = i;
        for (uint32_t j = 1; j <= m; j++) {
            const uint32_t cur = row[j];
            row[j] = std::min({ prev + (a[i - 1] != b[j - 1]), row[j] + 1, row[j - 1] + 1 });
            prev = cur;
        }
    }
    return row[m];
}
int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    srand(1);
    const char *al = "ACGTN-", *aa = "ACDEFGHIKLMNPQRSTVWY*X";
    long bad = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t n = rand() % 513, m = rand() % 700;
        if (it % 7 == 0) n = 64 * (rand() % 9), m = n + rand() % 3;             // block boundaries
        if (it % 3 == 1) n = rand() % 65, m = rand() % 90;                       // CDR3-sized
        std::vector<uint8_t> a(n + 1), b(m + 1);
        const int mode = rand() % 3;
        for (auto &x : a) x = al[rand() % (mode ? 4 : 6)];
        if (mode == 2 && n > 0) {                                                // b = a with sparse edits
            b.clear();
            for (uint32_t i = 0; i < n; i++) { const int r = rand() % 30; if (r == 0) continue; if (r == 1) b.push_back('G'); b.push_back(r == 2 ? 'T' : a[i]); }
            m = (uint32_t)b.size();
            b.push_back(0);
        } else for (auto &x : b) x = al[rand() % 4];
        if (lev_nt(a.data(), n, b.data(), m) != dp(a.data(), n, b.data(), m)) bad++;
        for (auto &x : a) x = aa[rand() % 22];
        for (uint32_t i = 0; i < m; i++) b[i] = (rand() % 4 && i < n) ? a[i] : aa[rand() % 22];
        const uint32_t daa = dp(a.data(), n, b.data(), m);
        if (lev_aa(a.data(), n, b.data(), m) != daa) bad++;
        if (n <= 64) {                                                           // the one-word form with a prebuilt table (asymmetric grouping)
            unsigned long long pq[32], tab[32][MAXW + 1];
            const uint32_t present = build_peq<32, true>(tab, a.data(), n);
            for (int q = 0; q < 32; q++) pq[q] = tab[q][0];
            if (myers64_aa(pq, present, n, b.data(), m) != daa) { bad++; if (bad < 6) printf("myers64 n=%u m=%u\n", n, m); }
        }
        for (int q = 0; q < 4; q++) {                                            // d <= k, around the true distance and at random
            const uint32_t k = q < 2 ? (daa + (uint32_t)(rand() % 5) >= 2 ? daa + (uint32_t)(rand() % 5) - 2 : 0) : (uint32_t)(rand() % 48);
            if ((n <= 512 || m <= 512) && lev_within<32, true>(a.data(), n, b.data(), m, k) != (daa <= k)) { bad++; if (bad < 6) printf("aa within n=%u m=%u d=%u k=%u\n", n, m, daa, k); }
        }
    }
    for (int it = 0; it < iters; it++) {                                         // near pairs: the band variant's home ground
        const uint32_t n = 1 + rand() % 512;
        std::vector<uint8_t> a(n), b;
        for (auto &x : a) x = al[rand() % 4];
        const int rate = 8 + rand() % 60;
        for (uint32_t i = 0; i < n; i++) { const int r = rand() % rate; if (r == 0) continue; if (r == 1) b.push_back(al[rand() % 4]); b.push_back(r == 2 ? al[rand() % 4] : a[i]); }
        const uint32_t m = (uint32_t)b.size();
        b.push_back(0);
        const uint32_t d = dp(a.data(), n, b.data(), m);
        for (int q = 0; q < 6; q++) {
            const uint32_t k = q < 3 ? (d + (uint32_t)q >= 1 ? d + (uint32_t)q - 1 : 0) : (uint32_t)(rand() % 40);
            if (lev_within<7, false>(a.data(), n, b.data(), m, k) != (d <= k)) { bad++; if (bad < 6) printf("nt within n=%u m=%u d=%u k=%u\n", n, m, d, k); }
            if (lev_within<7, false>(b.data(), m, a.data(), n, k) != (d <= k)) { bad++; if (bad < 6) printf("nt within' n=%u m=%u d=%u k=%u\n", n, m, d, k); }
        }
        const double min_r = (50 + rand() % 5000 / 100.0) / 100.0;                // the ratio test and its threshold form agree for every d
        const int dm = dmax_ratio(n, m, min_r);
        for (uint32_t dd = 0; dd <= (n > m ? n : m); dd++)
            if (ratio_ok(dd, n, m, min_r) != ((int)dd <= dm)) { bad++; if (bad < 6) printf("dmax n=%u m=%u r=%f d=%u dm=%d\n", n, m, min_r, dd, dm); break; }
    }
    printf("bad=%ld\n", bad);
    return bad != 0;
}
