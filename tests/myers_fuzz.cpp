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
        if (lev_aa(a.data(), n, b.data(), m) != dp(a.data(), n, b.data(), m)) bad++;
    }
    printf("bad=%ld\n", bad);
    return bad != 0;
}
