// Host-only check of the tier-1 work decomposition arithmetic (ejoin_dev.cuh): for every group size the
// (segment, row) enumeration is a bijection onto [0, items), the decode loop of item_decode() inverts it,
// every segment is at most `seg` tile steps long, and the extra slot rows are numbered 0 .. items - T - 1.
#include <cstdio>
#include <vector>

#include "../enclone_b200/csrc/ejoin_dev.cuh"

int main() {
    long checked = 0;
    for (uint32_t seg_min : { 8u, 4u, 2u, 1u })
    for (uint32_t T = 1; T <= 70000; T = T < 600 ? T + 1 : T + 997) {
        if (T > 65535) break;
        const uint32_t S = seg_of_tiles(T, seg_min), Q = (T + S - 1) / S;
        const uint32_t items = seg_item_off(T, S, Q);
        if (Q > SEG_QMAX) { printf("T=%u: %u segments\n", T, Q); return 1; }
        if (T <= seg_min && items != T) { printf("T=%u: small group must have one item per row\n", T); return 1; }
        std::vector<unsigned char> seen(items, 0);
        for (uint32_t q = 0; q < Q; q++)
            for (uint32_t ib = q * S; ib < T; ib++) {
                const uint32_t idx = seg_item_off(T, S, q) + (T - 1 - ib);
                if (idx >= items || seen[idx]) { printf("T=%u q=%u ib=%u: index %u out of range or repeated\n", T, q, ib, idx); return 1; }
                seen[idx] = 1;
                // decode as item_decode() does
                uint32_t rem = idx, qq = 0;
                while (rem >= T - qq * S) { rem -= T - qq * S; qq++; }
                if (qq != q || T - 1 - rem != ib) { printf("T=%u: decode(%u) = (%u,%u), expected (%u,%u)\n", T, idx, qq, T - 1 - rem, q, ib); return 1; }
                const uint32_t hi = (q + 1) * S - 1 < ib ? (q + 1) * S - 1 : ib;
                if (hi - q * S + 1 > S) { printf("T=%u: segment longer than %u\n", T, S); return 1; }
                if (q > 0 && idx - T >= items - T) { printf("T=%u: extra row out of range\n", T); return 1; }
                // phase A's view: row ib has ib / S + 1 segments
                if (q > ib / S) { printf("T=%u: row %u has no segment %u\n", T, ib, q); return 1; }
                checked++;
            }
        for (uint32_t i = 0; i < items; i++) if (!seen[i]) { printf("T=%u: item %u never produced\n", T, i); return 1; }
    }
    printf("ok %ld\n", checked);
    return 0;
}
