/* refine_oracle.c — CPU restatement of enclone's post-join refinement.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED.  The code this restates (define_mat, delete_doublets, the signature filter, weak_chains, split_orbits:
 * enclone_ranger@cbd15ef, Cargo.lock:632-634) is not in /root/reference and there is no Rust toolchain, so every clause
 * below is either the documentation's wording [DOC] or the builder's recollection of the upstream source [RECALL], marked
 * as such.  What the reference tree does hold: the prose of pages/heuristics.html.src:19-41 (chain grouping) and
 * pages/default_filters.html.src:292-353 (filters), and the before / after goldens of the doublet filter
 * (enclone_exec/testx/inputs/outputs/enclone_test170_output, enclone_test171_output; commands
 * enclone_testlist/src/main_testlist.rs:399-402), which tests/test_refine.py reproduces on a reconstruction of that
 * clonotype plus the partner clonotype the goldens do not print (so: consistent with, not pinned by).
 *
 * One "stage" = define_mat over every clonotype + the break-up + the pure subclonotypes:
 *  S1 [RECALL define_mat] the exact subclonotypes of a clonotype are listed by (origin, exact id); its chains are all
 *     (u, m) = (exact subclonotype, share index); an EquivRel over the chains says which share a column.
 *  S2 [DOC heuristics:25-30] every raw join (k1, k2) joins chain exact_cols[c] of k1's exact subclonotype with chain
 *     exact_cols[c] of k2's, c = heavy, light.
 *  S3 [DOC heuristics:36-38] "two exact subclonotypes are joined, and both have three chains": the third chains (the ones
 *     the join did not align) share a column if they are of the same kind, of equal length and differ in at most 10 bases.
 *  S4 [DOC heuristics:39-40] a onesie's chain is connected to every chain of its clonotype with the identical V..J.
 *     (heuristics:31-35, the re-test of join_one on pairs the join skipped, is the caller's: it appends what it recovers to
 *     raw_joins before the call.)
 *  S5 [RECALL define_mat] columns: the chain orbits ordered by the position of their representative — the first member in
 *     (u, m) order — in the list of all chains sorted by ("chain_type:cdr3_aa", seq.len(), u, m) (heavy before light); an
 *     orbit that holds up to mm chains of one exact subclonotype becomes mm columns, the j-th of them taking every exact
 *     subclonotype's j-th chain of the orbit.
 *  S6 [RECALL split_orbits; DOC default_filters:315-317] a clonotype breaks up into the connected components of "shares a
 *     column"; a clonotype is named by its smallest exact id.
 *  S7 [DOC default_filters:296-303; RECALL] pure subclonotype = the exact subclonotypes of one clonotype with entries in
 *     exactly the same columns.
 * Filters, each on the stage computed after the previous filter's deletions:
 *  F1 doublet [DOC :304-314] p0 is deleted if there are pure subclonotypes p1, p2 (anywhere in the data) with: p0 and p1 share
 *     an identical CDR3 DNA sequence, p0 and p2 do, p1 and p2 do not, and 5 ncells(p0) <= min(ncells(p1), ncells(p2)).
 *  F2 signature [DOC :330-339] p with at least two chains is deleted if the cells of the two-chain pure subclonotypes of
 *     its clonotype, other than p, that share a chain (column) with it number at least 20 ncells(p).
 *  F3 weak chains [DOC :345-353] in a clonotype of three or more chains, a chain supported by at most 20 cells, with 8 times
 *     that number less than the cells of the clonotype, takes every exact subclonotype that has it out.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/enclone_refine.h"

typedef struct {
    const erefine_pack_t *pk;
    const erefine_opts_t *op;
    uint32_t nx, nc;
    uint32_t *xof;            /* chain -> exact */
    uint8_t *alive;           /* exact */
    uint32_t *label;          /* exact -> clonotype (smallest exact id), EREFINE_NONE */
    uint32_t *pos, *pos2;     /* chain -> position in the two global orders */
    uint32_t *cp;             /* chain union-find */
    uint32_t *xp;             /* exact union-find */
    uint32_t *col;            /* chain -> column */
    uint32_t *ncols;          /* exact -> columns of its clonotype */
    uint32_t *pure;           /* exact -> pure subclonotype id */
    uint32_t n_pure;
} ctx_t;

static uint32_t uf_find(uint32_t *p, uint32_t a) { uint32_t r = a; while (p[r] != r) r = p[r]; while (p[a] != r) { uint32_t nx = p[a]; p[a] = r; a = nx; } return r; }
static void uf_join(uint32_t *p, uint32_t a, uint32_t b) { a = uf_find(p, a); b = uf_find(p, b); if (a == b) return; if (a < b) p[b] = a; else p[a] = b; }

typedef struct { uint64_t a, b; uint32_t i; } key3_t;
static int cmp_key3(const void *x, const void *y) {
    const key3_t *p = (const key3_t *)x, *q = (const key3_t *)y;
    if (p->a != q->a) return p->a < q->a ? -1 : 1;
    if (p->b != q->b) return p->b < q->b ? -1 : 1;
    return p->i < q->i ? -1 : p->i > q->i;
}

/* columns an exact subclonotype has entries in, sorted, as a fixed array (EREFINE_NONE padded) */
typedef struct { uint32_t label, c[EREFINE_MAX_CHAINS], x; } sig_t;
static int cmp_sig(const void *x, const void *y) {
    const sig_t *p = (const sig_t *)x, *q = (const sig_t *)y;
    if (p->label != q->label) return p->label < q->label ? -1 : 1;
    for (int i = 0; i < EREFINE_MAX_CHAINS; i++) if (p->c[i] != q->c[i]) return p->c[i] < q->c[i] ? -1 : 1;
    return p->x < q->x ? -1 : p->x > q->x;
}
static int same_sig(const sig_t *p, const sig_t *q) { return p->label == q->label && memcmp(p->c, q->c, sizeof(p->c)) == 0; }
static int cmp_u32(const void *x, const void *y) { uint32_t a = *(const uint32_t *)x, b = *(const uint32_t *)y; return a < b ? -1 : a > b; }

static void stage(ctx_t *C) {
    const erefine_pack_t *pk = C->pk;
    const uint32_t nx = C->nx, nc = C->nc;
    for (uint32_t c = 0; c < nc; c++) C->cp[c] = c;
    /* S2 + S3 */
    for (uint64_t j = 0; j < pk->n_joins; j++) {
        const uint32_t k1 = pk->join_k1[j], k2 = pk->join_k2[j];
        const uint32_t x1 = pk->exact_id[k1], x2 = pk->exact_id[k2];
        if (!C->alive[x1] || !C->alive[x2]) continue;
        for (int c = 0; c < 2; c++) {
            const uint8_t m1 = pk->exact_cols[c][k1], m2 = pk->exact_cols[c][k2];
            if (m1 == 0xFF || m2 == 0xFF) continue;
            uf_join(C->cp, pk->chain_off[x1] + m1, pk->chain_off[x2] + m2);
        }
        const uint32_t n1 = pk->chain_off[x1 + 1] - pk->chain_off[x1], n2 = pk->chain_off[x2 + 1] - pk->chain_off[x2];
        if (n1 == 3 && n2 == 3) {
            const uint8_t a0 = pk->exact_cols[0][k1], a1 = pk->exact_cols[1][k1], b0 = pk->exact_cols[0][k2], b1 = pk->exact_cols[1][k2];
            if (a0 < 3 && a1 < 3 && a0 != a1 && b0 < 3 && b1 < 3 && b0 != b1) {
                const uint32_t t1 = pk->chain_off[x1] + (3 - a0 - a1), t2 = pk->chain_off[x2] + (3 - b0 - b1);
                if (pk->left[t1] == pk->left[t2] && pk->seq_len[t1] == pk->seq_len[t2]) {
                    const uint8_t *s1 = pk->seq_arena + pk->seq_off[t1], *s2 = pk->seq_arena + pk->seq_off[t2];
                    uint32_t d = 0;
                    for (uint32_t i = 0; i < pk->seq_len[t1]; i++) d += s1[i] != s2[i];
                    if ((int64_t)d <= C->op->third_chain_diffs) uf_join(C->cp, t1, t2);
                }
            }
        }
    }
    /* S4 */
    {
        key3_t *v = (key3_t *)malloc(sizeof(key3_t) * (nc ? nc : 1));
        uint32_t n = 0;
        for (uint32_t c = 0; c < nc; c++) if (C->alive[C->xof[c]]) { v[n].a = C->label[C->xof[c]]; v[n].b = pk->seq_id[c]; v[n].i = c; n++; }
        qsort(v, n, sizeof(key3_t), cmp_key3);
        for (uint32_t i = 0; i < n;) {
            uint32_t j = i;
            int onesie = 0;
            while (j < n && v[j].a == v[i].a && v[j].b == v[i].b) {
                const uint32_t x = C->xof[v[j].i];
                if (pk->chain_off[x + 1] - pk->chain_off[x] == 1) onesie = 1;
                j++;
            }
            if (onesie) for (uint32_t l = i + 1; l < j; l++) uf_join(C->cp, v[i].i, v[l].i);
            i = j;
        }
        free(v);
    }
    /* S6: components of "shares a column" */
    for (uint32_t x = 0; x < nx; x++) C->xp[x] = x;
    for (uint32_t c = 0; c < nc; c++) if (C->alive[C->xof[c]]) uf_join(C->xp, C->xof[c], C->xof[uf_find(C->cp, c)]);
    for (uint32_t x = 0; x < nx; x++) C->label[x] = C->alive[x] ? uf_find(C->xp, x) : EREFINE_NONE;   /* union by smaller root: the smallest exact id */
    /* S5: representative of every chain orbit, multiplicities, column bases */
    uint64_t *rep = (uint64_t *)malloc(8 * (size_t)(nc ? nc : 1));
    uint32_t *mm = (uint32_t *)calloc(nc ? nc : 1, 4), *base = (uint32_t *)calloc(nc ? nc : 1, 4);
    for (uint32_t c = 0; c < nc; c++) rep[c] = ~0ull;
    for (uint32_t c = 0; c < nc; c++) if (C->alive[C->xof[c]]) {
        const uint32_t r = uf_find(C->cp, c);
        const uint64_t key = ((uint64_t)C->pos2[c] << 32) | C->pos[c];
        if (key < rep[r]) rep[r] = key;
    }
    for (uint32_t x = 0; x < nx; x++) if (C->alive[x]) {
        const uint32_t o = pk->chain_off[x], m = pk->chain_off[x + 1] - o;
        for (uint32_t a = 0; a < m; a++) {
            const uint32_t r = uf_find(C->cp, o + a);
            uint32_t cnt = 0;
            for (uint32_t b = 0; b < m; b++) cnt += uf_find(C->cp, o + b) == r;
            if (cnt > mm[r]) mm[r] = cnt;
        }
    }
    uint32_t *ncol_root = (uint32_t *)calloc(nc ? nc : 1, 4);
    {
        key3_t *v = (key3_t *)malloc(sizeof(key3_t) * (nc ? nc : 1));
        uint32_t n = 0;
        for (uint32_t c = 0; c < nc; c++) if (C->alive[C->xof[c]] && uf_find(C->cp, c) == c) { v[n].a = C->label[C->xof[c]]; v[n].b = rep[c] & 0xFFFFFFFFull; v[n].i = c; n++; }
        qsort(v, n, sizeof(key3_t), cmp_key3);
        for (uint32_t i = 0; i < n;) {
            uint32_t j = i, run = 0;
            while (j < n && v[j].a == v[i].a) { base[v[j].i] = run; run += mm[v[j].i]; j++; }
            for (uint32_t l = i; l < j; l++) ncol_root[v[l].i] = run;     /* mat.len() of the clonotype, parked on its orbit roots */
            i = j;
        }
        free(v);
    }
    sig_t *sg = (sig_t *)malloc(sizeof(sig_t) * (nx ? nx : 1));
    uint32_t ns = 0;
    for (uint32_t x = 0; x < nx; x++) {
        const uint32_t o = pk->chain_off[x], m = pk->chain_off[x + 1] - o;
        if (!C->alive[x]) { for (uint32_t a = 0; a < m; a++) C->col[o + a] = EREFINE_NONE; continue; }
        sig_t s;
        s.label = C->label[x]; s.x = x;
        for (int i = 0; i < EREFINE_MAX_CHAINS; i++) s.c[i] = EREFINE_NONE;
        for (uint32_t a = 0; a < m; a++) {
            const uint32_t r = uf_find(C->cp, o + a);
            uint32_t before = 0;
            for (uint32_t b = 0; b < a; b++) before += uf_find(C->cp, o + b) == r;
            C->col[o + a] = base[r] + before;
            s.c[a] = C->col[o + a];
        }
        qsort(s.c, EREFINE_MAX_CHAINS, 4, cmp_u32);
        sg[ns++] = s;
    }
    /* n_columns per exact: the total of its clonotype = ncol_root of any of its chain roots */
    for (uint32_t x = 0; x < nx; x++) {
        const uint32_t o = pk->chain_off[x], m = pk->chain_off[x + 1] - o;
        C->ncols[x] = (C->alive[x] && m) ? ncol_root[uf_find(C->cp, o)] : 0;
    }
    /* S7 */
    qsort(sg, ns, sizeof(sig_t), cmp_sig);
    uint32_t np = 0;
    for (uint32_t x = 0; x < nx; x++) C->pure[x] = EREFINE_NONE;
    for (uint32_t i = 0; i < ns; i++) {
        if (i > 0 && !same_sig(&sg[i], &sg[i - 1])) np++;
        C->pure[sg[i].x] = np;
    }
    C->n_pure = ns ? np + 1 : 0;
    free(sg); free(ncol_root); free(rep); free(mm); free(base);
}

/* F1: the doublet filter.  share(p, q) <=> the two pure subclonotypes hold an identical CDR3 DNA sequence. */
typedef struct { uint32_t cdr3, pure; } cp_t;
static int cmp_cp(const void *x, const void *y) {
    const cp_t *p = (const cp_t *)x, *q = (const cp_t *)y;
    if (p->cdr3 != q->cdr3) return p->cdr3 < q->cdr3 ? -1 : 1;
    return p->pure < q->pure ? -1 : p->pure > q->pure;
}
static int cmp_pc(const void *x, const void *y) {
    const cp_t *p = (const cp_t *)x, *q = (const cp_t *)y;
    if (p->pure != q->pure) return p->pure < q->pure ? -1 : 1;
    return p->cdr3 < q->cdr3 ? -1 : p->cdr3 > q->cdr3;
}
static uint64_t doublet(ctx_t *C, uint8_t *del, uint64_t *triples) {
    const erefine_pack_t *pk = C->pk;
    const uint32_t np = C->n_pure;
    uint64_t *npure = (uint64_t *)calloc(np ? np : 1, 8);
    cp_t *by_c = (cp_t *)malloc(sizeof(cp_t) * (C->nc ? C->nc : 1));
    uint32_t n = 0;
    for (uint32_t x = 0; x < C->nx; x++) if (C->alive[x]) {
        npure[C->pure[x]] += pk->ncells[x];
        for (uint32_t c = pk->chain_off[x]; c < pk->chain_off[x + 1]; c++) { by_c[n].cdr3 = pk->cdr3_id[c]; by_c[n].pure = C->pure[x]; n++; }
    }
    qsort(by_c, n, sizeof(cp_t), cmp_cp);
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; i++) if (i == 0 || cmp_cp(&by_c[i], &by_c[i - 1]) != 0) by_c[m++] = by_c[i];     /* content.dedup() */
    cp_t *by_p = (cp_t *)malloc(sizeof(cp_t) * (m ? m : 1));
    memcpy(by_p, by_c, sizeof(cp_t) * m);
    qsort(by_p, m, sizeof(cp_t), cmp_pc);
    uint32_t *p_off = (uint32_t *)calloc((size_t)np + 2, 4);
    for (uint32_t i = 0; i < m; i++) p_off[by_p[i].pure + 1]++;
    for (uint32_t p = 0; p < np; p++) p_off[p + 1] += p_off[p];
    uint32_t *part = (uint32_t *)malloc(4 * (size_t)(m ? m : 1)), *acdr = (uint32_t *)malloc(4 * (size_t)(m ? m : 1));
    uint64_t ndel = 0;
    uint8_t *pdel = (uint8_t *)calloc(np ? np : 1, 1);
    for (uint32_t p0 = 0; p0 < np; p0++) {
        /* partners of p0 with enough cells, each with the CDR3 of p0 it was found through (a partner may come up several times) */
        uint32_t npart = 0;
        const uint64_t need = (uint64_t)C->op->doublet_mult * npure[p0];
        for (uint32_t i = p_off[p0]; i < p_off[p0 + 1]; i++) {
            const uint32_t cd = by_p[i].cdr3;
            uint32_t lo = 0, hi = m;
            while (lo < hi) { const uint32_t mid = (lo + hi) / 2; if (by_c[mid].cdr3 < cd) lo = mid + 1; else hi = mid; }
            for (uint32_t j = lo; j < m && by_c[j].cdr3 == cd; j++) {
                const uint32_t q = by_c[j].pure;
                if (q == p0 || npure[q] < need) continue;
                if (npart == m) break;
                part[npart] = q; acdr[npart] = cd; npart++;
            }
        }
        int hit = 0;
        for (uint32_t a = 0; a < npart && !hit; a++) for (uint32_t b = a + 1; b < npart && !hit; b++) {
            const uint32_t p1 = part[a], p2 = part[b];
            if (p1 == p2) continue;
            (*triples)++;
            /* share(p1, p2): merge their sorted CDR3 lists */
            uint32_t i = p_off[p1], j = p_off[p2];
            int sh = 0;
            while (i < p_off[p1 + 1] && j < p_off[p2 + 1]) {
                if (by_p[i].cdr3 == by_p[j].cdr3) { sh = 1; break; }
                if (by_p[i].cdr3 < by_p[j].cdr3) i++; else j++;
            }
            if (!sh) hit = 1;
        }
        if (hit) pdel[p0] = 1;
    }
    for (uint32_t x = 0; x < C->nx; x++) if (C->alive[x] && pdel[C->pure[x]]) { del[x] = 1; ndel++; }
    free(npure); free(by_c); free(by_p); free(p_off); free(part); free(acdr); free(pdel);
    return ndel;
}

/* F2: the signature filter */
static uint64_t signature(ctx_t *C, uint8_t *del) {
    const erefine_pack_t *pk = C->pk;
    const uint32_t np = C->n_pure, nx = C->nx;
    uint64_t *npure = (uint64_t *)calloc(np ? np : 1, 8);
    uint32_t *first = (uint32_t *)malloc(4 * (size_t)(np ? np : 1));     /* one exact subclonotype of the pure subclonotype: its columns are the signature */
    for (uint32_t p = 0; p < np; p++) first[p] = EREFINE_NONE;
    for (uint32_t x = 0; x < nx; x++) if (C->alive[x]) { npure[C->pure[x]] += pk->ncells[x]; if (first[C->pure[x]] == EREFINE_NONE) first[C->pure[x]] = x; }
    /* pure ids are consecutive inside a clonotype (S7 sorts by clonotype first) */
    uint64_t ndel = 0;
    uint8_t *pdel = (uint8_t *)calloc(np ? np : 1, 1);
    for (uint32_t p = 0; p < np;) {
        uint32_t q = p;
        while (q < np && C->label[first[q]] == C->label[first[p]]) q++;
        for (uint32_t a = p; a < q; a++) {
            const uint32_t xa = first[a], oa = pk->chain_off[xa], ma = pk->chain_off[xa + 1] - oa;
            if (ma < 2) continue;
            uint64_t tot = 0;
            for (uint32_t b = p; b < q; b++) {
                if (b == a) continue;
                const uint32_t xb = first[b], ob = pk->chain_off[xb], mb = pk->chain_off[xb + 1] - ob;
                if (mb != 2) continue;
                int sh = 0;
                for (uint32_t i = 0; i < ma; i++) for (uint32_t j = 0; j < mb; j++) sh |= C->col[oa + i] == C->col[ob + j];
                if (sh) tot += npure[b];
            }
            if (tot >= (uint64_t)C->op->signature_mult * npure[a]) pdel[a] = 1;
        }
        p = q;
    }
    for (uint32_t x = 0; x < nx; x++) if (C->alive[x] && pdel[C->pure[x]]) { del[x] = 1; ndel++; }
    free(npure); free(first); free(pdel);
    return ndel;
}

/* F3: the weak-chain filter */
static uint64_t weak_chains(ctx_t *C, uint8_t *del) {
    const erefine_pack_t *pk = C->pk;
    const uint32_t nx = C->nx;
    /* cells per (clonotype, column) and per clonotype: keyed by clonotype label */
    uint64_t *tot = (uint64_t *)calloc(nx ? nx : 1, 8);
    for (uint32_t x = 0; x < nx; x++) if (C->alive[x]) tot[C->label[x]] += pk->ncells[x];
    key3_t *v = (key3_t *)malloc(sizeof(key3_t) * (C->nc ? C->nc : 1));
    uint32_t n = 0;
    for (uint32_t x = 0; x < nx; x++) if (C->alive[x] && C->ncols[x] >= 3)
        for (uint32_t c = pk->chain_off[x]; c < pk->chain_off[x + 1]; c++) { v[n].a = C->label[x]; v[n].b = C->col[c]; v[n].i = x; n++; }
    qsort(v, n, sizeof(key3_t), cmp_key3);
    uint64_t ndel = 0;
    for (uint32_t i = 0; i < n;) {
        uint32_t j = i;
        uint64_t cells = 0;
        while (j < n && v[j].a == v[i].a && v[j].b == v[i].b) { cells += pk->ncells[v[j].i]; j++; }
        if (cells <= (uint64_t)C->op->weak_max_cells && (uint64_t)C->op->weak_mult * cells < tot[v[i].a])
            for (uint32_t l = i; l < j; l++) if (!del[v[l].i]) { del[v[l].i] = 1; ndel++; }
        i = j;
    }
    free(tot); free(v);
    return ndel;
}

void oracle_refine_opts_default(erefine_opts_t *o) {
    memset(o, 0, sizeof(*o));
    o->doublet = o->signature = o->weak_chains = 1;
    o->doublet_mult = 5; o->signature_mult = 20; o->weak_max_cells = 20; o->weak_mult = 8; o->third_chain_diffs = 10;
}

int oracle_refine(const erefine_pack_t *pk, const erefine_opts_t *op, erefine_result_t *out, erefine_stats_t *st) {
    if (!pk || !op || !out || pk->abi_version != EREFINE_ABI_VERSION) return -1;
    const uint32_t nx = pk->n_exact, nc = pk->n_chain;
    for (uint32_t x = 0; x < nx; x++) if (pk->chain_off[x + 1] - pk->chain_off[x] > EREFINE_MAX_CHAINS) return -4;
    ctx_t C;
    memset(&C, 0, sizeof(C));
    C.pk = pk; C.op = op; C.nx = nx; C.nc = nc;
    C.xof = (uint32_t *)malloc(4 * (size_t)(nc ? nc : 1));
    C.alive = (uint8_t *)calloc(nx ? nx : 1, 1);
    C.label = (uint32_t *)malloc(4 * (size_t)(nx ? nx : 1));
    C.pos = (uint32_t *)malloc(4 * (size_t)(nc ? nc : 1)); C.pos2 = (uint32_t *)malloc(4 * (size_t)(nc ? nc : 1));
    C.cp = (uint32_t *)malloc(4 * (size_t)(nc ? nc : 1)); C.xp = (uint32_t *)malloc(4 * (size_t)(nx ? nx : 1));
    C.col = (uint32_t *)malloc(4 * (size_t)(nc ? nc : 1)); C.ncols = (uint32_t *)calloc((size_t)nx + 1, 4);
    C.pure = (uint32_t *)malloc(4 * (size_t)(nx ? nx : 1));
    for (uint32_t x = 0; x < nx; x++) for (uint32_t c = pk->chain_off[x]; c < pk->chain_off[x + 1]; c++) C.xof[c] = x;
    /* the two global chain orders of S5 */
    {
        key3_t *v = (key3_t *)malloc(sizeof(key3_t) * (nc ? nc : 1));
        for (uint32_t c = 0; c < nc; c++) { v[c].a = pk->origin_rank[C.xof[c]]; v[c].b = 0; v[c].i = c; }
        qsort(v, nc, sizeof(key3_t), cmp_key3);
        for (uint32_t i = 0; i < nc; i++) C.pos2[v[i].i] = i;
        for (uint32_t c = 0; c < nc; c++) { v[c].a = pk->order_rank[c]; v[c].b = C.pos2[c]; v[c].i = c; }
        qsort(v, nc, sizeof(key3_t), cmp_key3);
        for (uint32_t i = 0; i < nc; i++) C.pos[v[i].i] = i;
        free(v);
    }
    /* the clonotypes the join left: classes of class_of, named by their smallest exact id */
    {
        uint32_t *cmin = (uint32_t *)malloc(4 * (size_t)(pk->n_info ? pk->n_info : 1));
        for (uint32_t k = 0; k < pk->n_info; k++) cmin[k] = EREFINE_NONE;
        for (uint32_t k = 0; k < pk->n_info; k++) { const uint32_t c = pk->class_of[k]; if (pk->exact_id[k] < cmin[c]) cmin[c] = pk->exact_id[k]; }
        for (uint32_t x = 0; x < nx; x++) C.label[x] = EREFINE_NONE;
        for (uint32_t k = 0; k < pk->n_info; k++) { const uint32_t x = pk->exact_id[k]; C.alive[x] = 1; if (cmin[pk->class_of[k]] < C.label[x]) C.label[x] = cmin[pk->class_of[k]]; }
        free(cmin);
    }
    erefine_stats_t S;
    memset(&S, 0, sizeof(S));
    for (uint32_t x = 0; x < nx; x++) S.n_clonotypes_in += C.alive[x] && C.label[x] == x;
    memset(out->fate, 0, nx);
    uint8_t *del = (uint8_t *)calloc(nx ? nx : 1, 1);
    stage(&C);
    S.n_pure = C.n_pure;
    for (int f = 1; f <= 3; f++) {
        const int on = f == 1 ? op->doublet : f == 2 ? op->signature : op->weak_chains;
        if (!on) continue;
        memset(del, 0, nx);
        const uint64_t nd = f == 1 ? doublet(&C, del, &S.triples_tested) : f == 2 ? signature(&C, del) : weak_chains(&C, del);
        if (f == 1) S.n_doublet = nd; else if (f == 2) S.n_signature = nd; else S.n_weak = nd;
        if (nd) {
            for (uint32_t x = 0; x < nx; x++) if (del[x]) { C.alive[x] = 0; out->fate[x] = (uint8_t)f; }
            stage(&C);
        }
    }
    for (uint32_t x = 0; x < nx; x++) {
        out->clonotype_of[x] = C.alive[x] ? C.label[x] : EREFINE_NONE;
        out->n_columns[x] = C.alive[x] ? C.ncols[x] : 0;
        S.n_clonotypes_out += C.alive[x] && C.label[x] == x;
    }
    for (uint32_t c = 0; c < nc; c++) out->column_of[c] = C.col[c];
    if (st) *st = S;
    free(del); free(C.xof); free(C.alive); free(C.label); free(C.pos); free(C.pos2); free(C.cp); free(C.xp); free(C.col); free(C.ncols); free(C.pure);
    return 0;
}
