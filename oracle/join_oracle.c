/* join_oracle.c — CPU restatement of enclone's clonotype join.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle and the timed CPU baseline.  It is never linked into,
 * imported by, or called from the product path (enclone_b200/): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * WHAT IT RESTATES.  The reference's join lives in the un-vendored git dependency
 * 10XGenomics/enclone_ranger @ cbd15ef94236a91ada7227322303593166b63cde
 * (/root/reference/Cargo.lock:632-634: crates enclone, enclone_core, equiv) with the
 * arithmetic of debruijn 0.3.4 (Cargo.lock:527-530) and qd 0.2.0-alpha@0fb276d
 * (Cargo.lock:2133-2135).  None of that source is in /root/reference and there is no Rust
 * toolchain, so the reference cannot be compiled here (no oracle/_ref).  The algorithm is
 * restated from the in-tree specification:
 *   - enclone_help/src/help1.rs:270-373           the join algorithm, steps 4-11
 *   - pages/cellranger.html.src:130-247           current defaults of every knob
 *   - pages/heuristics.html.src:25-41             only-as-many-joins-as-needed
 *   - enclone_tail/src/print_clonotypes/print_utils5.rs:74-87   V/J window indexing vs germline
 *   - enclone_tail/src/grouper.rs:404-452         the EquivRel skip-if-same-class loop idiom
 *   - enclone_exec/testx/inputs/outputs/enclone_test256_output:29-70   known-answer vector
 * and SURVEY.md Appendix A (which tags what is documented and what is recollection).
 *
 * PINNING.  The one golden vector the reference holds for this path (regression test 256:
 * shares 5, indeps 23, cd 2, k 33, d 5, n 2373, p1 1.6600942157683246e-6,
 * mult 27576.76403240813, score 0.04578002645980871, FWR1/CDR1/CDR2/CDR3 diffs 2/0/3/0,
 * identities 97.3/92.3) is checked bit-exactly by tests/test_oracle_kat.py on the real
 * flaky8a/flaky8b contigs.  Beyond that one pair: PARITY WITH REAL ENCLONE IS UNPINNED
 * (DESIGN.md says the same); every clause that rests on recollection is marked [RECALL].
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/enclone_join.h"

/* ------------------------------------------------------------------------------------------
 * double-double arithmetic: stands in for qd::Double (Cargo.lock:2133-2135), which the
 * reference uses for the probability after the Jan-2022 precision fix
 * (pages/cellranger.html.src:160-162).  Compile with -ffp-contract=off.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double hi, lo; } dd_t;

static inline dd_t dd_two_sum(double a, double b) {
    double s = a + b, bb = s - a;
    dd_t r = { s, (a - (s - bb)) + (b - bb) };
    return r;
}
static inline dd_t dd_quick_two_sum(double a, double b) {
    double s = a + b;
    dd_t r = { s, b - (s - a) };
    return r;
}
static inline dd_t dd_two_prod(double a, double b) {
    double p = a * b;
    dd_t r = { p, fma(a, b, -p) };
    return r;
}
static inline dd_t dd_from(double a) { dd_t r = { a, 0.0 }; return r; }
static inline dd_t dd_add(dd_t a, dd_t b) {
    dd_t s = dd_two_sum(a.hi, b.hi), t = dd_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = dd_quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return dd_quick_two_sum(s.hi, s.lo);
}
static inline dd_t dd_neg(dd_t a) { dd_t r = { -a.hi, -a.lo }; return r; }
static inline dd_t dd_sub(dd_t a, dd_t b) { return dd_add(a, dd_neg(b)); }
static inline dd_t dd_mul(dd_t a, dd_t b) {
    dd_t p = dd_two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return dd_quick_two_sum(p.hi, p.lo);
}
static inline dd_t dd_mul_d(dd_t a, double b) {
    dd_t p = dd_two_prod(a.hi, b);
    p.lo += a.lo * b;
    return dd_quick_two_sum(p.hi, p.lo);
}
static inline dd_t dd_div(dd_t a, dd_t b) {
    double q1 = a.hi / b.hi;
    dd_t r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    dd_t q = dd_quick_two_sum(q1, q2);
    return dd_add(q, dd_from(q3));
}

/* ------------------------------------------------------------------------------------------
 * Stirling ratio table  sr[x][u] = S(x,u) * u! / u^x   (the `sr: &[Vec<Double>]` argument
 * of join_exacts, field EncloneExacts.sr, enclone_tail/src/print_clonotypes/mod.rs:327;
 * method pointer enclone_help/src/help1.rs:292-298 -> docs.rs/stirling_numbers).
 * Recurrence: sr[x][u] = sr[x-1][u] + sr[x-1][u-1] * ((u-1)/u)^(x-1).
 * Built once (outside the timed join, as in the reference where it is an input).
 * ---------------------------------------------------------------------------------------- */
#define SR_CAP 4096
static dd_t *g_sr_rows[SR_CAP + 1];
static _Atomic int g_sr_built = -1; /* highest x built */
static pthread_mutex_t g_sr_mu = PTHREAD_MUTEX_INITIALIZER;
static dd_t *g_sr_pw;    /* ((u-1)/u)^(x) running powers, x = g_sr_built */
static dd_t *g_sr_ratio; /* (u-1)/u */

static int sr_ensure(int xmax) {
    if (xmax > SR_CAP) return -1;
    if (atomic_load(&g_sr_built) >= xmax) return 0;
    pthread_mutex_lock(&g_sr_mu);
    int built = atomic_load(&g_sr_built);
    if (built < 0) {
        g_sr_pw = (dd_t *)malloc(sizeof(dd_t) * (SR_CAP + 2));
        g_sr_ratio = (dd_t *)malloc(sizeof(dd_t) * (SR_CAP + 2));
        for (int u = 1; u <= SR_CAP + 1; u++) {
            g_sr_ratio[u] = dd_div(dd_from((double)(u - 1)), dd_from((double)u));
            g_sr_pw[u] = dd_from(1.0); /* exponent 0 */
        }
        g_sr_rows[0] = (dd_t *)malloc(sizeof(dd_t));
        g_sr_rows[0][0] = dd_from(1.0);
        built = 0;
        atomic_store(&g_sr_built, 0);
    }
    for (int x = built + 1; x <= xmax; x++) {
        dd_t *row = (dd_t *)malloc(sizeof(dd_t) * (size_t)(x + 1));
        const dd_t *prev = g_sr_rows[x - 1];
        row[0] = dd_from(0.0);
        for (int u = 1; u <= x; u++) {
            dd_t a = (u <= x - 1) ? prev[u] : dd_from(0.0);
            /* g_sr_pw[u] holds ((u-1)/u)^(x-1) here */
            row[u] = dd_add(a, dd_mul(prev[u - 1], g_sr_pw[u]));
        }
        for (int u = 1; u <= SR_CAP + 1; u++) g_sr_pw[u] = dd_mul(g_sr_pw[u], g_sr_ratio[u]);
        g_sr_rows[x] = row;
        atomic_store(&g_sr_built, x);
    }
    pthread_mutex_unlock(&g_sr_mu);
    return 0;
}

/* p_at_most_m_distinct_in_sample_of_x_from_n (enclone_help/src/help1.rs:292-298):
 * p = 1 - sum_{u=m+1..x} sr[x][u] * (u/n)^x * C(n,u), clamped at 0, in double-double.
 * [RECALL: the reference multiplies the x factors u/n first and the u factors
 * (n-v)/(u-v) after; in that order the running product falls into the denormal range
 * from x ~ 200 on (sr[x][x] ~ e^-x) and loses its low word.  The oracle evaluates the
 * SAME product with the small and large factors interleaved so that it stays in range:
 * the mathematically intended value, at double-double accuracy.  For the golden vector
 * (x = 33) both orders give the identical double.] */
double oracle_p_at_most_m_distinct(int m, int x, int n) {
    if (x <= 0) return 1.0;
    if (sr_ensure(x) != 0) return NAN;
    dd_t p = dd_from(1.0);
    const dd_t *row = g_sr_rows[x];
    for (int u = m + 1; u <= x; u++) {
        dd_t z = row[u];
        const dd_t r = dd_div(dd_from((double)u), dd_from((double)n));
        int small_left = x, v = 0;
        while (small_left > 0 || v < u) {
            if (v < u && (z.hi < 1.0 || small_left == 0)) {
                z = dd_mul(z, dd_div(dd_from((double)(n - v)), dd_from((double)(u - v))));
                v++;
            } else {
                z = dd_mul(z, r);
                small_left--;
            }
        }
        p = dd_sub(p, z);
    }
    double out = p.hi + p.lo;
    if (out < 0.0) out = 0.0;
    return out;
}

/* p1 of join_one: k total mutations, d shared, n = 3 * (V..J lengths) (help1.rs:292-296). */
double oracle_p1(uint32_t k, uint32_t d, uint32_t n) {
    return oracle_p_at_most_m_distinct((int)k - (int)d, (int)k, (int)n);
}

/* ------------------------------------------------------------------------------------------
 * EquivRel (crate equiv; API seen in-tree at enclone_tail/src/grouper.rs:406-445):
 * union-find over i32 ids.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t *parent; int32_t n; } equiv_t;
static void eq_init(equiv_t *e, int32_t n) {
    e->parent = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    e->n = n;
    for (int32_t i = 0; i < n; i++) e->parent[i] = i;
}
static void eq_free(equiv_t *e) { free(e->parent); e->parent = NULL; }
static int32_t eq_class_id(equiv_t *e, int32_t a) {
    int32_t r = a;
    while (e->parent[r] != r) r = e->parent[r];
    while (e->parent[a] != r) { int32_t nx = e->parent[a]; e->parent[a] = r; a = nx; }
    return r;
}
static void eq_join(equiv_t *e, int32_t a, int32_t b) {
    int32_t ra = eq_class_id(e, a), rb = eq_class_id(e, b);
    if (ra == rb) return;
    if (ra < rb) e->parent[rb] = ra; else e->parent[ra] = rb;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const ejoin_pack_t *pk;
    const ejoin_params_t *pr;
} octx_t;

static inline const uint8_t *tig(const ejoin_pack_t *pk, int c, uint32_t k) {
    return pk->tig_arena + pk->tig_off[c][k];
}
static inline const uint8_t *cdr3(const ejoin_pack_t *pk, int c, uint32_t k) {
    return pk->cdr3_arena + pk->cdr3_off[c][k];
}
static inline const uint8_t *refseq(const ejoin_pack_t *pk, uint32_t id) {
    return pk->refseq_arena + pk->refseq_off[id];
}

typedef struct {
    uint64_t pairs_candidate, pairs_evaluated, pairs_same_cdr3_len, pairs_tier2, pairs_accepted;
} ostats_t;

/* Heavy-chain region difference counts for gate m: out = {FWR1 diffs, FWR1 length,
 * CDR1+CDR2 diffs, CDR1+CDR2 length}.  Returns 0 (gate skipped) unless every region
 * start of both entries is known; k1's bounds are used [RECALL]. */
int oracle_fwr1_cdr12_counts(const ejoin_pack_t *pk, uint32_t k1, uint32_t k2, int out[4]) {
    const uint16_t f1 = pk->fr1_start[k1], c1 = pk->cdr1_start[k1], f2 = pk->fr2_start[k1],
                   c2 = pk->cdr2_start[k1], f3 = pk->fr3_start[k1];
    const int all1 = f1 != EJOIN_NONE16 && c1 != EJOIN_NONE16 && f2 != EJOIN_NONE16 &&
                     c2 != EJOIN_NONE16 && f3 != EJOIN_NONE16;
    const int all2 = pk->fr1_start[k2] != EJOIN_NONE16 && pk->cdr1_start[k2] != EJOIN_NONE16 &&
                     pk->fr2_start[k2] != EJOIN_NONE16 && pk->cdr2_start[k2] != EJOIN_NONE16 &&
                     pk->fr3_start[k2] != EJOIN_NONE16;
    const int L0 = pk->len[0][k1];
    if (!(all1 && all2 && f1 < c1 && c1 <= f2 && f2 <= c2 && c2 <= f3 && f3 <= L0 &&
          (f2 - c1) + (f3 - c2) > 0)) return 0;
    const uint8_t *a = tig(pk, 0, k1), *b = tig(pk, 0, k2);
    int dfw = 0, dc12 = 0;
    for (int p = f1; p < c1; p++) dfw += (a[p] != b[p]);
    for (int p = c1; p < f2; p++) dc12 += (a[p] != b[p]);
    for (int p = c2; p < f3; p++) dc12 += (a[p] != b[p]);
    out[0] = dfw; out[1] = c1 - f1; out[2] = dc12; out[3] = (f2 - c1) + (f3 - c2);
    return 1;
}

/* Number of positions at which the germline sequences assigned to k1 and k2 differ
 * (help1.rs:345-348, literal reading): V references compared left-aligned, J references
 * right-aligned, unmatched tail / head positions of the longer one count as differing. */
int oracle_ref_positions_differing(const ejoin_pack_t *pk, uint32_t k1, uint32_t k2) {
    int tot = 0;
    for (int c = 0; c < 2; c++) {
        const uint32_t v1 = pk->vs_seq[c][k1], v2 = pk->vs_seq[c][k2];
        if (v1 != v2) {
            const int l1 = (int)pk->refseq_len[v1], l2 = (int)pk->refseq_len[v2];
            const int m = l1 < l2 ? l1 : l2;
            const uint8_t *a = refseq(pk, v1), *b = refseq(pk, v2);
            for (int p = 0; p < m; p++) tot += (a[p] != b[p]);
            tot += (l1 > l2 ? l1 - l2 : l2 - l1);
        }
        const uint32_t j1 = pk->js_seq[c][k1], j2 = pk->js_seq[c][k2];
        if (j1 != j2) {
            const int l1 = (int)pk->refseq_len[j1], l2 = (int)pk->refseq_len[j2];
            const int m = l1 < l2 ? l1 : l2;
            const uint8_t *a = refseq(pk, j1) + (l1 - m), *b = refseq(pk, j2) + (l2 - m);
            for (int p = 0; p < m; p++) tot += (a[p] != b[p]);
            tot += (l1 > l2 ? l1 - l2 : l2 - l1);
        }
    }
    return tot;
}

/* join_one: the pairwise test.  SURVEY Appendix A.4; help1.rs:270-363.
 * Pure predicate of (k1,k2); fills *e on accept (and the count fields on every call that
 * reaches them).  Gate order is cheap-first; the order only affects logging in the
 * reference, never the result. */
int oracle_join_one(const ejoin_pack_t *pk, const ejoin_params_t *pr, uint32_t k1, uint32_t k2,
                    ejoin_edge_t *e, ostats_t *st) {
    memset(e, 0, sizeof(*e));
    e->k1 = k1; e->k2 = k2;
    /* a. chain-count gate (help1.rs:270-276): 2- or 3-chain exact subclonotypes only,
     *    and both info entries carry two chains. */
    if (pk->len[1][k1] == 0 || pk->len[1][k2] == 0) return 0;
    const uint32_t e1 = pk->exact_id[k1], e2 = pk->exact_id[k2];
    if (pk->nchains[e1] < 2 || pk->nchains[e1] > 3) return 0;
    if (pk->nchains[e2] < 2 || pk->nchains[e2] > 3) return 0;
    /* b. CDR3 length gate (help1.rs:278-282) */
    for (int c = 0; c < 2; c++)
        if (pk->cdr3_len[c][k1] != pk->cdr3_len[c][k2]) return 0;
    if (st) st->pairs_same_cdr3_len++;
    const int n0 = pk->cdr3_len[0][k1], n1 = pk->cdr3_len[1][k1];
    const int L[2] = { pk->len[0][k1], pk->len[1][k1] };

    /* c. total differences; capped only for TCR or when MAX_DIFFS is given
     *    (pages/cellranger.html.src:163-164; PREHISTORY:1771) [RECALL: condition] */
    int have_diffs = 0;
    uint32_t diffs = 0;
    if (!pk->is_bcr || pr->max_diffs < 1000000) {
        for (int c = 0; c < 2; c++) {
            const uint8_t *a = tig(pk, c, k1), *b = tig(pk, c, k2);
            for (int p = 0; p < L[c]; p++) diffs += (a[p] != b[p]);
        }
        have_diffs = 1;
        int cap = pr->max_diffs;
        if (!pk->is_bcr && pr->max_diffs_tcr < cap) cap = pr->max_diffs_tcr;
        if ((int)diffs > cap) return 0;
    }
    /* d. CDR3 differences per chain (help1.rs:300-304) */
    int cdc[2] = { 0, 0 };
    for (int c = 0; c < 2; c++) {
        const uint8_t *a = cdr3(pk, c, k1), *b = cdr3(pk, c, k2);
        const int n = c ? n1 : n0;
        for (int p = 0; p < n; p++) cdc[c] += (a[p] != b[p]);
    }
    const int cd = cdc[0] + cdc[1];
    if (!pk->is_bcr && cd > 0) return 0;              /* PREHISTORY:829 */
    if (cd > pr->max_cdr3_diffs) return 0;
    /*    CDR3 nucleotide identity >= JOIN_CDR3_IDENT (help1.rs:308-311) */
    {
        double ident = 100.0 * (1.0 - (double)cd / (double)(n0 + n1));
        if (ident < pr->join_cdr3_ident) return 0;
    }
    if (st) st->pairs_tier2++;
    /* e. donor gate (enclone_help/src/help2.rs:185; golden test256 l.29 "JOIN ERROR") */
    int err = 0;
    {
        uint32_t lo1 = pk->donor_lo[e1], hi1 = pk->donor_hi[e1];
        uint32_t lo2 = pk->donor_lo[e2], hi2 = pk->donor_hi[e2];
        if (lo1 != EJOIN_NONE && lo2 != EJOIN_NONE)
            err = !(lo1 == hi1 && lo2 == hi2 && lo1 == lo2);
        if (err && !pr->mix_donors) return 0;
    }
    /* f. light-chain constant regions differ and cd > 0 (help1.rs:354-356) */
    if (pk->is_bcr && !pr->old_light && cd > 0 &&
        pk->c_ref_id_light[k1] != pk->c_ref_id_light[k2]) return 0;
    /* g. V gene names differ and (V refs differ after right-truncation, or 5'UTR refs
     *    differ after left-truncation) (help1.rs:328-331) */
    for (int c = 0; c < 2; c++) {
        if (pk->vname_id[c][k1] == pk->vname_id[c][k2]) continue;
        uint32_t a = pk->vuni_seq[c][k1], b = pk->vuni_seq[c][k2];
        uint32_t la = pk->refseq_len[a], lb = pk->refseq_len[b];
        uint32_t m = la < lb ? la : lb;
        if (memcmp(refseq(pk, a), refseq(pk, b), m) != 0) return 0;
        uint32_t ua = pk->utr_seq[c][k1], ub = pk->utr_seq[c][k2];
        if (ua != EJOIN_NONE && ub != EJOIN_NONE) {
            uint32_t lua = pk->refseq_len[ua], lub = pk->refseq_len[ub];
            uint32_t mu = lua < lub ? lua : lub;
            if (memcmp(refseq(pk, ua) + (lua - mu), refseq(pk, ub) + (lub - mu), mu) != 0) return 0;
        }
    }
    /* h. reference choice (help1.rs:345-348; mechanics [RECALL]) */
    int nrefs = 1;
    for (int c = 0; c < 2; c++)
        if (pk->v_ref_id[c][k1] != pk->v_ref_id[c][k2] || pk->dref_id[c][k1] != pk->dref_id[c][k2] ||
            pk->j_ref_id[c][k1] != pk->j_ref_id[c][k2]) nrefs = 2;
    /* i. shares / indeps against each reference choice; windows exactly as
     *    print_utils5.rs:78-81:  V: p < |v| - ref_v_trim vs v[p];
     *                            J: p >= n - (|j| - ref_j_trim) vs j[|j| - (n - p)] */
    int shares[2] = { 0, 0 }, indeps[2] = { 0, 0 }, indeps_out[2] = { 0, 0 };
    for (int u = 0; u < nrefs; u++) {
        const uint32_t ku = u == 0 ? k1 : k2;
        for (int c = 0; c < 2; c++) {
            const uint8_t *t1 = tig(pk, c, k1), *t2 = tig(pk, c, k2);
            const int n = L[c];
            const uint32_t vi = pk->vs_seq[c][ku], ji = pk->js_seq[c][ku];
            const uint8_t *v = refseq(pk, vi), *j = refseq(pk, ji);
            const int vl = (int)pk->refseq_len[vi], jl = (int)pk->refseq_len[ji];
            /* CDR3 ranges of both entries (for the CDR3_MULT rule) */
            const int a1 = pk->cdr3_start[c][k1], a2 = pk->cdr3_start[c][k2];
            const int cl = pk->cdr3_len[c][k1];
            int vend = vl - pr->ref_v_trim;
            if (vend > n) vend = n;
            for (int p = 0; p < vend; p++) {
                const uint8_t r = v[p];
                const int out = !((p >= a1 && p < a1 + cl) || (p >= a2 && p < a2 + cl));
                if (t1[p] == t2[p]) {
                    if (t1[p] != r) { shares[u]++; e->shares_details[u][2 * c]++; }
                } else {
                    int q = (t1[p] != r) + (t2[p] != r);
                    indeps[u] += q;
                    if (out) indeps_out[u] += q;
                }
            }
            int jstart = n - (jl - pr->ref_j_trim);
            if (jstart < 0) jstart = 0;
            if (jl - pr->ref_j_trim > 0)
                for (int p = jstart; p < n; p++) {
                    const uint8_t r = j[jl - (n - p)];
                    const int out = !((p >= a1 && p < a1 + cl) || (p >= a2 && p < a2 + cl));
                    if (t1[p] == t2[p]) {
                        if (t1[p] != r) { shares[u]++; e->shares_details[u][2 * c + 1]++; }
                    } else {
                        int q = (t1[p] != r) + (t2[p] != r);
                        indeps[u] += q;
                        if (out) indeps_out[u] += q;
                    }
                }
        }
    }
    /*    MAX_DEGRADATION (help1.rs:345-348): "We do not join two clonotypes which were assigned
     *    different reference sequences unless those reference sequences differ by at most 2
     *    positions."  degradation_mode 0 (default) implements that sentence literally [DOC]:
     *    the number of positions at which the two entries' V references (left-aligned, as the
     *    V window indexes them) and J references (right-aligned, as the J window indexes them)
     *    differ, a length difference counting one per unmatched position, summed over both
     *    chains.  degradation_mode 1 is the builder's recollection of the upstream code
     *    [RECALL]: |shares[0] - shares[1]| > MAX_DEGRADATION.  SURVEY App. C-6. */
    if (nrefs == 2) {
        if (pr->degradation_mode == 1) {
            if (abs(shares[0] - shares[1]) > pr->max_degradation) return 0;
        } else {
            if (oracle_ref_positions_differing(pk, k1, k2) > pr->max_degradation) return 0;
        }
    }
    int d = shares[0], mi = indeps[0], mo = indeps_out[0];
    if (nrefs == 2) {
        if (shares[1] < d) d = shares[1];
        if (indeps[1] < mi) mi = indeps[1];
        if (indeps_out[1] < mo) mo = indeps_out[1];
    }
    /* j. p1 (help1.rs:292-298) */
    const int k = mi + 2 * d;
    const uint32_t n = 3u * (uint32_t)(L[0] + L[1]);
    const double p1 = oracle_p1((uint32_t)k, (uint32_t)d, n);
    /* k. N (help1.rs:300-306); floating-point order verified against the golden
     *    (SURVEY Appendix B): sum over chains of (42*cd_c)/n_c, then pow. */
    double mult;
    if (pr->old_mult) {
        /* OLD_MULT: sum_{i<=cd} C(3*cn, i), cn = total CDR3 aa (cellranger.html.src:142-144) */
        int cn = (n0 + n1) / 3;
        double s = 0.0, t = 1.0;
        for (int i = 0; i <= cd; i++) {
            s += t;
            t = t * (double)(3 * cn - i) / (double)(i + 1);
        }
        mult = s;
    } else {
        double cdx = 0.0;
        cdx += (double)pr->cdr3_normal_len * (double)cdc[0] / (double)n0;
        cdx += (double)pr->cdr3_normal_len * (double)cdc[1] / (double)n1;
        mult = pow(pr->mult_pow, cdx);
    }
    /* l. score and the three ways to accept (help1.rs:313-344) */
    const double score = p1 * mult;
    int reason = 0;
    if (!(score > pr->max_score)) reason |= 1;
    if (d >= pr->auto_share) reason |= 2;
    {
        int hc = pk->hcomp[k1] < pk->hcomp[k2] ? pk->hcomp[k1] : pk->hcomp[k2];
        if (hc - cd >= pr->comp_filt) {
            if (!have_diffs) {
                for (int c = 0; c < 2; c++) {
                    const uint8_t *a = tig(pk, c, k1), *b = tig(pk, c, k2);
                    for (int p = 0; p < L[c]; p++) diffs += (a[p] != b[p]);
                }
                have_diffs = 1;
            }
            if ((int)diffs - cd <= pr->comp_filt_bound) reason |= 4;
        }
    }
    if (!reason) return 0;
    /* m. heavy-chain FWR1 identity vs CDR1+CDR2 identity (help1.rs:357-359; golden l.47-54) */
    if (pk->is_bcr) {
        int rc[4];
        if (oracle_fwr1_cdr12_counts(pk, k1, k2, rc)) {
            const double ifw = 100.0 * (1.0 - (double)rc[0] / (double)rc[1]);
            const double i12 = 100.0 * (1.0 - (double)rc[2] / (double)rc[3]);
            if (ifw - i12 >= pr->fwr1_cdr12_delta) return 0;
        }
    }
    /* n. CDR3 concentration (help1.rs:360-363): cd >= CDR3_MULT * max(1, non-shared
     *    mutations outside CDR3) */
    {
        int x = mo > 1 ? mo : 1;
        if (cd >= pr->cdr3_mult * x) return 0;
    }
    /* o. the PotentialJoin */
    if (!have_diffs) {
        for (int c = 0; c < 2; c++) {
            const uint8_t *a = tig(pk, c, k1), *b = tig(pk, c, k2);
            for (int p = 0; p < L[c]; p++) diffs += (a[p] != b[p]);
        }
    }
    e->nrefs = (uint16_t)nrefs;
    e->cd = (uint16_t)cd; e->cd_chain[0] = (uint16_t)cdc[0]; e->cd_chain[1] = (uint16_t)cdc[1];
    e->diffs = diffs;
    for (int u = 0; u < 2; u++) { e->shares[u] = u < nrefs ? shares[u] : 0; e->indeps[u] = u < nrefs ? indeps[u] : 0; }
    e->k = (uint16_t)k; e->d = (uint16_t)d; e->n = n;
    e->err = (uint8_t)err; e->accept_reason = (uint8_t)reason;
    e->p1 = p1; e->mult = mult; e->score = score;
    if (st) st->pairs_accepted++;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * One bucket: join_core (A.3) + two-cell passes (A.6) + ordered emit (A.7).
 * ---------------------------------------------------------------------------------------- */
typedef struct { ejoin_edge_t *v; size_t n, cap; } evec_t;
static void ev_push(evec_t *x, const ejoin_edge_t *e) {
    if (x->n == x->cap) {
        x->cap = x->cap ? x->cap * 2 : 16;
        x->v = (ejoin_edge_t *)realloc(x->v, x->cap * sizeof(ejoin_edge_t));
    }
    x->v[x->n++] = *e;
}

/* ------------------------------------------------------------------------------------------
 * Tile-parallel join_core for ONE large bucket (test infrastructure for the 200,000-entry
 * configuration, where the sequential scan below would take hours on one core).
 * join_one is a pure predicate of (k1,k2), so `pot` = the replay of
 * join_core's skip-if-same-class rule over the FULL accept set in (k1,k2) order.  Rows k1 are
 * scored by worker threads; a pair is handed to oracle_join_one only when it passes a packed
 * restatement of join_one's own first gates (b: CDR3 lengths equal; d: TCR cd = 0,
 * cd <= MAX_CDR3_DIFFS, CDR3 identity >= JOIN_CDR3_IDENT with the same double expression) --
 * pairs it drops are pairs oracle_join_one returns 0 for.  tests/test_oracle_kat.py checks that
 * this mode and the sequential scan give identical edges on the same buckets.
 * ---------------------------------------------------------------------------------------- */
static uint32_t g_parallel_bucket_min = 0xFFFFFFFFu;     /* off unless a test asks for it */
static int g_parallel_threads = 1;
static uint32_t g_parallel_row_limit = 0xFFFFFFFFu;     /* bench only: score just the first rows of the triangle (a bounded timing sample) */
void oracle_set_row_limit(uint32_t rows) { g_parallel_row_limit = rows ? rows : 0xFFFFFFFFu; }
void oracle_set_parallel_bucket(uint32_t min_entries, int threads) {
    g_parallel_bucket_min = min_entries ? min_entries : 0xFFFFFFFFu;
    g_parallel_threads = threads < 1 ? 1 : threads;
}

typedef struct {
    const ejoin_pack_t *pk; const ejoin_params_t *pr;
    uint32_t i, j;
    const uint64_t *packed;      /* [nb][pw] 2-bit CDR3s (both chains concatenated), NULL = no prefilter */
    const uint8_t *packable;     /* [nb] CDR3s are ACGT only */
    uint32_t pw;
    _Atomic uint32_t *next_row;
    evec_t acc;
    ostats_t st;
} prow_t;

static int cdr3_prefilter_pass(const prow_t *w, uint32_t k1, uint32_t k2) {
    const ejoin_pack_t *pk = w->pk; const ejoin_params_t *pr = w->pr;
    if (pk->cdr3_len[0][k1] != pk->cdr3_len[0][k2] || pk->cdr3_len[1][k1] != pk->cdr3_len[1][k2]) return 0;
    if (!w->packed || !w->packable[k1 - w->i] || !w->packable[k2 - w->i]) return 1;
    const uint64_t *a = w->packed + (size_t)(k1 - w->i) * w->pw, *b = w->packed + (size_t)(k2 - w->i) * w->pw;
    int cd = 0;
    for (uint32_t q = 0; q < w->pw; q++) {
        const uint64_t z = a[q] ^ b[q];
        cd += __builtin_popcountll((z | (z >> 1)) & 0x5555555555555555ull);
    }
    if (!pk->is_bcr && cd > 0) return 0;
    if (cd > pr->max_cdr3_diffs) return 0;
    const int tot = pk->cdr3_len[0][k1] + pk->cdr3_len[1][k1];
    const double ident = 100.0 * (1.0 - (double)cd / (double)tot);
    if (ident < pr->join_cdr3_ident) return 0;
    return 1;
}

static void *prow_main(void *arg) {
    prow_t *w = (prow_t *)arg;
    ejoin_edge_t e;
    for (;;) {
        const uint32_t r0 = atomic_fetch_add(w->next_row, 64);
        uint32_t rows = w->j - w->i;
        if (g_parallel_row_limit < rows) rows = g_parallel_row_limit;
        if (r0 >= rows) break;
        const uint32_t r1 = r0 + 64 < rows ? r0 + 64 : rows;
        for (uint32_t k1 = w->i + r0; k1 < w->i + r1; k1++)
            for (uint32_t k2 = k1 + 1; k2 < w->j; k2++) {
                if (!cdr3_prefilter_pass(w, k1, k2)) continue;
                w->st.pairs_evaluated++;
                if (oracle_join_one(w->pk, w->pr, k1, k2, &e, &w->st)) ev_push(&w->acc, &e);
            }
    }
    return NULL;
}

static int edge_cmp(const void *a, const void *b) {
    const ejoin_edge_t *x = (const ejoin_edge_t *)a, *y = (const ejoin_edge_t *)b;
    if (x->k1 != y->k1) return x->k1 < y->k1 ? -1 : 1;
    if (x->k2 != y->k2) return x->k2 < y->k2 ? -1 : 1;
    return 0;
}

/* fills `pot` with join_core's result for the bucket [i,j) */
static void join_core_parallel(const ejoin_pack_t *pk, const ejoin_params_t *pr, uint32_t i, uint32_t j, evec_t *pot, ostats_t *st) {
    const uint32_t nb = j - i;
    uint32_t maxlen = 0;
    for (uint32_t k = i; k < j; k++) { uint32_t t = (uint32_t)pk->cdr3_len[0][k] + pk->cdr3_len[1][k]; if (t > maxlen) maxlen = t; }
    const uint32_t pw = (maxlen + 31) / 32 ? (maxlen + 31) / 32 : 1;
    uint64_t *packed = (uint64_t *)calloc((size_t)nb * pw, sizeof(uint64_t));
    uint8_t *packable = (uint8_t *)malloc(nb);
    for (uint32_t k = i; k < j; k++) {
        uint64_t *row = packed + (size_t)(k - i) * pw;
        uint32_t pos = 0; int ok = 1;
        for (int c = 0; c < 2; c++) {
            const uint8_t *q = cdr3(pk, c, k);
            for (uint32_t p = 0; p < pk->cdr3_len[c][k]; p++, pos++) {
                uint64_t code;
                switch (q[p]) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break; default: code = 0; ok = 0; }
                row[pos >> 5] |= code << (2 * (pos & 31));
            }
        }
        packable[k - i] = (uint8_t)ok;
    }
    const int T = g_parallel_threads;
    prow_t *ws = (prow_t *)calloc((size_t)T, sizeof(prow_t));
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    _Atomic uint32_t next_row = 0;
    for (int t = 0; t < T; t++) {
        ws[t].pk = pk; ws[t].pr = pr; ws[t].i = i; ws[t].j = j; ws[t].packed = packed; ws[t].packable = packable; ws[t].pw = pw;
        ws[t].next_row = &next_row;
    }
    if (T == 1) prow_main(&ws[0]);
    else {
        for (int t = 0; t < T; t++) pthread_create(&th[t], NULL, prow_main, &ws[t]);
        for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    }
    evec_t all = { 0, 0, 0 };
    for (int t = 0; t < T; t++) {
        for (size_t q = 0; q < ws[t].acc.n; q++) ev_push(&all, &ws[t].acc.v[q]);
        free(ws[t].acc.v);
        st->pairs_evaluated += ws[t].st.pairs_evaluated; st->pairs_same_cdr3_len += ws[t].st.pairs_same_cdr3_len;
        st->pairs_tier2 += ws[t].st.pairs_tier2; st->pairs_accepted += ws[t].st.pairs_accepted;
    }
    if (all.n) qsort(all.v, all.n, sizeof(ejoin_edge_t), edge_cmp);
    /* the replay: exactly join_core's rule, applied to the accepted pairs in scan order */
    equiv_t eq;
    eq_init(&eq, (int32_t)nb);
    for (size_t q = 0; q < all.n; q++) {
        const int32_t a = (int32_t)(all.v[q].k1 - i), b = (int32_t)(all.v[q].k2 - i);
        if (eq_class_id(&eq, a) == eq_class_id(&eq, b)) continue;
        eq_join(&eq, a, b);
        ev_push(pot, &all.v[q]);
    }
    eq_free(&eq);
    free(all.v); free(ws); free(th); free(packed); free(packable);
}

static void bucket_join(const ejoin_pack_t *pk, const ejoin_params_t *pr, uint32_t i, uint32_t j,
                        evec_t *out, ostats_t *st) {
    const int32_t nb = (int32_t)(j - i);
    st->pairs_candidate += (uint64_t)nb * (uint64_t)(nb - 1) / 2;
    if (nb < 2) return;
    evec_t pot = { 0, 0, 0 };
    equiv_t eq;
    if ((uint32_t)nb >= g_parallel_bucket_min) join_core_parallel(pk, pr, i, j, &pot, st);
    else {
    /* join_core: upper-triangle loop, skipping pairs already in one class
     * (pages/heuristics.html.src:30-32; idiom grouper.rs:406-412) */
    eq_init(&eq, nb);
    ejoin_edge_t e;
    for (uint32_t k1 = i; k1 < j; k1++)
        for (uint32_t k2 = k1 + 1; k2 < j; k2++) {
            if (eq_class_id(&eq, (int32_t)(k1 - i)) == eq_class_id(&eq, (int32_t)(k2 - i))) continue;
            st->pairs_evaluated++;
            if (oracle_join_one(pk, pr, k1, k2, &e, st)) {
                ev_push(&pot, &e);
                eq_join(&eq, (int32_t)(k1 - i), (int32_t)(k2 - i));
            }
        }
    eq_free(&eq);
    }
    /* two-cell rule, applied in two passes (help1.rs:349-353; PREHISTORY:14607) */
    int32_t *size = (int32_t *)malloc(sizeof(int32_t) * (size_t)nb);
    for (int pass = 0; pass < 2; pass++) {
        eq_init(&eq, nb);
        for (size_t t = 0; t < pot.n; t++) eq_join(&eq, (int32_t)(pot.v[t].k1 - i), (int32_t)(pot.v[t].k2 - i));
        memset(size, 0, sizeof(int32_t) * (size_t)nb);
        for (int32_t a = 0; a < nb; a++) size[eq_class_id(&eq, a)]++;
        size_t w = 0;
        for (size_t t = 0; t < pot.n; t++) {
            const ejoin_edge_t *q = &pot.v[t];
            int keep = 1;
            if (size[eq_class_id(&eq, (int32_t)(q->k1 - i))] == 2) {
                uint32_t cells = pk->ncells[pk->exact_id[q->k1]] + pk->ncells[pk->exact_id[q->k2]];
                if (cells == 2 && !pr->easy && !(q->cd <= q->d)) keep = 0;
            }
            if (keep) pot.v[w++] = *q;
        }
        pot.n = w;
        eq_free(&eq);
    }
    free(size);
    /* ordered emit: skip if already joined, else join and record (A.7) */
    eq_init(&eq, nb);
    for (size_t t = 0; t < pot.n; t++) {
        int32_t a = (int32_t)(pot.v[t].k1 - i), b = (int32_t)(pot.v[t].k2 - i);
        if (eq_class_id(&eq, a) == eq_class_id(&eq, b)) continue;
        eq_join(&eq, a, b);
        ev_push(out, &pot.v[t]);
    }
    eq_free(&eq);
    free(pot.v);
}

typedef struct {
    const ejoin_pack_t *pk;
    const ejoin_params_t *pr;
    const uint32_t *bstart;
    uint32_t nb;
    _Atomic uint32_t *next;
    evec_t *per_bucket;
    uint32_t stride, phase;
    ostats_t st;
} worker_t;

static void *worker_main(void *arg) {
    worker_t *w = (worker_t *)arg;
    for (;;) {
        uint32_t b = atomic_fetch_add(w->next, 1);
        if (b >= w->nb) break;
        if (b % w->stride != w->phase) continue;
        bucket_join(w->pk, w->pr, w->bstart[b], w->bstart[b + 1], &w->per_bucket[b], &w->st);
    }
    return NULL;
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

typedef struct {
    uint64_t n_buckets, pairs_candidate, pairs_evaluated, pairs_same_cdr3_len, pairs_tier2, pairs_accepted;
    double ms_total;
    int threads;
} oracle_stats_t;

/* bucket scan (A.2): maximal runs of equal lens.  Returns malloc'ed starts [nb+1]. */
uint32_t *oracle_bucket_scan(const ejoin_pack_t *pk, uint32_t *nb_out) {
    uint32_t n = pk->n_info, nb = 0;
    uint32_t *bs = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n + 2));
    for (uint32_t k = 0; k < n; k++)
        if (k == 0 || pk->len[0][k] != pk->len[0][k - 1] || pk->len[1][k] != pk->len[1][k - 1]) bs[nb++] = k;
    bs[nb] = n;
    *nb_out = nb;
    return bs;
}

/* join_exacts + finish_join.  bucket_lo/bucket_hi restrict the work to a bucket range
 * (used to time a bounded sample); pass 0, UINT32_MAX for everything. */
int oracle_join_exacts_strided(const ejoin_pack_t *pk, const ejoin_params_t *pr, int threads,
                               uint32_t bucket_lo, uint32_t bucket_hi, uint32_t stride, uint32_t phase,
                               ejoin_result_t *out, oracle_stats_t *ost);
int oracle_join_exacts(const ejoin_pack_t *pk, const ejoin_params_t *pr, int threads,
                       uint32_t bucket_lo, uint32_t bucket_hi, ejoin_result_t *out,
                       oracle_stats_t *ost) {
    return oracle_join_exacts_strided(pk, pr, threads, bucket_lo, bucket_hi, 1, 0, out, ost);
}
/* stride/phase: only buckets b with b % stride == phase are joined (a representative
 * bounded sample for timing); the rest are left unjoined. */
int oracle_join_exacts_strided(const ejoin_pack_t *pk, const ejoin_params_t *pr, int threads,
                               uint32_t bucket_lo, uint32_t bucket_hi, uint32_t stride, uint32_t phase,
                               ejoin_result_t *out, oracle_stats_t *ost) {
    memset(out, 0, sizeof(*out));
    double t0 = now_ms();
    uint32_t nb_all;
    uint32_t *bs_all = oracle_bucket_scan(pk, &nb_all);
    if (bucket_hi > nb_all) bucket_hi = nb_all;
    if (bucket_lo > bucket_hi) bucket_lo = bucket_hi;
    const uint32_t nb = bucket_hi - bucket_lo;
    const uint32_t *bs = bs_all + bucket_lo;
    if (threads < 1) threads = 1;
    /* the sr table is an input of join_exacts in the reference: build it before timing
     * the join proper is the caller's business (oracle_prepare) */
    evec_t *per_bucket = (evec_t *)calloc(nb ? nb : 1, sizeof(evec_t));
    _Atomic uint32_t next = 0;
    worker_t *ws = (worker_t *)calloc((size_t)threads, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        ws[t].pk = pk; ws[t].pr = pr; ws[t].bstart = bs; ws[t].nb = nb; ws[t].next = &next;
        ws[t].per_bucket = per_bucket; ws[t].stride = stride ? stride : 1; ws[t].phase = phase;
    }
    if (threads == 1) worker_main(&ws[0]);
    else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker_main, &ws[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    /* raw_joins = per-bucket join lists concatenated in bucket order (A.7) */
    size_t ne = 0;
    for (uint32_t b = 0; b < nb; b++) ne += per_bucket[b].n;
    out->edges = (ejoin_edge_t *)malloc(sizeof(ejoin_edge_t) * (ne ? ne : 1));
    out->n_edges = ne;
    size_t w = 0;
    for (uint32_t b = 0; b < nb; b++) {
        if (per_bucket[b].n) memcpy(out->edges + w, per_bucket[b].v, per_bucket[b].n * sizeof(ejoin_edge_t));
        w += per_bucket[b].n;
        free(per_bucket[b].v);
    }
    free(per_bucket);
    /* finish_join (A.7): global EquivRel; join emitted edges; then join the info entries
     * that share a clonotype_id (PREHISTORY:14747) */
    const uint32_t n = pk->n_info;
    equiv_t eq;
    eq_init(&eq, (int32_t)n);
    for (size_t t = 0; t < ne; t++) eq_join(&eq, (int32_t)out->edges[t].k1, (int32_t)out->edges[t].k2);
    {
        int64_t *first = (int64_t *)malloc(sizeof(int64_t) * (size_t)(pk->n_exact ? pk->n_exact : 1));
        for (uint32_t x = 0; x < pk->n_exact; x++) first[x] = -1;
        for (uint32_t k = 0; k < n; k++) {
            uint32_t x = pk->exact_id[k];
            if (first[x] < 0) first[x] = k; else eq_join(&eq, (int32_t)first[x], (int32_t)k);
        }
        free(first);
    }
    out->class_of = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    out->n_info = n;
    uint32_t ncls = 0;
    for (uint32_t k = 0; k < n; k++) {
        out->class_of[k] = (uint32_t)eq_class_id(&eq, (int32_t)k); /* root = min member (eq_join keeps the smaller) */
        if (out->class_of[k] == k) ncls++;
    }
    out->n_classes = ncls;
    eq_free(&eq);
    ostats_t tot = { 0, 0, 0, 0, 0 };
    for (int t = 0; t < threads; t++) {
        tot.pairs_candidate += ws[t].st.pairs_candidate; tot.pairs_evaluated += ws[t].st.pairs_evaluated;
        tot.pairs_same_cdr3_len += ws[t].st.pairs_same_cdr3_len; tot.pairs_tier2 += ws[t].st.pairs_tier2;
        tot.pairs_accepted += ws[t].st.pairs_accepted;
    }
    out->stats.n_buckets = nb;
    out->stats.pairs_candidate = tot.pairs_candidate;
    out->stats.pairs_compared = tot.pairs_same_cdr3_len;
    out->stats.pairs_tier2 = tot.pairs_tier2;
    out->stats.pairs_accepted = tot.pairs_accepted;
    out->stats.ms_total = now_ms() - t0;
    if (ost) {
        ost->n_buckets = nb; ost->pairs_candidate = tot.pairs_candidate; ost->pairs_evaluated = tot.pairs_evaluated;
        ost->pairs_same_cdr3_len = tot.pairs_same_cdr3_len; ost->pairs_tier2 = tot.pairs_tier2;
        ost->pairs_accepted = tot.pairs_accepted; ost->ms_total = out->stats.ms_total; ost->threads = threads;
    }
    free(ws); free(th); free(bs_all);
    return 0;
}

/* Evaluate join_one on every pair of a bucket WITHOUT the skip (the full accept graph);
 * used by tests to check the GPU's accepted-pair set, not only the spanning forest. */
int oracle_accept_graph(const ejoin_pack_t *pk, const ejoin_params_t *pr, uint32_t i, uint32_t j,
                        ejoin_edge_t **edges, uint64_t *n_edges) {
    evec_t v = { 0, 0, 0 };
    ejoin_edge_t e;
    for (uint32_t k1 = i; k1 < j; k1++)
        for (uint32_t k2 = k1 + 1; k2 < j; k2++)
            if (oracle_join_one(pk, pr, k1, k2, &e, NULL)) ev_push(&v, &e);
    *edges = v.v; *n_edges = v.n;
    return 0;
}

void oracle_prepare(int kmax) { sr_ensure(kmax < 64 ? 64 : kmax); }
void oracle_result_free(ejoin_result_t *r) { free(r->edges); free(r->class_of); memset(r, 0, sizeof(*r)); }
void oracle_free(void *p) { free(p); }

/* Statistic for design work: how many info entries have at least one partner in their bucket that
 * passes the chain-count, CDR3-length and CDR3-identity gates (i.e. would reach the full scoring).
 * Single-threaded scan of every pair; not part of the join. */
uint64_t oracle_count_entries_with_cdr3_partner(const ejoin_pack_t *pk, const ejoin_params_t *pr) {
    uint32_t nb;
    uint32_t *bs = oracle_bucket_scan(pk, &nb);
    uint8_t *hit = (uint8_t *)calloc(pk->n_info ? pk->n_info : 1, 1);
    for (uint32_t b = 0; b < nb; b++)
        for (uint32_t k1 = bs[b]; k1 < bs[b + 1]; k1++) {
            if (pk->len[1][k1] == 0) continue;
            uint32_t e1 = pk->exact_id[k1];
            if (pk->nchains[e1] < 2 || pk->nchains[e1] > 3) continue;
            const int n0 = pk->cdr3_len[0][k1], n1 = pk->cdr3_len[1][k1];
            for (uint32_t k2 = k1 + 1; k2 < bs[b + 1]; k2++) {
                if (hit[k1] && hit[k2]) continue;
                if (pk->cdr3_len[0][k2] != n0 || pk->cdr3_len[1][k2] != n1) continue;
                uint32_t e2 = pk->exact_id[k2];
                if (pk->nchains[e2] < 2 || pk->nchains[e2] > 3) continue;
                int cd = 0;
                const uint8_t *a = cdr3(pk, 0, k1), *c = cdr3(pk, 0, k2);
                for (int p = 0; p < n0; p++) cd += a[p] != c[p];
                a = cdr3(pk, 1, k1); c = cdr3(pk, 1, k2);
                for (int p = 0; p < n1; p++) cd += a[p] != c[p];
                if (!pk->is_bcr && cd > 0) continue;
                if (cd > pr->max_cdr3_diffs) continue;
                if (100.0 * (1.0 - (double)cd / (double)(n0 + n1)) < pr->join_cdr3_ident) continue;
                hit[k1] = hit[k2] = 1;
            }
        }
    uint64_t n = 0;
    for (uint32_t k = 0; k < pk->n_info; k++) n += hit[k];
    free(hit); free(bs);
    return n;
}
