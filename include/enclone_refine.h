/* enclone_refine.h — C ABI of the post-join refinement (libenclone_join.so), SURVEY.md §8(f)2.
 *
 * Replaces the data-parallel part of what enclone does with `eq` and `raw_joins` right after join_exacts returns
 * (enclone_main/src/main_enclone.rs:36 -> enclone_stuff::start, and process_clonotypes, enclone_main/src/stop.rs:56; both
 * in the un-vendored enclone_ranger crate, Cargo.lock:632-634):
 *   - clonotype chain grouping ("define_mat"): which chains of which exact subclonotypes share a column of the clonotype
 *     table.  Spec: pages/heuristics.html.src:19-41.
 *   - pure subclonotypes, the doublet filter, the signature filter, the weak-chain filter.  Spec:
 *     pages/default_filters.html.src:292-353 (filters 13, 14 and 16 of the table at :28-150); option names NDOUBLET, NSIG,
 *     NWEAK_CHAINS (enclone_tail/src/fate.rs:21,76,81).
 *   - the break-up of clonotypes that are "no longer held together by shared chains" after a filter
 *     (default_filters.html.src:315-317).
 * Out of this boundary: the filters that need per-cell data (UMI, UMI ratio, GEX, whitelist, cross, barcode duplication,
 * quality), the onesie merger and everything that prints.
 *
 * Flat view of (exact_clonotypes, info, raw_joins, eq).  Strings never cross the boundary: the caller interns them
 * (equal id <=> equal string), as it does for the join's reference ids.
 */
#ifndef ENCLONE_REFINE_H
#define ENCLONE_REFINE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EREFINE_ABI_VERSION 1
#define EREFINE_NONE 0xFFFFFFFFu
#define EREFINE_MAX_CHAINS 8         /* chains per exact subclonotype (enclone's maximum-contigs filter keeps <= 4) */

/* fate of an exact subclonotype (result.fate) */
#define EREFINE_KEPT 0
#define EREFINE_DOUBLET 1
#define EREFINE_SIGNATURE 2
#define EREFINE_WEAK_CHAIN 3

typedef struct erefine_pack {
    uint32_t abi_version;
    uint32_t n_info;                 /* info.len()                                                      */
    uint32_t n_exact;                /* exact_clonotypes.len()                                          */
    uint32_t n_chain;                /* sum of exact_clonotypes[x].share.len()                          */
    /* ---- per info entry (CloneInfo) ---- */
    const uint32_t *exact_id;        /* [n_info] CloneInfo.clonotype_index                              */
    const uint8_t  *exact_cols[2];   /* [n_info] CloneInfo.exact_cols[c]: which share[] entries the two chains of the entry
                                      * are; exact_cols[1][k] = 0xFF for a one-chain entry              */
    const uint32_t *class_of;        /* [n_info] the EquivRel finish_join returned (ejoin_result_t.class_of) */
    /* ---- raw_joins, in the join's emit order (ejoin_result_t.edges[].k1 / .k2) ---- */
    uint64_t        n_joins;
    const uint32_t *join_k1, *join_k2;
    /* ---- per exact subclonotype ---- */
    const uint32_t *chain_off;       /* [n_exact+1] share[m] of exact x = chain record chain_off[x] + m  */
    const uint32_t *ncells;          /* [n_exact] ExactClonotype.ncells()                               */
    const uint32_t *origin_rank;     /* [n_exact] rank of CloneInfo.origin (the sorted origin ids of its cells) among all
                                      * distinct origin vectors, in their lexicographic order: the exact subclonotypes of a
                                      * clonotype are listed by (origin, exact id)                      */
    /* ---- per chain (TigData1 = exact_clonotypes[x].share[m]) ---- */
    const uint32_t *seq_id;          /* equal <=> share[m].seq (V..J nucleotides) equal                 */
    const uint32_t *cdr3_id;         /* equal <=> share[m].cdr3_dna equal                               */
    const uint32_t *order_rank;      /* rank of (chain_type with TRB -> TRX, TRA -> TRY, ':', cdr3_aa ; seq.len()) among all
                                      * chains (equal keys, equal rank): the order of the table's columns */
    const uint8_t  *left;            /* share[m].left                                                   */
    const uint32_t *seq_len;
    const uint64_t *seq_off;         /* ASCII V..J in seq_arena (read by the third-chain rule only)     */
    const uint8_t  *seq_arena;
    uint64_t        seq_arena_bytes;
} erefine_pack_t;

/* clono_filt_opt_def.{doublet, signature, weak_chains} and the constants of the three rules */
typedef struct erefine_opts {
    int32_t doublet;                 /* 0 = NDOUBLET                                                    */
    int32_t signature;               /* 0 = NSIG                                                        */
    int32_t weak_chains;             /* 0 = NWEAK_CHAINS                                                */
    int32_t doublet_mult;            /* 5:  5 * ncells(p0) <= min(ncells(p1), ncells(p2))                */
    int32_t signature_mult;          /* 20: cells of the two-chain pure subclonotypes sharing a chain >= 20 * ncells(p) */
    int32_t weak_max_cells;          /* 20 (5 for the Cell Ranger of the day)                           */
    int32_t weak_mult;               /* 8:  8 * cells supporting the chain < cells of the clonotype      */
    int32_t third_chain_diffs;       /* 10: two joined three-chain exact subclonotypes: third chains of equal length
                                      * differing in at most 10 bases share a column                    */
    int32_t reserved[4];
} erefine_opts_t;

typedef struct erefine_stats {
    uint64_t n_clonotypes_in;        /* classes of class_of, counted over exact subclonotypes            */
    uint64_t n_clonotypes_out;       /* clonotypes after the three filters and the break-ups             */
    uint64_t n_pure;                 /* pure subclonotypes before the doublet filter                     */
    uint64_t n_doublet, n_signature, n_weak;   /* exact subclonotypes deleted by each filter             */
    uint64_t triples_tested;         /* (p0, p1, p2) candidates examined by the doublet filter           */
    uint64_t gpu_launches;
    uint64_t h2d_bytes;              /* bytes uploaded: the tables, and of seq_arena only the third chains the rule compares */
    double   ms_device, ms_total;    /* device: CUDA events around the device work (the host's gather of those bytes taken out) */
} erefine_stats_t;

/* Caller-allocated result arrays. */
typedef struct erefine_result {
    uint8_t  *fate;                  /* [n_exact] EREFINE_KEPT / _DOUBLET / _SIGNATURE / _WEAK_CHAIN      */
    uint32_t *clonotype_of;          /* [n_exact] smallest exact id of the exact subclonotype's final clonotype; EREFINE_NONE
                                      * for a deleted one and for one without an info entry              */
    uint32_t *column_of;             /* [n_chain] column of the chain in its final clonotype's table (mat[col][u] = Some(m)
                                      * <=> column_of[chain_off[x] + m] = col), EREFINE_NONE for a deleted exact subclonotype */
    uint32_t *n_columns;             /* [n_exact] mat.len() of the exact subclonotype's final clonotype   */
} erefine_result_t;

void erefine_opts_default(erefine_opts_t *o);    /* all three filters on, enclone's constants */
/* The whole refinement on one GPU.  No CPU fallback: EJOIN_E_NOGPU (-3) without a CUDA device.  Returns 0 or a negative
 * EJOIN_E_* code (include/enclone_join.h); erefine_last_error() has the message. */
int  erefine_run(const erefine_pack_t *pack, const erefine_opts_t *opts, int device, erefine_result_t *out, erefine_stats_t *stats);
const char *erefine_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ENCLONE_REFINE_H */
