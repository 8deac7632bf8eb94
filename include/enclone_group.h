/* enclone_group.h — C ABI of the symmetric clonotype grouping (libenclone_join.so), SURVEY.md §8(f)3.
 *
 * Replaces the hot part of enclone_tail::grouper::grouper(), "Case 1: symmetric grouping"
 * (/root/reference/enclone_tail/src/grouper.rs:55-766): the all-pairs distance passes with an EquivRel inside every
 * group (heavy_pc / light_pc :386-457, cdr3_heavy_pc_hf :459-535, aa_heavy_pc / aa_light_pc :537-612, cdr3_heavy_pc /
 * cdr3_light_pc :614-687, cdr3_aa_heavy_pc / cdr3_aa_light_pc :689-758) run on the GPU (bit-parallel Levenshtein, one
 * thread per clonotype pair, GPU union-find); the key partitions (:64-384) are host-side sorts.  The result is the
 * partition of the final EquivRel `e` (:760-766) as min-member labels; keeper_group, the ordering of groups by cell
 * count stay with the caller.  Case 2 (asymmetric grouping) is egroup_run_asym below.  Each distance pass reproduces the reference's own orbit idiom,
 * including `ee.orbit(i)` for i < #orbits (:441-452; see oracle/grouper_oracle.c).
 *
 * Flat view of (exacts, exact_clonotypes, rsi, opt_d_val, refdata.name): clonotype i = exacts[i] (a list of exact
 * subclonotype ids); exact subclonotype e = exact_clonotypes[e].share (a list of chains).
 */
#ifndef ENCLONE_GROUP_H
#define ENCLONE_GROUP_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGROUP_ABI_VERSION 1

typedef struct egroup_pack {
    uint32_t abi_version;
    uint32_t n_clono;                /* exacts.len()                                                */
    uint32_t n_exact;                /* exact_clonotypes.len()                                      */
    uint32_t n_chain;                /* total number of chains over all exact subclonotypes         */
    const uint32_t *clono_exact_off; /* [n_clono+1] exacts[i] = clono_exact[off[i] .. off[i+1])       */
    const uint32_t *clono_exact;
    const uint32_t *exact_chain_off; /* [n_exact+1] share[j] of exact e = chain records off[e] + j  */
    /* per chain (TigData1) */
    const uint8_t  *left;            /* share[j].left (heavy / TRB)                                 */
    const uint64_t *seq_off;         /* share[j].seq: ASCII nucleotides (ACGT, N, -) in seq_arena   */
    const uint32_t *seq_len;
    const uint32_t *seq_del_len;     /* share[j].seq_del.len()                                      */
    const uint32_t *cdr3_start;      /* share[j].cdr3_start (in seq coordinates)                    */
    const uint64_t *cdr3_aa_off;     /* share[j].cdr3_aa: ASCII amino acids in aa_arena             */
    const uint32_t *cdr3_aa_len;
    const uint8_t  *seq_arena;  uint64_t seq_arena_bytes;
    const uint8_t  *aa_arena;   uint64_t aa_arena_bytes;
    /* per clonotype: the columns of rsi[i] (ColInfo), names as interned ids of refdata.name[..] */
    const uint32_t *col_off;         /* [n_clono+1]                                                 */
    const uint32_t *col_vname;       /* refdata.name[rsi[i].vids[col]]                              */
    const uint32_t *col_jname;       /* refdata.name[rsi[i].jids[col]]                              */
    const uint32_t *col_dname;       /* refdata.name[opt_d_val[i].1[col][0][0]] or the id of "null"  */
    const uint8_t  *col_left;        /* rsi[i].left[col]                                            */
} egroup_pack_t;

/* ctl.clono_group_opt: flags for the key partitions; percentages for the distance passes (negative = not asked for) */
typedef struct egroup_opts {
    int32_t vj_refname, vdj_refname, v_heavy_refname, vj_heavy_refname, vdj_heavy_refname;
    int32_t vj_len, cdr3_len, cdr3_heavy_len, cdr3_light_len;
    int32_t reserved[3];
    double  heavy_pc, light_pc;
    double  cdr3_heavy_pc_hf;        /* .0 of cdr3_heavy_pc_hf                                      */
    const double *hf_matrix;         /* .1: 20 x 20 penalties, amino acids in the order ACDEFGHIKLMNPQRSTVWY */
    double  aa_heavy_pc, aa_light_pc;
    double  cdr3_heavy_pc, cdr3_light_pc;
    double  cdr3_aa_heavy_pc, cdr3_aa_light_pc;
} egroup_opts_t;

typedef struct egroup_stats {
    uint64_t pairs_candidate;        /* clonotype pairs inside groups, summed over the distance passes */
    uint64_t pairs_evaluated;        /* of those, pairs whose chains were actually compared (not yet in one class) */
    uint64_t chain_compares;         /* Levenshtein / penalty evaluations                            */
    uint64_t gpu_launches;
    uint64_t n_groups;               /* classes of the final partition                              */
    double   ms_device;              /* CUDA events around the distance passes                      */
    double   ms_total;
} egroup_stats_t;

void egroup_opts_default(egroup_opts_t *o);      /* nothing asked for */
/* The whole symmetric grouping.  group_of[i] = smallest clonotype index of i's group.  Fails with EJOIN_E_NOGPU (-3)
 * without a CUDA device when a distance pass is asked for: there is no CPU fallback for those. */
int  egroup_run(const egroup_pack_t *pack, const egroup_opts_t *opts, int device, uint32_t *group_of, egroup_stats_t *stats);
const char *egroup_last_error(void);

/* ---- Case 2: asymmetric grouping (grouper.rs:828-928; AGROUP, AG_CENTER, AG_DIST_FORMULA=cdr3_edit_distance, AG_DIST_BOUND) ----
 * center = the clonotypes with in_center[i] != 0 (:832-853).  distance(center c, clonotype j) = min(1e9, min over the pairs of
 * exact subclonotypes of (heavy + light)), heavy / light = the smallest Levenshtein distance between heavy / light CDR3 amino acid
 * strings of the two, 1e9 when there is no such pair of chains (:856-889).  Per center the clonotypes are sorted by (distance,
 * index); with top=N the sorted positions 0..N are taken, with max=D the ones with distance <= D; the center itself is listed
 * first with distance 0 and skipped in the walk; groups smaller than min_group are dropped (:891-925).  The distance matrix and
 * the selection run on the GPU; the result arrays belong to the library until egroup_asym_free. */
#define EGROUP_BOUND_TOP 0
#define EGROUP_BOUND_MAX 1
#define EGROUP_BOUND_NONE 2          /* neither: every clonotype is in every group */
typedef struct egroup_asym_opts {
    int32_t  bound_kind;
    uint32_t top;                    /* AG_DIST_BOUND=top=N */
    double   max_dist;               /* AG_DIST_BOUND=max=D */
    uint32_t min_group;              /* ctl.clono_group_opt.min_group */
    uint32_t reserved;
} egroup_asym_opts_t;
typedef struct egroup_asym_result {
    uint64_t  n_groups;
    uint64_t *group_off;             /* [n_groups + 1] into member / dist */
    uint32_t *member;                /* member[group_off[g]] is the group's center */
    double   *dist;                  /* the number printed as "distance = {}" */
    void     *owner;
} egroup_asym_result_t;
int  egroup_run_asym(const egroup_pack_t *pack, const uint8_t *in_center, const egroup_asym_opts_t *opts, int device,
                     egroup_asym_result_t *out, egroup_stats_t *stats);
void egroup_asym_free(egroup_asym_result_t *r);

#ifdef __cplusplus
}
#endif
#endif /* ENCLONE_GROUP_H */
