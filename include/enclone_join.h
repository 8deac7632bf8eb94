/* enclone_join.h — C ABI of the B200 clonotype-join library (libenclone_join.so).
 *
 * This is the drop-in boundary for ONE hot path of 10XGenomics/enclone: the call
 *
 *   enclone::join::join_exacts(is_bcr, to_bc, refdata, ctl, exact_clonotypes, info,
 *                              raw_joins, sr, dref) -> EquivRel
 *
 * made once per run from enclone_stuff::start::main_enclone_start, which the in-tree
 * code reaches at enclone_main/src/main_enclone.rs:36 (the callee crates are the
 * un-vendored git dependency enclone_ranger@cbd15ef, Cargo.lock:632-634; see
 * SURVEY.md §0.2 and §8(b)).  There is no FFI in the reference; these entry points
 * are what a thin Rust `extern "C"` crate would bind (INTEGRATION.md shows the stub).
 *
 * Conventions: plain pointers and sizes, little-endian PODs, no callbacks, no
 * exceptions across the boundary; every function returns 0 or a negative EJOIN_E_*.
 * The caller owns everything reachable from ejoin_pack_t for the duration of the
 * call; the library owns everything reachable from ejoin_result_t until
 * ejoin_result_free().  One ejoin_run*() at a time per handle.
 */
#ifndef ENCLONE_JOIN_H
#define ENCLONE_JOIN_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EJOIN_ABI_VERSION 2
#define EJOIN_NONE 0xFFFFFFFFu   /* Option::None for u32 ids                      */
#define EJOIN_NONE16 0xFFFFu     /* Option::None for u16 region bounds            */
/* ejoin_pack_t.flags: tig_arena is pinned host memory the device can read (ejoin_host_alloc /
 * ejoin_host_register), padded with >= 64 readable bytes after tig_arena_bytes: the library then
 * reads the V..J bytes IN PLACE over PCIe, and only for the entries that have a CDR3-compatible
 * partner (about half of a repertoire), instead of uploading the whole arena. */
#define EJOIN_F_TIGS_IN_PLACE 1u
/* The V..J bases of every chain WITHOUT a deletion marker are given 2 bits per base in tig2_arena, in the layout the
 * reference already holds them in (CloneInfo.tigsp: debruijn 0.3.4 DnaString, Cargo.lock:527-530 — 32 bases per u64,
 * base i of a chain in word i >> 5 at bits 63-2(i&31) (high) and 62-2(i&31) (low), A,C,G,T = 0,1,2,3; every chain
 * starts on a word boundary) — the same split join_one makes: ndiffs on tigsp when !has_del, bytes otherwise.
 * tig_off[c][k] (or entry_rec[k].tig_off[c]) is then a WORD offset into tig2_arena for those chains; chains with
 * has_del[c][k] != 0 (deletion marker '-', or any byte outside ACGT) stay ASCII: tig_off is a byte offset into
 * tig_arena, which needs to hold only those chains.  With EJOIN_F_TIGS_IN_PLACE the in-place reads go to tig2_arena. */
#define EJOIN_F_TIGS_2BIT 2u
/* With EJOIN_F_TIGS_2BIT: every word of tig2_arena holds its 32 bases as the two bit planes of their 2-bit codes instead of
 * interleaved: bits 63..32 = the codes' high bits, bits 31..0 = the codes' low bits, base i at bit 31-i of each half — i.e.
 * (pext(w, 0xAAAAAAAAAAAAAAAA) << 32) | pext(w, 0x5555555555555555) of the DnaString word w.  Same size, same offsets; the GPU has
 * no pext, the host does it in two instructions per word, and the packing kernel then skips a 45-instruction bit transpose. */
#define EJOIN_F_TIGS_PLANES 16u
/* The per-entry fields that only the scoring reads are given as one 64-byte record per entry (entry_rec) instead of
 * 22 parallel arrays (which may then be NULL): 4.5x fewer bytes per entry and one cache line per entry.
 * Requires every id / germline index to be < 65535. */
#define EJOIN_F_ENTRY_RECS 4u
/* The CDR3s come 2 bits per base: for entry k the two CDR3 strings concatenated (chain 0 then chain 1, no gap), in the
 * DnaString bit layout, in ceil((cdr3_len[0][k] + cdr3_len[1][k]) / 32) words starting at cdr3_2bit[cdr3_2bit_off[k]].
 * cdr3_arena / cdr3_off may then be NULL.  (CDR3 characters outside ACGT are outside the documented limits either way.) */
#define EJOIN_F_CDR3_2BIT 8u

/* error codes */
#define EJOIN_OK 0
#define EJOIN_E_ARG (-1)         /* bad argument / inconsistent pack              */
#define EJOIN_E_CUDA (-2)        /* CUDA runtime error (message in last_error)    */
#define EJOIN_E_NOGPU (-3)       /* no CUDA device: the product path has no CPU fallback */
#define EJOIN_E_UNSUPPORTED (-4) /* input outside documented limits               */
#define EJOIN_E_NOMEM (-5)

/* ---------------------------------------------------------------------------
 * JoinPack: the flattened view of (exact_clonotypes, info, refdata, dref) that
 * join_exacts reads.  Field provenance: SURVEY.md §8(a) table T.
 * Index k runs over `info` (Vec<CloneInfo>) IN THE CALLER'S SORTED ORDER
 * (CloneInfo's derived Ord, `lens` first); chain c = 0 is the heavy/TRB chain
 * (TigData1.left), c = 1 the light/TRA chain.  An entry with one chain
 * (CloneInfo.lens.len() == 1) has len[1][k] == 0 and is never joined
 * (enclone_help/src/help1.rs:270-276).
 * Arena layout: offsets are free within an arena, except that the sharded entry points (ejoin_run_shard*,
 * ejoin_upload_shard, ejoin_run_sharded) upload only the slice between a shard's first and last entry, so for those
 * the arenas must be laid out in entry order (offsets non-decreasing in k — what any packer that walks `info` once
 * writes).  An entry whose bytes fall outside its slice, or outside the arena, is an EJOIN_E_ARG error.
 * ------------------------------------------------------------------------- */
/* One info entry's "late" fields (EJOIN_F_ENTRY_RECS): same meaning as the arrays of the same name below;
 * 0xFFFF stands for None (EJOIN_NONE / EJOIN_NONE16). */
typedef struct ejoin_entry_rec {
    uint32_t tig_off[2];         /* word offset into tig2_arena (2-bit chain) or byte offset into tig_arena (ASCII chain) */
    uint16_t cdr3_start[2];
    uint16_t v_ref_id[2], j_ref_id[2], dref_id[2], vs_seq[2], js_seq[2], vname_id[2], vuni_seq[2], utr_seq[2];
    uint16_t c_ref_id_light, hcomp;
    uint16_t fr1_start, cdr1_start, fr2_start, cdr2_start, fr3_start;
    uint8_t  has_del[2];
    uint8_t  pad[4];
} ejoin_entry_rec_t;             /* 64 bytes */

typedef struct ejoin_pack {
    uint32_t abi_version;        /* EJOIN_ABI_VERSION                             */
    uint32_t n_info;             /* info.len()                                    */
    uint32_t n_exact;            /* exact_clonotypes.len()                        */
    uint32_t n_refseq;           /* entries in the germline sequence table        */
    uint32_t is_bcr;             /* join_exacts(is_bcr, ..)                       */
    uint32_t flags;              /* EJOIN_F_*                                     */

    /* ---- per info entry ---- */
    const uint32_t *exact_id;    /* [n_info] CloneInfo.clonotype_id               */
    const uint16_t *len[2];      /* [n_info] CloneInfo.lens[c] = tigs[c].len()    */
    const uint64_t *tig_off[2];  /* [n_info] byte offset of tigs[c] in tig_arena  */
    const uint8_t  *tig_arena;   /* CloneInfo.tigs: ASCII V..J, '-' marks a deletion */
    uint64_t        tig_arena_bytes;
    const uint8_t  *has_del[2];  /* [n_info] CloneInfo.has_del[c]                 */
    const uint16_t *cdr3_len[2]; /* [n_info] CloneInfo.cdr3s[c].len() (nt)        */
    const uint64_t *cdr3_off[2]; /* [n_info] byte offset of cdr3s[c] in cdr3_arena */
    const uint8_t  *cdr3_arena;  /* CloneInfo.cdr3s: ASCII, ACGT only             */
    uint64_t        cdr3_arena_bytes;
    const uint16_t *cdr3_start[2]; /* [n_info] TigData1.cdr3_start (in V..J coords) */
    const uint32_t *v_ref_id[2]; /* [n_info] CloneInfo.vsids[c] (universal ref id) */
    const uint32_t *j_ref_id[2]; /* [n_info] CloneInfo.jsids[c]                   */
    const uint32_t *dref_id[2];  /* [n_info] CloneInfo.dref[c], EJOIN_NONE if no donor allele */
    const uint32_t *vs_seq[2];   /* [n_info] refseq index of CloneInfo.vs[c] (donor allele / indel-edited V actually compared) */
    const uint32_t *js_seq[2];   /* [n_info] refseq index of CloneInfo.js[c]      */
    const uint32_t *vname_id[2]; /* [n_info] id of refdata.name[v_ref_id] with any trailing "*.." removed */
    const uint32_t *vuni_seq[2]; /* [n_info] refseq index of refdata.refs[v_ref_id] (universal V) */
    const uint32_t *utr_seq[2];  /* [n_info] refseq index of refdata.refs[u_ref_id], EJOIN_NONE if u_ref_id is None */
    const uint32_t *c_ref_id_light; /* [n_info] share[light].c_ref_id, EJOIN_NONE if None */
    const uint16_t *hcomp;       /* [n_info] share[heavy].jun.hcomp               */
    /* heavy-chain region starts in V..J coordinates (TigData1.fr1_start, cdr1_start, ...); EJOIN_NONE16 = None */
    const uint16_t *fr1_start, *cdr1_start, *fr2_start, *cdr2_start, *fr3_start;

    /* ---- per exact subclonotype ---- */
    const uint32_t *nchains;     /* [n_exact] ExactClonotype.share.len()          */
    const uint32_t *ncells;      /* [n_exact] ExactClonotype.ncells()             */
    const uint32_t *donor_lo;    /* [n_exact] min over cells of clones[i][0].donor_index, EJOIN_NONE if all None */
    const uint32_t *donor_hi;    /* [n_exact] max over cells, EJOIN_NONE if all None */

    /* ---- germline sequence table (ASCII): universal V/J refs, donor alleles,
     *      indel-edited V refs, 5'UTR refs ---- */
    const uint64_t *refseq_off;  /* [n_refseq] byte offset into refseq_arena      */
    const uint32_t *refseq_len;  /* [n_refseq]                                    */
    const uint8_t  *refseq_arena;
    uint64_t        refseq_arena_bytes;

    /* ---- compact forms (ABI 2; see EJOIN_F_TIGS_2BIT / EJOIN_F_ENTRY_RECS) ---- */
    const uint64_t *tig2_arena;  /* 2-bit V..J bases (CloneInfo.tigsp), >= 8 readable words of slack behind the last one */
    uint64_t        tig2_arena_words;
    const ejoin_entry_rec_t *entry_rec; /* [n_info] */
    const uint64_t *cdr3_2bit;   /* EJOIN_F_CDR3_2BIT */
    uint64_t        cdr3_2bit_words;
    const uint32_t *cdr3_2bit_off; /* [n_info] word offset into cdr3_2bit */
} ejoin_pack_t;

/* Join parameters: EncloneControl.{join_alg_opt, heur, clono_filt_opt_def} as
 * scalars (defaults from enclone_help/src/help1.rs:300-363
 * and pages/cellranger.html.src:55-247).  ejoin_params_default() fills them. */
typedef struct ejoin_params {
    double  max_score;           /* MAX_SCORE, 100000                              */
    double  mult_pow;            /* MULT_POW, 80                                   */
    double  join_cdr3_ident;     /* JOIN_CDR3_IDENT, 85 (%)                        */
    double  fwr1_cdr12_delta;    /* 20 (%)                                         */
    int32_t cdr3_normal_len;     /* CDR3_NORMAL_LEN, 42                            */
    int32_t auto_share;          /* AUTO_SHARES, 15                                */
    int32_t comp_filt;           /* hcomp - cd >= 8                                */
    int32_t comp_filt_bound;     /* diffs outside junction <= 80                   */
    int32_t max_cdr3_diffs;      /* MAX_CDR3_DIFFS, 1000 (off)                     */
    int32_t cdr3_mult;           /* CDR3_MULT, 5                                   */
    int32_t max_degradation;     /* MAX_DEGRADATION, 2                             */
    int32_t max_diffs;           /* MAX_DIFFS, 1000000 (off for BCR)               */
    int32_t max_diffs_tcr;       /* total-difference cap applied when !is_bcr, 5 */
    int32_t ref_v_trim;          /* 15                                             */
    int32_t ref_j_trim;          /* 15                                             */
    int32_t easy;                /* EASY                                           */
    int32_t old_light;           /* OLD_LIGHT                                      */
    int32_t old_mult;            /* OLD_MULT: N = sum_{i<=cd} C(3 cn, i)              */
    int32_t mix_donors;          /* MIX_DONORS (clono_filt_opt_def.donor)          */
    int32_t degradation_mode;    /* how MAX_DEGRADATION is read (help1.rs:345-348):
                                  * 0 = the documented rule: the references assigned to the two entries may differ in at most
                                  *     MAX_DEGRADATION positions (V left-aligned, J right-aligned, both chains summed);
                                  * 1 = |shares[0] - shares[1]| <= MAX_DEGRADATION (recollection of the upstream code) */
    int32_t reserved[4];
} ejoin_params_t;

/* One emitted join = one PotentialJoin that survived. */
typedef struct ejoin_edge {
    uint32_t k1, k2;             /* info indices, k1 < k2                          */
    int32_t  shares[2];          /* per reference choice u (u < nrefs)             */
    int32_t  indeps[2];
    uint16_t shares_details[2][4]; /* [u][V_heavy, J_heavy, V_light, J_light]       */
    uint16_t nrefs;              /* 1 or 2                                         */
    uint16_t cd;                 /* CDR3 nt differences, both chains               */
    uint16_t cd_chain[2];
    uint32_t diffs;              /* total V..J differences, both chains            */
    uint16_t k, d;               /* k = min indeps + 2 d, d = min shares           */
    uint32_t n;                  /* 3 * (len0 + len1)                              */
    uint8_t  err;                /* cross-donor join under MIX_DONORS ("JOIN ERROR") */
    uint8_t  accept_reason;      /* bit0 score<=max_score, bit1 shares>=auto_share, bit2 hcomp exception */
    uint16_t reserved;
    double   p1, mult, score;
} ejoin_edge_t;

typedef struct ejoin_stats {
    uint64_t n_buckets;          /* runs of equal lens                             */
    uint64_t n_groups;           /* (lens, cdr3 lens) comparison groups            */
    uint64_t pairs_candidate;    /* sum over buckets of n_b (n_b-1)/2 : the metric's "pairs" */
    uint64_t pairs_compared;     /* pairs whose CDR3s were XOR/popcount-compared (tier 1)    */
    uint64_t pairs_tier2;        /* tier-1 survivors fully scored                  */
    uint64_t pairs_accepted;     /* pairs passing join_one                         */
    uint64_t pairs_generic;      /* tier-2 pairs routed to the bytewise path       */
    uint64_t edges_after_prune;  /* accepted edges kept for the ordered replay     */
    uint64_t gpu_launches;       /* kernels launched by this call                  */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t alg_bytes_tier1;    /* sum over compared pairs of the two c3 rows read (2 * 8W bytes)   */
    uint64_t alg_bytes_tier2;    /* sum over tier-2 pairs of the two t2 rows + meta records read     */
    uint64_t alg_bytes_pack;     /* ASCII bytes read + packed rows and meta written by the pack kernel */
    double   ms_h2d, ms_device, ms_d2h, ms_host_post, ms_total; /* wall, host clock */
    double   ms_kernel_tier1, ms_kernel_tier2, ms_kernel_pack, ms_kernel_other; /* CUDA events */
    uint64_t alg_bytes_pack_t2;  /* k_pack_t2 alone: V..J bytes read + t2 row bytes written, needed entries only */
    double   ms_kernel_pack_t2;  /* k_pack_t2 alone (CUDA events); with a chunked or in-place upload it includes the wait for the bytes */
    uint64_t mma_pairs;          /* of pairs_compared, the pairs of groups whose tier 1 ran on the tensor cores (k_sweep_mma): large groups whose
                                  * two CDR3s total at most 96 bases; alg_bytes_tier1 counts the other groups only */
    uint64_t mma_macs;           /* int8 multiply-accumulates of those groups' tile pairs (128 x 128 x K each, K = 3 * bases rounded up to 32) */
} ejoin_stats_t;

typedef struct ejoin_result {
    uint64_t      n_edges;
    ejoin_edge_t *edges;         /* raw_joins, in the reference's emit order       */
    uint32_t     *class_of;      /* [n_info] EquivRel partition as min-member labels (finish_join); NULL when n_info == 0 */
    uint32_t      n_info;
    uint32_t      n_classes;
    void         *owner;         /* non-NULL: edges/class_of live in that handle's pinned result buffer and stay valid
                                    until its next ejoin_run*/ /*ejoin_finish or ejoin_destroy (one live result per handle) */
    ejoin_stats_t stats;
} ejoin_result_t;

/* Raw accepted edge as produced on one device before the ordered replay; the unit
 * gathered across GPUs (the only collective). */
typedef struct ejoin_raw_edge {
    uint32_t k1, k2;
    uint16_t cd, d;              /* for the two-cell rule                          */
    uint32_t flags;
} ejoin_raw_edge_t;

typedef struct ejoin_handle ejoin_handle_t;

void ejoin_params_default(ejoin_params_t *p);

/* device = CUDA ordinal.  Fails with EJOIN_E_NOGPU when no device is usable. */
int  ejoin_create(ejoin_handle_t **h, const ejoin_params_t *p, int device);
void ejoin_destroy(ejoin_handle_t *h);
const char *ejoin_last_error(const ejoin_handle_t *h);

/* The whole join stage on one GPU: upload (overlapped with the first kernels), pack, score, the spanning
 * forest + two-cell rule + ordered raw_joins on the device, per-edge payload, union-find, D2H into the
 * handle's pinned result buffer.  Equivalent of join_exacts + finish_join.  `out` points into that buffer:
 * valid until the next call on the handle. */
int  ejoin_run(ejoin_handle_t *h, const ejoin_pack_t *pack, ejoin_result_t *out);
void ejoin_result_free(ejoin_result_t *r);
/* Per-stage device times: after ejoin_set_stage_timing(h, 1) every pipeline run records a CUDA event on the launching
 * stream between its stages (a stage = one kernel or a short group: "k_sweep", "k_pack_t2", "sort entries", ...), and
 * ejoin_stage_time returns the time of stage `index` (0, 1, ...) of the last run, EJOIN_E_ARG past the last stage.
 * Off by default: the events cost a few microseconds per stage. */
int  ejoin_set_stage_timing(ejoin_handle_t *h, int on);
int  ejoin_stage_time(const ejoin_handle_t *h, int index, const char **name, double *ms);

/* Multi-GPU decomposition (one process per GPU):
 *   every rank: ejoin_plan_shards() -> which buckets this rank owns (contiguous, cost-balanced ranges),
 *               ejoin_run_shard()   -> accepted raw edges for those buckets (global info ids),
 *   all-gather the raw edges (NCCL; done by the host framework),
 *   then       ejoin_finish()       -> ordered replay, per-edge payload, partition. */
int  ejoin_plan_shards(const ejoin_pack_t *pack, int world_size, uint32_t *bucket_owner /* [n_buckets] */,
                       uint32_t *bucket_start /* [n_buckets+1] */, uint32_t cap_buckets, uint32_t *n_buckets);
int  ejoin_run_shard(ejoin_handle_t *h, const ejoin_pack_t *pack, int rank, int world_size,
                     ejoin_raw_edge_t **edges, uint64_t *n_edges, ejoin_stats_t *stats);
void ejoin_raw_free(ejoin_raw_edge_t *edges);
/* Same decomposition, one call per rank.  *mode = 0 (contiguous bucket ranges: the rank's buckets are
 * independent of the others, so `final` receives their emitted joins with full payload, class_of NULL,
 * and no raw edges are returned) or 1 (stride mode, a giant bucket is shared: `raw` receives the
 * accepted edges for the gather + ejoin_finish, `final` stays empty). */
int  ejoin_run_shard_final(ejoin_handle_t *h, const ejoin_pack_t *pack, int rank, int world_size, int *mode,
                           ejoin_result_t *final, ejoin_raw_edge_t **raw, uint64_t *n_raw, ejoin_stats_t *stats);
int  ejoin_finish(ejoin_handle_t *h, const ejoin_pack_t *pack, const ejoin_raw_edge_t *edges,
                  uint64_t n_edges, ejoin_result_t *out);

/* The same decomposition with the exchange step inside the library (NCCL, bound at run time from libnccl.so.2): one
 * process per GPU, each with its own handle.  Rank 0 obtains an id (ejoin_comm_id), the host framework hands it to every
 * rank (any transport: MPI, a file, torch.distributed), every rank calls ejoin_comm_init, then ejoin_run_sharded once per
 * join.  Contiguous bucket ranges: each rank's emitted joins, with their PotentialJoin records, go device to device to
 * rank 0, which runs finish_join on its GPU and returns the whole result (raw_joins order, class_of over all entries);
 * the other ranks return an empty result with their own stats.  A giant bucket shared by the ranks (stride mode): the
 * accepted raw edges are gathered to rank 0 and replayed there. */
#define EJOIN_COMM_ID_BYTES 128
int  ejoin_comm_id(uint8_t *id /* [EJOIN_COMM_ID_BYTES] */);
int  ejoin_comm_init(ejoin_handle_t *h, const uint8_t *id, int rank, int world_size);
int  ejoin_run_sharded(ejoin_handle_t *h, const ejoin_pack_t *pack, ejoin_result_t *out);

/* finish_join on the host from an emitted edge list (k1[i], k2[i]): partition labels (min member)
 * over all n_info entries, including the same-exact-subclonotype links.  Used by rank 0 after
 * the gather; needs no GPU and no handle. */
int  ejoin_partition(const ejoin_pack_t *pack, const uint32_t *k1, const uint32_t *k2, uint64_t n_edges,
                     uint32_t *class_of /* [n_info] */, uint32_t *n_classes);

/* The same on the handle's GPU (uploads exact_id and the edge list: ~9 bytes per info entry). */
int  ejoin_partition_gpu(ejoin_handle_t *h, const ejoin_pack_t *pack, const uint32_t *k1, const uint32_t *k2, uint64_t n_edges,
                         uint32_t *class_of /* [n_info] */, uint32_t *n_classes);

/* Device-resident variant used to time the device side alone: upload once, then
 * ejoin_run_resident() repeats pack+score+compact+union-find on the resident input. */
int  ejoin_upload(ejoin_handle_t *h, const ejoin_pack_t *pack);
int  ejoin_upload_shard(ejoin_handle_t *h, const ejoin_pack_t *pack, int rank, int world_size);   /* this rank's share, as ejoin_run_shard* would take it */
int  ejoin_run_resident(ejoin_handle_t *h, ejoin_stats_t *stats);

/* Host-side conversion of an ASCII / parallel-array pack into the compact form (EJOIN_F_TIGS_2BIT | EJOIN_F_ENTRY_RECS |
 * EJOIN_F_CDR3_2BIT | EJOIN_F_TIGS_IN_PLACE): what a caller that holds CloneInfo.tigsp would build directly.  The arenas
 * of the new pack are pinned (ejoin_host_alloc); arrays that do not change are shared with `in`, which must outlive it.
 * Fails with EJOIN_E_UNSUPPORTED when an id does not fit 16 bits.  Needs no handle. */
int  ejoin_pack_compact(const ejoin_pack_t *in, int threads, ejoin_pack_t **out);
void ejoin_pack_compact_free(ejoin_pack_t *p);

/* Pinned host memory for the caller's flat buffers (so H2D runs at link speed). */
void *ejoin_host_alloc(size_t bytes);
void  ejoin_host_free(void *p);
/* Pin caller-owned buffers in place (e.g. the Vec<u8> arenas the Rust side already holds). */
int   ejoin_host_register(void *p, size_t bytes);
int   ejoin_host_unregister(void *p);

/* JoinInfo.log of one emitted join (enclone_tail/src/group.rs:70-74,434-441): the text block enclone prints under
 * PRINT_JOINS-style options, rebuilt from the edge record and the caller's inputs; golden:
 * enclone_exec/testx/inputs/outputs/enclone_test256_output:29-70, reproduced byte for byte by tests/test_join_log.py.
 * share_pos_v / share_pos_j are recomputed on the host for that one edge.  `names` carries the strings only the caller
 * has.  Writes at most cap-1 bytes + NUL; *len = full length.  Host code: needs no GPU and no handle. */
typedef struct ejoin_log_names {
    int32_t index;               /* "[0] = ..."                                                  */
    const char *id1, *id2;       /* "<dataset>.clono<exact id>" of the two entries               */
    const char *vsegs1, *vsegs2; /* "107=ENST00000632630, 219=ENST00000632910"                   */
    const char *dataset1, *dataset2; /* "flaky8b = flaky8b"                                      */
    const char *cdr3_aa1, *cdr3_aa2; /* "IGK:CMQSTHWPTWTF;IGH:CAFDFW"                            */
    const char *bcs1, *bcs2;     /* barcodes, comma separated                                    */
} ejoin_log_names_t;
int  ejoin_format_join_log(const ejoin_pack_t *pack, const ejoin_params_t *params, const ejoin_edge_t *edge,
                           const ejoin_log_names_t *names, char *buf, size_t cap, size_t *len);

/* The probability table entry the device uses: P(at most k-d distinct values in k
 * draws with replacement from n) — exposed so tests can pin it to the known-answer
 * vector (enclone_exec/testx/inputs/outputs/enclone_test256_output:55-56). */
double ejoin_p1(uint32_t k, uint32_t d, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* ENCLONE_JOIN_H */
