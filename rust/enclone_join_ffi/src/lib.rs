//! Raw bindings to `libenclone_join.so` (include/enclone_join.h, ABI version 2).
//! NOT COMPILED HERE: the build image has no Rust toolchain.  Field order and types mirror the C
//! header one for one; `tests/test_cpu.py::test_library_exports_every_declared_symbol` keeps the
//! header and the library's export list in step, and the same layouts are exercised through ctypes
//! (`enclone_b200/_abi.py`) and C++ (`enclone_b200/host/join_exacts.hpp`).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const EJOIN_ABI_VERSION: u32 = 2;
pub const EJOIN_NONE: u32 = 0xFFFF_FFFF;
pub const EJOIN_NONE16: u16 = 0xFFFF;
pub const EJOIN_F_TIGS_IN_PLACE: u32 = 1;
pub const EJOIN_F_TIGS_2BIT: u32 = 2;
pub const EJOIN_F_ENTRY_RECS: u32 = 4;
pub const EJOIN_F_CDR3_2BIT: u32 = 8;
pub const EJOIN_F_TIGS_PLANES: u32 = 16;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ejoin_entry_rec_t {
    pub tig_off: [u32; 2], pub cdr3_start: [u16; 2],
    pub v_ref_id: [u16; 2], pub j_ref_id: [u16; 2], pub dref_id: [u16; 2], pub vs_seq: [u16; 2], pub js_seq: [u16; 2],
    pub vname_id: [u16; 2], pub vuni_seq: [u16; 2], pub utr_seq: [u16; 2],
    pub c_ref_id_light: u16, pub hcomp: u16,
    pub fr1_start: u16, pub cdr1_start: u16, pub fr2_start: u16, pub cdr2_start: u16, pub fr3_start: u16,
    pub has_del: [u8; 2], pub pad: [u8; 4],
}

#[repr(C)]
pub struct ejoin_pack_t {
    pub abi_version: u32, pub n_info: u32, pub n_exact: u32, pub n_refseq: u32, pub is_bcr: u32, pub flags: u32,
    pub exact_id: *const u32,
    pub len: [*const u16; 2], pub tig_off: [*const u64; 2], pub tig_arena: *const u8, pub tig_arena_bytes: u64,
    pub has_del: [*const u8; 2], pub cdr3_len: [*const u16; 2], pub cdr3_off: [*const u64; 2], pub cdr3_arena: *const u8,
    pub cdr3_arena_bytes: u64, pub cdr3_start: [*const u16; 2],
    pub v_ref_id: [*const u32; 2], pub j_ref_id: [*const u32; 2], pub dref_id: [*const u32; 2],
    pub vs_seq: [*const u32; 2], pub js_seq: [*const u32; 2], pub vname_id: [*const u32; 2], pub vuni_seq: [*const u32; 2],
    pub utr_seq: [*const u32; 2], pub c_ref_id_light: *const u32, pub hcomp: *const u16,
    pub fr1_start: *const u16, pub cdr1_start: *const u16, pub fr2_start: *const u16, pub cdr2_start: *const u16, pub fr3_start: *const u16,
    pub nchains: *const u32, pub ncells: *const u32, pub donor_lo: *const u32, pub donor_hi: *const u32,
    pub refseq_off: *const u64, pub refseq_len: *const u32, pub refseq_arena: *const u8, pub refseq_arena_bytes: u64,
    pub tig2_arena: *const u64, pub tig2_arena_words: u64, pub entry_rec: *const ejoin_entry_rec_t,
    pub cdr3_2bit: *const u64, pub cdr3_2bit_words: u64, pub cdr3_2bit_off: *const u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ejoin_params_t {
    pub max_score: f64, pub mult_pow: f64, pub join_cdr3_ident: f64, pub fwr1_cdr12_delta: f64,
    pub cdr3_normal_len: i32, pub auto_share: i32, pub comp_filt: i32, pub comp_filt_bound: i32, pub max_cdr3_diffs: i32,
    pub cdr3_mult: i32, pub max_degradation: i32, pub max_diffs: i32, pub max_diffs_tcr: i32, pub ref_v_trim: i32, pub ref_j_trim: i32,
    pub easy: i32, pub old_light: i32, pub old_mult: i32, pub mix_donors: i32, pub degradation_mode: i32, pub reserved: [i32; 4],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ejoin_edge_t {
    pub k1: u32, pub k2: u32, pub shares: [i32; 2], pub indeps: [i32; 2], pub shares_details: [[u16; 4]; 2],
    pub nrefs: u16, pub cd: u16, pub cd_chain: [u16; 2], pub diffs: u32, pub k: u16, pub d: u16, pub n: u32,
    pub err: u8, pub accept_reason: u8, pub reserved: u16, pub p1: f64, pub mult: f64, pub score: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ejoin_stats_t {
    pub n_buckets: u64, pub n_groups: u64, pub pairs_candidate: u64, pub pairs_compared: u64, pub pairs_tier2: u64,
    pub pairs_accepted: u64, pub pairs_generic: u64, pub edges_after_prune: u64, pub gpu_launches: u64, pub h2d_bytes: u64,
    pub d2h_bytes: u64, pub alg_bytes_tier1: u64, pub alg_bytes_tier2: u64, pub alg_bytes_pack: u64,
    pub ms_h2d: f64, pub ms_device: f64, pub ms_d2h: f64, pub ms_host_post: f64, pub ms_total: f64,
    pub ms_kernel_tier1: f64, pub ms_kernel_tier2: f64, pub ms_kernel_pack: f64, pub ms_kernel_other: f64,
    pub alg_bytes_pack_t2: u64, pub ms_kernel_pack_t2: f64, pub mma_pairs: u64, pub mma_macs: u64,
}

#[repr(C)]
pub struct ejoin_result_t {
    pub n_edges: u64, pub edges: *mut ejoin_edge_t, pub class_of: *mut u32, pub n_info: u32, pub n_classes: u32,
    pub owner: *mut c_void, pub stats: ejoin_stats_t,
}

#[repr(C)]
pub struct ejoin_log_names_t {
    pub index: i32, pub id1: *const c_char, pub id2: *const c_char, pub vsegs1: *const c_char, pub vsegs2: *const c_char,
    pub dataset1: *const c_char, pub dataset2: *const c_char, pub cdr3_aa1: *const c_char, pub cdr3_aa2: *const c_char,
    pub bcs1: *const c_char, pub bcs2: *const c_char,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct ejoin_raw_edge_t { pub k1: u32, pub k2: u32, pub cd: u16, pub d: u16, pub flags: u32 }

#[link(name = "enclone_join")]
extern "C" {
    pub fn ejoin_params_default(p: *mut ejoin_params_t);
    pub fn ejoin_create(h: *mut *mut c_void, p: *const ejoin_params_t, device: c_int) -> c_int;
    pub fn ejoin_destroy(h: *mut c_void);
    pub fn ejoin_last_error(h: *const c_void) -> *const c_char;
    pub fn ejoin_run(h: *mut c_void, pack: *const ejoin_pack_t, out: *mut ejoin_result_t) -> c_int;
    pub fn ejoin_result_free(r: *mut ejoin_result_t);
    pub fn ejoin_plan_shards(pack: *const ejoin_pack_t, world: c_int, owner: *mut u32, bstart: *mut u32, cap: u32, nb: *mut u32) -> c_int;
    pub fn ejoin_run_shard(h: *mut c_void, pack: *const ejoin_pack_t, rank: c_int, world: c_int, edges: *mut *mut ejoin_raw_edge_t,
                           n_edges: *mut u64, stats: *mut ejoin_stats_t) -> c_int;
    pub fn ejoin_run_shard_final(h: *mut c_void, pack: *const ejoin_pack_t, rank: c_int, world: c_int, mode: *mut c_int,
                                 fin: *mut ejoin_result_t, raw: *mut *mut ejoin_raw_edge_t, n_raw: *mut u64, stats: *mut ejoin_stats_t) -> c_int;
    pub fn ejoin_raw_free(e: *mut ejoin_raw_edge_t);
    pub fn ejoin_finish(h: *mut c_void, pack: *const ejoin_pack_t, edges: *const ejoin_raw_edge_t, n: u64, out: *mut ejoin_result_t) -> c_int;
    pub fn ejoin_partition(pack: *const ejoin_pack_t, k1: *const u32, k2: *const u32, n: u64, class_of: *mut u32, n_classes: *mut u32) -> c_int;
    pub fn ejoin_partition_gpu(h: *mut c_void, pack: *const ejoin_pack_t, k1: *const u32, k2: *const u32, n: u64, class_of: *mut u32,
                               n_classes: *mut u32) -> c_int;
    pub fn ejoin_upload(h: *mut c_void, pack: *const ejoin_pack_t) -> c_int;
    pub fn ejoin_upload_shard(h: *mut c_void, pack: *const ejoin_pack_t, rank: c_int, world: c_int) -> c_int;
    pub fn ejoin_run_resident(h: *mut c_void, stats: *mut ejoin_stats_t) -> c_int;
    pub fn ejoin_host_alloc(bytes: usize) -> *mut c_void;
    pub fn ejoin_host_free(p: *mut c_void);
    pub fn ejoin_host_register(p: *mut c_void, bytes: usize) -> c_int;
    pub fn ejoin_host_unregister(p: *mut c_void) -> c_int;
    pub fn ejoin_pack_compact(input: *const ejoin_pack_t, threads: c_int, out: *mut *mut ejoin_pack_t) -> c_int;
    pub fn ejoin_pack_compact_free(p: *mut ejoin_pack_t);
    pub fn ejoin_comm_id(id: *mut u8) -> c_int;
    pub fn ejoin_comm_init(h: *mut c_void, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn ejoin_run_sharded(h: *mut c_void, pack: *const ejoin_pack_t, out: *mut ejoin_result_t) -> c_int;
    pub fn ejoin_format_join_log(pack: *const ejoin_pack_t, params: *const ejoin_params_t, edge: *const ejoin_edge_t, names: *const ejoin_log_names_t,
                                 buf: *mut c_char, cap: usize, len: *mut usize) -> c_int;
    pub fn ejoin_set_stage_timing(h: *mut c_void, on: c_int) -> c_int;
    pub fn ejoin_stage_time(h: *const c_void, index: c_int, name: *mut *const c_char, ms: *mut f64) -> c_int;
    pub fn ejoin_p1(k: u32, d: u32, n: u32) -> f64;
}

// ---- include/enclone_group.h: the symmetric grouping (enclone_tail/src/grouper.rs, Case 1) ----
#[repr(C)]
pub struct egroup_pack_t {
    pub abi_version: u32, pub n_clono: u32, pub n_exact: u32, pub n_chain: u32,
    pub clono_exact_off: *const u32, pub clono_exact: *const u32, pub exact_chain_off: *const u32,
    pub left: *const u8, pub seq_off: *const u64, pub seq_len: *const u32, pub seq_del_len: *const u32,
    pub cdr3_start: *const u32, pub cdr3_aa_off: *const u64, pub cdr3_aa_len: *const u32,
    pub seq_arena: *const u8, pub seq_arena_bytes: u64, pub aa_arena: *const u8, pub aa_arena_bytes: u64,
    pub col_off: *const u32, pub col_vname: *const u32, pub col_jname: *const u32, pub col_dname: *const u32, pub col_left: *const u8,
}
#[repr(C)]
pub struct egroup_opts_t {
    pub vj_refname: i32, pub vdj_refname: i32, pub v_heavy_refname: i32, pub vj_heavy_refname: i32, pub vdj_heavy_refname: i32,
    pub vj_len: i32, pub cdr3_len: i32, pub cdr3_heavy_len: i32, pub cdr3_light_len: i32, pub reserved: [i32; 3],
    pub heavy_pc: f64, pub light_pc: f64, pub cdr3_heavy_pc_hf: f64, pub hf_matrix: *const f64,
    pub aa_heavy_pc: f64, pub aa_light_pc: f64, pub cdr3_heavy_pc: f64, pub cdr3_light_pc: f64, pub cdr3_aa_heavy_pc: f64, pub cdr3_aa_light_pc: f64,
}
#[repr(C)]
pub struct egroup_stats_t {
    pub pairs_candidate: u64, pub pairs_evaluated: u64, pub chain_compares: u64, pub gpu_launches: u64, pub n_groups: u64,
    pub ms_device: f64, pub ms_total: f64,
}
#[link(name = "enclone_join")]
extern "C" {
    pub fn egroup_opts_default(o: *mut egroup_opts_t);
    pub fn egroup_run(pack: *const egroup_pack_t, opts: *const egroup_opts_t, device: c_int, group_of: *mut u32, stats: *mut egroup_stats_t) -> c_int;
    pub fn egroup_last_error() -> *const c_char;
    pub fn egroup_run_asym(pack: *const egroup_pack_t, in_center: *const u8, opts: *const egroup_asym_opts_t, device: c_int,
                           out: *mut egroup_asym_result_t, stats: *mut egroup_stats_t) -> c_int;
    pub fn egroup_asym_free(r: *mut egroup_asym_result_t);
}
#[repr(C)]
pub struct egroup_asym_opts_t { pub bound_kind: i32, pub top: u32, pub max_dist: f64, pub min_group: u32, pub reserved: u32 }
#[repr(C)]
pub struct egroup_asym_result_t { pub n_groups: u64, pub group_off: *mut u64, pub member: *mut u32, pub dist: *mut f64, pub owner: *mut c_void }
// ---- include/enclone_refine.h: the post-join refinement (define_mat, doublet / signature / weak-chain filters) ----
#[repr(C)]
pub struct erefine_pack_t {
    pub abi_version: u32, pub n_info: u32, pub n_exact: u32, pub n_chain: u32,
    pub exact_id: *const u32, pub exact_cols: [*const u8; 2], pub class_of: *const u32,
    pub n_joins: u64, pub join_k1: *const u32, pub join_k2: *const u32,
    pub chain_off: *const u32, pub ncells: *const u32, pub origin_rank: *const u32,
    pub seq_id: *const u32, pub cdr3_id: *const u32, pub order_rank: *const u32, pub left: *const u8,
    pub seq_len: *const u32, pub seq_off: *const u64, pub seq_arena: *const u8, pub seq_arena_bytes: u64,
}
#[repr(C)]
pub struct erefine_opts_t {
    pub doublet: i32, pub signature: i32, pub weak_chains: i32, pub doublet_mult: i32, pub signature_mult: i32,
    pub weak_max_cells: i32, pub weak_mult: i32, pub third_chain_diffs: i32, pub reserved: [i32; 4],
}
#[repr(C)]
pub struct erefine_stats_t {
    pub n_clonotypes_in: u64, pub n_clonotypes_out: u64, pub n_pure: u64, pub n_doublet: u64, pub n_signature: u64, pub n_weak: u64,
    pub triples_tested: u64, pub gpu_launches: u64, pub h2d_bytes: u64, pub ms_device: f64, pub ms_total: f64,
}
#[repr(C)]
pub struct erefine_result_t { pub fate: *mut u8, pub clonotype_of: *mut u32, pub column_of: *mut u32, pub n_columns: *mut u32 }
#[link(name = "enclone_join")]
extern "C" {
    pub fn erefine_opts_default(o: *mut erefine_opts_t);
    pub fn erefine_run(pack: *const erefine_pack_t, opts: *const erefine_opts_t, device: c_int, out: *mut erefine_result_t, stats: *mut erefine_stats_t) -> c_int;
    pub fn erefine_last_error() -> *const c_char;
}
// The safe wrapper `join_exacts(..) -> EquivRel` with `Flat::pack` is sketched in INTEGRATION.md §1-§3;
// enclone_b200/host/join_exacts.hpp is the same logic in C++ and is compiled and tested in this repository.
