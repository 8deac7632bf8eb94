// ejoin.cu — host orchestration and C ABI (include/enclone_join.h) of the B200 clonotype join.
//
// Replaces, behind the reference's own seam, the call
//   enclone::join::join_exacts(..) -> EquivRel  (+ raw_joins, join_info)
// reached in-tree at enclone_main/src/main_enclone.rs:36.
// There is NO CPU fallback in this file: without a CUDA device every entry point fails
// with EJOIN_E_NOGPU.  The host does only what the reference's own host-side epilogue does
// (ordered replay of the accepted edges = A.7/A.8, finish_join) plus table preparation.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>          // types and enums only: the library is bound at run time (dlopen), never linked

#include "ejoin_kernels.cuh"

#define P1_CACHE_MAX 512          // p1 tables kept across inputs (1.3 MB each: 0.67 GB)

namespace {

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// bump allocator over one DevBuf
struct Arena {
    DevBuf buf;
    size_t off = 0;
    void reset() { off = 0; }
    template <typename T> T *take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T *r = reinterpret_cast<T *>(static_cast<char *>(buf.p) + off);
        off += count * sizeof(T);
        return r;
    }
};
struct Sizer {                 // same walk as Arena, to size it first
    size_t off = 0;
    void take(size_t bytes) { off = (off + 255) & ~(size_t)255; off += bytes; }
};

}  // namespace

// NCCL, bound at run time: the copy already loaded in the process (torch's bundled one) or the system libnccl.so.2.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommAbort)(ncclComm_t);          // optional
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi *nccl_api() {
    static NcclApi api;
    static int state = 0;              // 0 = not tried, 1 = ok, -1 = unavailable
    if (state == 0) {
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        state = -1;
        if (lib) {
            bool ok = true;
            auto sym = [&](const char *name) { void *p = dlsym(lib, name); if (!p) ok = false; return p; };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.CommAbort = (decltype(api.CommAbort))dlsym(lib, "ncclCommAbort");
            api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
            api.Send = (decltype(api.Send))sym("ncclSend");
            api.Recv = (decltype(api.Recv))sym("ncclRecv");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            if (ok) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

struct ejoin_handle {
    int device = 0;
    ejoin_params_t params;
    std::string err;
    cudaStream_t st = nullptr, st_copy = nullptr;
    cudaEvent_t ev[8] = {};
    // EJOIN_DEBUG: per-kernel marks (events created on first use) printed after each pipeline run
    bool debug = false;
    std::vector<cudaEvent_t> mark_ev;
    std::vector<const char *> mark_name;
    std::vector<float> mark_ms;            // stage times of the last pipeline run (CUDA events on the launching stream)
    bool stage_timing = false;
    size_t n_marks = 0;
    // overlapped upload: small arrays first, then the tig arena in chunks, each with an event
    cudaEvent_t ev_copy0 = nullptr, ev_early = nullptr, ev_small = nullptr, ev_chunk[8] = {};
    int n_chunks = 0;
    uint64_t chunk_end[8] = {};     // arena byte offset (caller coordinates) up to which chunk i has landed
    bool copy_pending = false;
    bool tigs_in_place = false;    // zero-copy: din.tig_arena aliases the caller's pinned arena
    bool recs_pending = false;     // EJOIN_F_ENTRY_RECS: k_unpack_recs still has to run on this upload
    const ejoin_entry_rec_t *d_rec = nullptr;
    RecOut rec_out;
    int is_bcr = 1;
    int tables_is_bcr = -1;
    bool tables_built = false;       // prepare_tables has run on this handle: later joins do not wait for the key pass (device_pipeline)
    int sm_count = 148;
    int bor_grid = 0;          // cooperative grid of k_boruvka
    int mma_grid = 0;          // grid of k_sweep_mma (resident CTAs)
    // resident input + work buffers
    Arena in_arena, work;
    DevIn din;
    uint32_t n = 0;            // entries uploaded
    uint32_t k_base = 0;       // info index of uploaded entry 0
    RangeTab rt;               // the uploaded entry ranges: local index <-> info index
    uint64_t tig_bytes = 0;
    uint64_t h2d_bytes = 0;
    bool uploaded = false;
    uint32_t last_stride = 1, last_phase = 0;   // tile-row share of the last run_shard (stride mode), reused by run_resident
    bool device_final = false;  // the last pipeline run produced the final emitted edge list on the device
    // tables (cached across runs)
    DevBuf d_part;              // ejoin_partition_gpu scratch
    DevBuf d_counters, d_presL, d_presC, d_p1ptrs, d_multinfo, d_multpool, d_cdmax, d_p1scratch, d_Ls, d_tabptrs;
    std::vector<double *> p1_host_ptrs;            // [8192] device pointers (host copy)
    std::vector<MultInfo> multinfo_host;           // [65536]
    std::vector<double> multpool_host;
    std::vector<uint16_t> cdmax_host;
    std::vector<DevBuf> p1_tables;
    unsigned long long xrow_cap = 0;
    uint32_t xrow_seg = 0;
    std::vector<uint32_t> ref_plane_off;   // host copy of DevIn::ref_plane_off (kept alive for the async upload)
    unsigned long long cand_cap = 0;    // survivor list (tier-1 survivors beyond the phase-A slots): 16-byte mask records
    unsigned long long rest_cap = 0;    // survivor pairs kept for round 2
    unsigned long long edge_cap = 0;    // to-be-scored lists, accepted edges, sort arrays
    // work pointers of the last pipeline run
    EntryMeta *meta = nullptr; uint32_t *c3 = nullptr; uint4 *t2 = nullptr; uint32_t *k_to_s = nullptr;
    unsigned long long *kept_keys = nullptr, *kept_vals = nullptr;
    uint32_t *w_parent = nullptr, *w_comp_size = nullptr; unsigned long long *w_best = nullptr, *w_f0_key = nullptr, *w_f0_val = nullptr;   // forest scratch of the last run
    DevCounters counters;
    ejoin_stats_t stats;
    uint64_t launches = 0;
    // pinned staging for results
    void *pin = nullptr; size_t pin_cap = 0;
    void *pin2 = nullptr; size_t pin2_cap = 0;   // ejoin_partition_gpu staging (h->pin may hold a live result)
    // multi-GPU (ejoin_comm_init): the one exchange step of the path runs inside the library, device to device
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    DevBuf d_gather, d_xcounts, d_exact_all;
    // host replay scratch (kept across runs; only touched entries are reset)
    std::vector<uint32_t> uf_parent, cnt_tmp;
    std::vector<size_t> forest_tmp;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return EJOIN_E_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int fail(ejoin_handle *h, int code, const std::string &msg) { h->err = msg; return code; }

DevParams dev_params(const ejoin_params_t &p, int is_bcr) {
    DevParams d;
    d.max_score = p.max_score; d.join_cdr3_ident = p.join_cdr3_ident; d.fwr1_cdr12_delta = p.fwr1_cdr12_delta;
    d.auto_share = p.auto_share; d.comp_filt = p.comp_filt; d.comp_filt_bound = p.comp_filt_bound;
    d.max_cdr3_diffs = p.max_cdr3_diffs; d.cdr3_mult = p.cdr3_mult; d.max_degradation = p.max_degradation;
    d.max_diffs = p.max_diffs; d.max_diffs_tcr = p.max_diffs_tcr; d.ref_v_trim = p.ref_v_trim; d.ref_j_trim = p.ref_j_trim;
    d.old_light = p.old_light; d.mix_donors = p.mix_donors; d.is_bcr = is_bcr; d.easy = p.easy; d.degradation_mode = p.degradation_mode;
    return d;
}

int ensure_pin(ejoin_handle *h, size_t bytes) {
    if (bytes <= h->pin_cap) return 0;
    if (h->pin) cudaFreeHost(h->pin);
    h->pin = nullptr; h->pin_cap = 0;
    CK(cudaMallocHost(&h->pin, bytes + bytes / 2 + 4096));
    h->pin_cap = bytes + bytes / 2 + 4096;
    return 0;
}

// cd cap per total CDR3 length: the largest cd passing MAX_CDR3_DIFFS, the TCR rule and
// "100*(1 - cd/total) >= JOIN_CDR3_IDENT" evaluated with the oracle's exact expression.
void build_cdmax(ejoin_handle *h, int is_bcr) {
    h->cdmax_host.assign(C3_MAX_TOTAL + 1, 0xFFFF);
    for (int total = 0; total <= C3_MAX_TOTAL; total++) {
        int best = -1;
        for (int cd = 0; cd <= total; cd++) {
            if (!is_bcr && cd > 0) break;
            if (cd > h->params.max_cdr3_diffs) break;
            double ident = 100.0 * (1.0 - (double)cd / (double)(total));
            if (ident < h->params.join_cdr3_ident) break;
            best = cd;
        }
        h->cdmax_host[total] = best < 0 ? 0xFFFF : (uint16_t)best;
    }
}

// mult = MULT_POW ^ (sum_c CDR3_NORMAL_LEN * cd_c / n_c), with the host libm pow, in the
// oracle's floating-point order (help1.rs:300-306).
double mult_value(const ejoin_params_t &p, int cd0, int n0, int cd1, int n1) {
    if (p.old_mult) {
        // OLD_MULT (pages/cellranger.html.src:142-144): N = sum_{i <= cd} C(3 * cn, i), cn = total CDR3 amino acids; same
        // floating-point order as the oracle
        const int cn = (n0 + n1) / 3, cd = cd0 + cd1;
        double s = 0.0, t = 1.0;
        for (int i = 0; i <= cd; i++) { s += t; t = t * (double)(3 * cn - i) / (double)(i + 1); }
        return s;
    }
    double cdx = 0.0;
    cdx += (double)p.cdr3_normal_len * (double)cd0 / (double)n0;
    cdx += (double)p.cdr3_normal_len * (double)cd1 / (double)n1;
    return pow(p.mult_pow, cdx);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// p1 on the host: the same double-double recurrence as k_p1_tables (dd.cuh).
// ---------------------------------------------------------------------------------------------
extern "C" double ejoin_p1(uint32_t k, uint32_t d, uint32_t n) {
    if (d == 0) return 1.0;
    if (d > k) d = k;
    const double nn = (double)n;
    std::vector<dd_t> F(k + 2);
    F[0] = dd_from(1.0);
    for (uint32_t u = 0; u <= k; u++) F[u + 1] = p1_f_step(F[u], (int)u, nn);
    std::vector<dd_t> prev(d + 1), cur(d + 1);
    for (uint32_t j = 0; j <= d; j++) prev[j] = dd_from(j == 0 ? 1.0 : 0.0);
    for (uint32_t kk = 1; kk <= k; kk++) {
        for (uint32_t j = 0; j < d; j++) cur[j] = p1_b_step((int)kk, (int)j, prev[j > 0 ? j - 1 : 0], prev[j], nn);
        std::swap(prev, cur);
    }
    dd_t p = dd_from(1.0);
    for (uint32_t j = 0; j < d; j++) {
        dd_t T = j <= k ? dd_mul(prev[j], F[k - j]) : dd_from(0.0);
        p = dd_sub(p, T);
    }
    double v = dd_to_double(p);
    return v < 0.0 ? 0.0 : v;
}

extern "C" void ejoin_params_default(ejoin_params_t *p) {
    memset(p, 0, sizeof(*p));
    p->max_score = 100000.0; p->mult_pow = 80.0; p->join_cdr3_ident = 85.0; p->fwr1_cdr12_delta = 20.0;
    p->cdr3_normal_len = 42; p->auto_share = 15; p->comp_filt = 8; p->comp_filt_bound = 80; p->max_cdr3_diffs = 1000;
    p->cdr3_mult = 5; p->max_degradation = 2; p->max_diffs = 1000000; p->max_diffs_tcr = 5; p->ref_v_trim = 15; p->ref_j_trim = 15;
}

extern "C" int ejoin_create(ejoin_handle_t **out, const ejoin_params_t *p, int device) {
    if (!out || !p) return EJOIN_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) { cudaGetLastError(); return EJOIN_E_NOGPU; }
    ejoin_handle *h = new ejoin_handle();
    h->device = device; h->params = *p;
    if (const char *cc = getenv("EJOIN_CAND_CAP")) h->cand_cap = h->rest_cap = h->edge_cap = strtoull(cc, nullptr, 10);   // tests: force the overflow-and-retry path
    *out = h;                                    // so the caller can read last_error on failure
    if (p->auto_share > P1_DCAP) return fail(h, EJOIN_E_UNSUPPORTED, "AUTO_SHARES above the p1 table depth");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
    for (auto &e : h->ev) CK(cudaEventCreate(&e));
    CK(cudaEventCreate(&h->ev_copy0)); CK(cudaEventCreate(&h->ev_small)); CK(cudaEventCreate(&h->ev_early));
    for (auto &e : h->ev_chunk) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventDestroy(h->ev_chunk[7])); CK(cudaEventCreate(&h->ev_chunk[7]));
    CK(h->d_counters.ensure(sizeof(DevCounters)));
    CK(h->d_presL.ensure(8192 / 8));
    CK(h->d_presC.ensure(65536 / 8));
    CK(h->d_p1ptrs.ensure(8192 * sizeof(double *)));
    CK(h->d_multinfo.ensure(65536 * sizeof(MultInfo)));
    CK(h->d_cdmax.ensure((C3_MAX_TOTAL + 1) * sizeof(uint16_t)));
    h->p1_host_ptrs.assign(8192, nullptr);
    MultInfo none; none.off = 0xFFFFFFFFu; none.d1p1 = 0;
    h->multinfo_host.assign(65536, none);
    CK(cudaMemsetAsync(h->d_p1ptrs.p, 0, 8192 * sizeof(double *), h->st));
    CK(cudaStreamSynchronize(h->st));
    return EJOIN_OK;
}

extern "C" void ejoin_destroy(ejoin_handle_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    // abort, not destroy: tearing a handle down must never wait for the other ranks (they may already be gone)
    if (h->comm && nccl_api()) { NcclApi *nc = nccl_api(); if (nc->CommAbort) nc->CommAbort(h->comm); else nc->CommDestroy(h->comm); h->comm = nullptr; }
    h->in_arena.buf.release(); h->work.buf.release(); h->d_part.release(); h->d_gather.release(); h->d_xcounts.release(); h->d_exact_all.release();
    for (DevBuf *b : { &h->d_counters, &h->d_presL, &h->d_presC, &h->d_p1ptrs, &h->d_multinfo, &h->d_multpool, &h->d_cdmax, &h->d_p1scratch,
                       &h->d_Ls, &h->d_tabptrs })
        b->release();
    for (auto &t : h->p1_tables) t.release();
    if (h->pin) cudaFreeHost(h->pin);
    if (h->pin2) cudaFreeHost(h->pin2);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (auto &e : h->mark_ev) if (e) cudaEventDestroy(e);
    for (auto &e : h->ev_chunk) if (e) cudaEventDestroy(e);
    if (h->ev_copy0) cudaEventDestroy(h->ev_copy0);
    if (h->ev_small) cudaEventDestroy(h->ev_small);
    if (h->ev_early) cudaEventDestroy(h->ev_early);
    if (h->st_copy) { cudaStreamSynchronize(h->st_copy); cudaStreamDestroy(h->st_copy); }
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

extern "C" const char *ejoin_last_error(const ejoin_handle_t *h) { return h ? h->err.c_str() : "null handle"; }

extern "C" void *ejoin_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void ejoin_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int ejoin_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return EJOIN_E_ARG;
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return EJOIN_E_CUDA; }
    return EJOIN_OK;
}
extern "C" int ejoin_host_unregister(void *p) {
    if (!p) return EJOIN_E_ARG;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return EJOIN_E_CUDA; }
    return EJOIN_OK;
}

namespace {

// every array the library dereferences, given the counts (the header promises EJOIN_E_ARG, not a crash)
const char *pack_null_field(const ejoin_pack_t *pk) {
    if (pk->n_info) {
        const bool recs = pk->flags & EJOIN_F_ENTRY_RECS;
        if (!pk->exact_id) return "exact_id";
        if (recs && !pk->entry_rec) return "entry_rec";
        if ((pk->flags & EJOIN_F_TIGS_2BIT) && !pk->tig2_arena && pk->tig2_arena_words) return "tig2_arena";
        if (!(pk->flags & EJOIN_F_TIGS_2BIT) && !pk->tig_arena) return "tig_arena";
        if (pk->tig_arena_bytes && !pk->tig_arena) return "tig_arena";
        const bool c2 = pk->flags & EJOIN_F_CDR3_2BIT;
        if (c2 && (!pk->cdr3_2bit_off || (!pk->cdr3_2bit && pk->cdr3_2bit_words))) return "cdr3_2bit / cdr3_2bit_off";
        if (!c2 && !pk->cdr3_arena) return "cdr3_arena";
        for (int c = 0; c < 2; c++) {
            if (!pk->len[c]) return "len";
            if (!pk->cdr3_len[c]) return "cdr3_len";
            if (!c2 && !pk->cdr3_off[c]) return "cdr3_off";
        }
        if (!recs) {
            if (!pk->c_ref_id_light) return "c_ref_id_light";
            if (!pk->hcomp) return "hcomp";
            if (!pk->fr1_start || !pk->cdr1_start || !pk->fr2_start || !pk->cdr2_start || !pk->fr3_start) return "region starts";
            for (int c = 0; c < 2; c++) {
                if (!pk->tig_off[c]) return "tig_off";
                if (!pk->has_del[c]) return "has_del";
                if (!pk->cdr3_start[c]) return "cdr3_start";
                if (!pk->v_ref_id[c] || !pk->j_ref_id[c] || !pk->dref_id[c]) return "v_ref_id/j_ref_id/dref_id";
                if (!pk->vs_seq[c] || !pk->js_seq[c] || !pk->vuni_seq[c] || !pk->utr_seq[c]) return "vs_seq/js_seq/vuni_seq/utr_seq";
                if (!pk->vname_id[c]) return "vname_id";
            }
        }
        if (!pk->n_exact) return "n_exact is 0 with n_info > 0";
    }
    if (pk->n_exact && (!pk->nchains || !pk->ncells || !pk->donor_lo || !pk->donor_hi)) return "nchains/ncells/donor_lo/donor_hi";
    if (pk->n_refseq && (!pk->refseq_off || !pk->refseq_len || !pk->refseq_arena)) return "refseq_off/refseq_len/refseq_arena";
    return nullptr;
}

int check_pack(ejoin_handle *h, const ejoin_pack_t *pk) {
    if (!pk) return fail(h, EJOIN_E_ARG, "null pack");
    if (pk->abi_version != EJOIN_ABI_VERSION) return fail(h, EJOIN_E_ARG, "abi_version mismatch");
    const char *bad = pack_null_field(pk);
    if (bad) return fail(h, EJOIN_E_ARG, std::string("null array in pack: ") + bad);
    return 0;
}

// H2D of this rank's entries: one or several ascending entry ranges [lo,hi) (whole arrays for the per-exact and germline
// tables).  Per-entry arrays are concatenated (local index = position in that concatenation, h->rt maps it back to the info
// index); the arenas keep the caller's offsets: a device buffer spans from the first to the last referenced byte and only
// the owned slices are copied into it.
typedef std::vector<std::pair<uint32_t, uint32_t>> Ranges;
int upload_ranges(ejoin_handle *h, const ejoin_pack_t *pk, const Ranges &rg_in, bool allow_in_place = true) {
    int rc = check_pack(h, pk);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    Ranges rg;
    for (auto &r : rg_in) if (r.second > r.first) rg.push_back(r);
    if (rg.size() > EJOIN_MAX_RANGES) return fail(h, EJOIN_E_ARG, "too many entry ranges");
    uint32_t n = 0;
    RangeTab &rt = h->rt;
    memset(&rt, 0, sizeof(rt));
    for (auto &r : rg) { rt.lo[rt.n] = r.first; rt.base[rt.n] = n; rt.n++; n += r.second - r.first; }
    const bool everything = rg.size() == 1 && rg[0].first == 0 && rg[0].second == pk->n_info;
    const bool two_bit = pk->flags & EJOIN_F_TIGS_2BIT, recs = pk->flags & EJOIN_F_ENTRY_RECS, c2 = pk->flags & EJOIN_F_CDR3_2BIT;
    auto tig_off_of = [&](int c, uint32_t k) -> uint64_t { return recs ? (uint64_t)pk->entry_rec[k].tig_off[c] : pk->tig_off[c][k]; };
    auto has_del_of = [&](int c, uint32_t k) -> bool { return recs ? pk->entry_rec[k].has_del[c] != 0 : pk->has_del[c][k] != 0; };
    // arena slices referenced by each range: ASCII bytes [tlo,thi), 2-bit words [wlo,whi), CDR3 bytes or words [clo,chi);
    // the unsuffixed variables are the spans over all ranges.  In 2-bit mode the ASCII arena only holds the few chains
    // with a deletion marker: it goes up whole.
    struct Slice { uint64_t tlo, thi, clo, chi, wlo, whi; };
    std::vector<Slice> sl(rg.size());
    uint64_t tlo = 0, thi = pk->tig_arena_bytes, clo = 0, chi = c2 ? pk->cdr3_2bit_words : pk->cdr3_arena_bytes, wlo = 0, whi = two_bit ? pk->tig2_arena_words : 0;
    if (everything) sl[0] = Slice{ tlo, thi, clo, chi, wlo, whi };
    else if (n == 0) { clo = chi = wlo = whi = 0; if (!two_bit) tlo = thi = 0; }
    else {
        // A range's slices are read off its first and last entries: the arenas must be laid out in entry order (offsets
        // non-decreasing in k, as every packer writes them; include/enclone_join.h).  The kernels check each entry against
        // the span (ERR_ARENA_RANGE -> EJOIN_E_ARG) instead of the host walking 10^6 entries per call.
        for (size_t ri = 0; ri < rg.size(); ri++) {
            Slice x{ ~0ull, 0, ~0ull, 0, ~0ull, 0 };
            auto visit = [&](uint32_t k) {
                for (int c = 0; c < 2; c++) {
                    uint64_t o = tig_off_of(c, k), e;
                    if (two_bit && !has_del_of(c, k)) { e = o + (pk->len[c][k] + 31) / 32; x.wlo = std::min(x.wlo, o); x.whi = std::max(x.whi, e); }
                    else if (!two_bit) { e = o + pk->len[c][k]; x.tlo = std::min(x.tlo, o); x.thi = std::max(x.thi, e); }
                    if (!c2) { o = pk->cdr3_off[c][k]; e = o + pk->cdr3_len[c][k]; x.clo = std::min(x.clo, o); x.chi = std::max(x.chi, e); }
                }
                if (c2) {
                    const uint64_t o = pk->cdr3_2bit_off[k], e = o + ((uint64_t)pk->cdr3_len[0][k] + pk->cdr3_len[1][k] + 31) / 32;
                    x.clo = std::min(x.clo, o); x.chi = std::max(x.chi, e);
                }
            };
            auto plain2 = [&](uint32_t k) { return two_bit && ((pk->len[0][k] && !has_del_of(0, k)) || (pk->len[1][k] && !has_del_of(1, k))); };
            uint32_t a = rg[ri].first, b = rg[ri].second - 1;
            visit(a); visit(b);
            if (two_bit) {          // the first / last entry that has a 2-bit chain (nearly always the end entries themselves)
                while (a < b && !plain2(a)) visit(++a);
                while (b > a && !plain2(b)) visit(--b);
            }
            if (x.chi < x.clo) x.clo = x.chi = 0;
            if (x.whi <= x.wlo) x.wlo = x.whi = 0;
            if (x.thi < x.tlo) x.tlo = x.thi = 0;
            if (two_bit) { x.tlo = 0; x.thi = pk->tig_arena_bytes; }
            sl[ri] = x;
        }
        auto span = [&](uint64_t Slice::*lo_m, uint64_t Slice::*hi_m, uint64_t &lo_o, uint64_t &hi_o) {
            lo_o = ~0ull; hi_o = 0;
            for (auto &x : sl) if (x.*hi_m > x.*lo_m) { lo_o = std::min(lo_o, x.*lo_m); hi_o = std::max(hi_o, x.*hi_m); }
            if (hi_o <= lo_o) lo_o = hi_o = 0;
        };
        span(&Slice::clo, &Slice::chi, clo, chi); span(&Slice::wlo, &Slice::whi, wlo, whi);
        if (!two_bit) span(&Slice::tlo, &Slice::thi, tlo, thi);
    }
    // zero-copy: the caller's arena is pinned, device-readable and padded (EJOIN_F_TIGS_IN_PLACE): k_pack_t2 (and, for
    // ASCII, the bytewise path through the mirror) reads it in place, and only the needed entries ever cross PCIe
    bool in_place = false;
    const uint8_t *in_place_ptr = nullptr;
    {
        const uint8_t *a_lo = two_bit ? (const uint8_t *)(pk->tig2_arena + wlo) : pk->tig_arena + tlo;
        const uint8_t *a_hi = two_bit ? (const uint8_t *)(pk->tig2_arena + whi) - 1 : pk->tig_arena + thi - 1;
        if (allow_in_place && (pk->flags & EJOIN_F_TIGS_IN_PLACE) && a_hi > a_lo && !getenv("EJOIN_NO_ZERO_COPY")) {
            cudaPointerAttributes a0, a1;
            if (cudaPointerGetAttributes(&a0, a_lo) == cudaSuccess && cudaPointerGetAttributes(&a1, a_hi) == cudaSuccess &&
                a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost && a0.devicePointer) {
                in_place_ptr = (const uint8_t *)a0.devicePointer;     // device alias of the first referenced byte / word
                in_place = true;
            } else cudaGetLastError();
        }
    }
    Sizer sz;
    const size_t n1 = n ? n : 1;
    sz.take(4 * n1);                                   // exact_id
    for (int c = 0; c < 2; c++) { sz.take(2 * n1); sz.take(8 * n1); sz.take(n1); sz.take(2 * n1); sz.take(8 * n1); sz.take(2 * n1); for (int i = 0; i < 8; i++) sz.take(4 * n1); }
    sz.take(4 * n1); for (int i = 0; i < 6; i++) sz.take(2 * n1);
    for (int i = 0; i < 4; i++) sz.take(4 * (size_t)std::max(1u, pk->n_exact));
    sz.take(8 * (size_t)std::max(1u, pk->n_refseq)); sz.take(4 * (size_t)std::max(1u, pk->n_refseq));
    sz.take((c2 ? 8 : 1) * (chi - clo) + 64); sz.take(pk->refseq_arena_bytes + 64); sz.take(thi - tlo + 64 + 512);
    sz.take(8 * (whi - wlo) + 256); sz.take(sizeof(ejoin_entry_rec_t) * n1);
    sz.take(4 * (size_t)std::max(1u, pk->n_refseq)); sz.take(12 * ((size_t)pk->refseq_arena_bytes / 32 + 3 * (size_t)pk->n_refseq + 8));   // germline planes
    CK(h->in_arena.buf.ensure(sz.off + 4096));
    h->in_arena.reset();
    Arena &A = h->in_arena;
    uint64_t bytes = 0;
    DevIn &d = h->din;
    memset(&d, 0, sizeof(d));
    d.n_info = n; d.n_exact = pk->n_exact; d.n_refseq = pk->n_refseq;
    auto up = [&](auto *&dst, const auto *src, size_t count) -> cudaError_t {
        using T = std::remove_cv_t<std::remove_pointer_t<std::remove_reference_t<decltype(dst)>>>;
        T *p = A.take<T>(count ? count : 1);
        dst = p;
        if (!count) return cudaSuccess;
        bytes += count * sizeof(T);
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, h->st_copy);
    };
    // a per-entry array: the owned ranges, concatenated
    auto up_e = [&](auto *&dst, const auto *src_all) -> cudaError_t {
        using T = std::remove_cv_t<std::remove_pointer_t<std::remove_reference_t<decltype(dst)>>>;
        T *p = A.take<T>(n ? n : 1);
        dst = p;
        for (size_t ri = 0; ri < rg.size(); ri++) {
            const size_t cnt = rg[ri].second - rg[ri].first;
            bytes += cnt * sizeof(T);
            cudaError_t e = cudaMemcpyAsync(p + rt.base[ri], src_all + rg[ri].first, cnt * sizeof(T), cudaMemcpyHostToDevice, h->st_copy);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    // an arena: a buffer over the span [lo,hi) (in elements of T), the owned slices copied to their own offsets
    auto up_arena = [&](auto *&dst, const auto *src_all, uint64_t lo_, uint64_t hi_, uint64_t Slice::*lo_m, uint64_t Slice::*hi_m, size_t slack) -> cudaError_t {
        using T = std::remove_cv_t<std::remove_pointer_t<std::remove_reference_t<decltype(dst)>>>;
        T *p = A.take<T>(hi_ - lo_ + slack);
        dst = p;
        uint64_t done_hi = 0;
        for (auto &x : sl) {
            uint64_t a_ = std::max(x.*lo_m, done_hi), b_ = x.*hi_m;          // slices ascend; never copy an overlap twice
            if (b_ <= a_) continue;
            bytes += (b_ - a_) * sizeof(T);
            cudaError_t e = cudaMemcpyAsync(p + (a_ - lo_), src_all + a_, (b_ - a_) * sizeof(T), cudaMemcpyHostToDevice, h->st_copy);
            if (e != cudaSuccess) return e;
            done_hi = b_;
        }
        return cudaSuccess;
    };
    // the copy stream must not run ahead of the previous run's kernels that still read the old input
    CK(cudaStreamSynchronize(h->st));
    CK(cudaEventRecord(h->ev_copy0, h->st_copy));
    // early set: what the keys, the sort, the c3 rows and the sweep read
    CK(up_e(d.exact_id, pk->exact_id));
    CK(up(d.nchains, pk->nchains, pk->n_exact));
    for (int c = 0; c < 2; c++) {
        CK(up_e(d.len[c], pk->len[c]));
        CK(up_e(d.cdr3_len[c], pk->cdr3_len[c]));
        if (!c2) CK(up_e(d.cdr3_off[c], pk->cdr3_off[c]));
    }
    if (c2) { CK(up_e(d.cdr3_2bit_off, pk->cdr3_2bit_off)); CK(up_arena(d.cdr3_2bit, pk->cdr3_2bit, clo, chi, &Slice::clo, &Slice::chi, 8)); }
    else CK(up_arena(d.cdr3_arena, pk->cdr3_arena, clo, chi, &Slice::clo, &Slice::chi, 64));
    CK(cudaEventRecord(h->ev_early, h->st_copy));
    // late set: what k_meta and the scoring read; goes up while the sweep runs
    h->recs_pending = false;
    if (recs) {
        // one 64-byte record per entry crosses PCIe; k_unpack_recs (first thing after the wait for the late set) widens it
        // into the parallel arrays the kernels read
        const ejoin_entry_rec_t *d_rec = nullptr;
        CK(up_e(d_rec, pk->entry_rec));
        RecOut &o = h->rec_out;
        auto carve = [&](auto *&dev_const, auto *&dev_mut) {
            using T = std::remove_pointer_t<std::remove_reference_t<decltype(dev_mut)>>;
            dev_mut = A.take<T>(n1); dev_const = dev_mut;
        };
        for (int c = 0; c < 2; c++) {
            carve(d.tig_off[c], o.tig_off[c]); carve(d.has_del[c], o.has_del[c]); carve(d.cdr3_start[c], o.cdr3_start[c]);
            carve(d.v_ref_id[c], o.v_ref_id[c]); carve(d.j_ref_id[c], o.j_ref_id[c]); carve(d.dref_id[c], o.dref_id[c]);
            carve(d.vs_seq[c], o.vs_seq[c]); carve(d.js_seq[c], o.js_seq[c]); carve(d.vname_id[c], o.vname_id[c]);
            carve(d.vuni_seq[c], o.vuni_seq[c]); carve(d.utr_seq[c], o.utr_seq[c]);
        }
        carve(d.c_ref_id_light, o.c_ref_id_light); carve(d.hcomp, o.hcomp);
        carve(d.fr1, o.fr1); carve(d.cdr1, o.cdr1); carve(d.fr2, o.fr2); carve(d.cdr2, o.cdr2); carve(d.fr3, o.fr3);
        h->d_rec = d_rec; h->recs_pending = true;
    } else {
        for (int c = 0; c < 2; c++) {
            CK(up_e(d.tig_off[c], pk->tig_off[c]));
            CK(up_e(d.has_del[c], pk->has_del[c]));
            CK(up_e(d.cdr3_start[c], pk->cdr3_start[c]));
            CK(up_e(d.v_ref_id[c], pk->v_ref_id[c]));
            CK(up_e(d.j_ref_id[c], pk->j_ref_id[c]));
            CK(up_e(d.dref_id[c], pk->dref_id[c]));
            CK(up_e(d.vs_seq[c], pk->vs_seq[c]));
            CK(up_e(d.js_seq[c], pk->js_seq[c]));
            CK(up_e(d.vname_id[c], pk->vname_id[c]));
            CK(up_e(d.vuni_seq[c], pk->vuni_seq[c]));
            CK(up_e(d.utr_seq[c], pk->utr_seq[c]));
        }
        CK(up_e(d.c_ref_id_light, pk->c_ref_id_light));
        CK(up_e(d.hcomp, pk->hcomp));
        CK(up_e(d.fr1, pk->fr1_start)); CK(up_e(d.cdr1, pk->cdr1_start)); CK(up_e(d.fr2, pk->fr2_start));
        CK(up_e(d.cdr2, pk->cdr2_start)); CK(up_e(d.fr3, pk->fr3_start));
    }
    CK(up(d.ncells, pk->ncells, pk->n_exact));
    CK(up(d.donor_lo, pk->donor_lo, pk->n_exact)); CK(up(d.donor_hi, pk->donor_hi, pk->n_exact));
    CK(up(d.refseq_off, pk->refseq_off, pk->n_refseq)); CK(up(d.refseq_len, pk->refseq_len, pk->n_refseq));
    CK(up(d.refseq_arena, pk->refseq_arena, pk->refseq_arena_bytes));
    {
        // germline bit planes: offsets from the host's lengths; the planes themselves are built by k_pack_ref
        h->ref_plane_off.resize(pk->n_refseq ? pk->n_refseq : 1);
        uint32_t words = 0;
        for (uint32_t i = 0; i < pk->n_refseq; i++) { h->ref_plane_off[i] = words + 1; words += (pk->refseq_len[i] + 31) / 32 + 2; }
        d.ref_plane_words = words + 2;
        CK(up(d.ref_plane_off, h->ref_plane_off.data(), pk->n_refseq));
        d.ref_planes = A.take<uint32_t>(3 * (size_t)d.ref_plane_words);
        CK(cudaMemsetAsync(d.ref_planes, 0, 12 * (size_t)d.ref_plane_words, h->st_copy));
    }
    CK(cudaEventRecord(h->ev_small, h->st_copy));       // everything but the V..J bytes: keys / sort / groups / meta / c3 can start
    h->tigs_in_place = in_place;
    if (two_bit) {
        // ASCII chains (deletion marker / non-ACGT): the small arena goes up whole, with the late set
        uint8_t *dst = A.take<uint8_t>(thi - tlo + 64);
        if (thi > tlo) { CK(cudaMemcpyAsync(dst, pk->tig_arena + tlo, thi - tlo, cudaMemcpyHostToDevice, h->st_copy)); bytes += thi - tlo; }
        d.tig_arena = dst; d.tig_src = dst; d.tig_copy = nullptr;
        CK(cudaEventRecord(h->ev_small, h->st_copy));
        if (in_place) {
            d.tig2_src = (const uint64_t *)in_place_ptr;                 // k_pack_t2 pulls the needed entries' words over PCIe
            h->n_chunks = 0;
        } else {
            const uint64_t *w = nullptr;
            CK(up_arena(w, pk->tig2_arena, wlo, whi, &Slice::wlo, &Slice::whi, 8));
            d.tig2_src = w;
            h->n_chunks = 1; h->chunk_end[0] = ~0ull;
            CK(cudaEventRecord(h->ev_chunk[7], h->st_copy));
        }
        h->copy_pending = true;
    } else if (in_place) {
        // device mirror with the same 16-byte phase as the host arena (k_pack_t2 copies chunk for chunk)
        uint8_t *mir = A.take<uint8_t>(thi - tlo + 64 + 256);
        mir += ((uintptr_t)in_place_ptr & 15);
        d.tig_src = in_place_ptr; d.tig_copy = mir; d.tig_arena = mir;
        h->n_chunks = 0;
        h->copy_pending = true;                                   // still wait for ev_small
    } else {
        // the V..J arena (3/4 of the bytes) goes up in chunks; k_pack_t2 packs each chunk's entries as it lands
        const uint64_t tb = thi - tlo;
        uint8_t *dst = A.take<uint8_t>(tb + 64);
        d.tig_arena = dst; d.tig_src = dst; d.tig_copy = nullptr;
        const int nch = sl.size() > 1 ? 0 : (tb >= (32ull << 20) ? 8 : 1);
        h->n_chunks = nch ? nch : 1;
        uint64_t done = 0;
        if (nch == 0) {          // several slices: each goes to its own offset; k_pack_t2 runs once when all have landed
            uint64_t done_hi = 0;
            for (auto &x : sl) {
                const uint64_t a_ = std::max(x.tlo, done_hi), b_ = x.thi;
                if (b_ <= a_) continue;
                CK(cudaMemcpyAsync(dst + (a_ - tlo), pk->tig_arena + a_, b_ - a_, cudaMemcpyHostToDevice, h->st_copy));
                bytes += b_ - a_; done_hi = b_;
            }
            h->chunk_end[0] = ~0ull;
            CK(cudaEventRecord(h->ev_chunk[7], h->st_copy));
        }
        for (int i = 0; i < nch; i++) {
            uint64_t end = i == nch - 1 ? tb : ((tb * (uint64_t)(i + 1) / nch) & ~(uint64_t)4095);
            if (end > done) CK(cudaMemcpyAsync(dst + done, pk->tig_arena + tlo + done, end - done, cudaMemcpyHostToDevice, h->st_copy));
            done = end;
            h->chunk_end[i] = tlo + end;
            CK(cudaEventRecord(h->ev_chunk[i == nch - 1 ? 7 : i], h->st_copy));
        }
        if (nch) bytes += tb;
        h->copy_pending = true;
    }
    // the offsets index the original arenas: rebase the device pointers instead of the offsets
    d.tig_arena -= tlo; d.tig_src -= tlo; if (d.tig_copy) d.tig_copy -= tlo;
    if (d.tig2_src) d.tig2_src -= wlo;
    d.tig2_planes = (pk->flags & EJOIN_F_TIGS_PLANES) ? 1u : 0u;
    if (c2) d.cdr3_2bit -= clo; else d.cdr3_arena -= clo;
    d.tig_lo = tlo; d.tig_hi = thi; d.tig2_lo = wlo; d.tig2_hi = whi; d.cdr3_lo = clo; d.cdr3_hi = chi;
    h->n = n; h->k_base = rt.n ? rt.lo[0] : 0; h->tig_bytes = two_bit ? 32 * (whi - wlo) + (thi - tlo) : thi - tlo; h->h2d_bytes = bytes; h->uploaded = true;
    return 0;
}

int upload_range(ejoin_handle *h, const ejoin_pack_t *pk, uint32_t lo, uint32_t hi, bool allow_in_place = true) {
    return upload_ranges(h, pk, Ranges{ { lo, hi } }, allow_in_place);
}
// local index (position in the concatenated ranges) <-> info index
uint32_t host_to_global(const RangeTab &rt, uint32_t local) { uint32_t r = 0; while (r + 1 < rt.n && rt.base[r + 1] <= local) r++; return rt.lo[r] + (local - rt.base[r]); }
bool host_to_local(const RangeTab &rt, uint32_t n, uint32_t g, uint32_t *local) {
    for (uint32_t r = 0; r < rt.n; r++) {
        const uint32_t cnt = (r + 1 < rt.n ? rt.base[r + 1] : n) - rt.base[r];
        if (g >= rt.lo[r] && g - rt.lo[r] < cnt) { *local = rt.base[r] + (g - rt.lo[r]); return true; }
    }
    return false;
}

// Prepare p1 / mult / cdmax tables for what this input contains (one small D2H + sync).
int prepare_tables(ejoin_handle *h, int is_bcr) {
    std::vector<uint32_t> presL(8192 / 32), presC(65536 / 32);
    CK(cudaMemcpyAsync(presL.data(), h->d_presL.p, 8192 / 8, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(presC.data(), h->d_presC.p, 65536 / 8, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(&h->counters, h->d_counters.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (h->counters.err & ERR_CDR3_TOO_LONG) return fail(h, EJOIN_E_UNSUPPORTED, "a CDR3 exceeds 255 nt (or 256 nt for both chains)");
    if (h->counters.err & ERR_LEN_TOO_LARGE) return fail(h, EJOIN_E_UNSUPPORTED, "V..J lengths of one entry sum to 8192 or more");
    if (h->counters.err & ERR_BAD_EXACT) return fail(h, EJOIN_E_ARG, "exact_id outside [0, n_exact)");
    if (h->counters.err & ERR_ARENA_RANGE) return fail(h, EJOIN_E_ARG, "an arena offset lies outside the slice of its shard (arenas must be laid out in entry order) or outside the arena");
    // p1 tables for new L
    std::vector<uint32_t> newL;
    for (uint32_t L = 0; L < 8192; L++)
        if ((presL[L >> 5] >> (L & 31)) & 1u)
            if (!h->p1_host_ptrs[L]) newL.push_back(L);
    const size_t tab_bytes = (size_t)P1_DCAP * P1_KCAP * sizeof(double);
    const size_t BATCH = 32;           // tables built per launch: their 84 MB of scratch stay in L2 (192 at once ran 10x slower: 311 ms for 160 tables)
    if (!newL.empty()) {
        CK(h->d_p1scratch.ensure(BATCH * (size_t)P1_KCAP * P1_DCAP * sizeof(dd_t)));
        CK(h->d_Ls.ensure(BATCH * sizeof(uint32_t)));
        CK(h->d_tabptrs.ensure(BATCH * sizeof(double *)));
    }
    for (size_t b0 = 0; b0 < newL.size(); b0 += BATCH) {
        size_t nb = std::min(BATCH, newL.size() - b0);
        std::vector<double *> ptrs(nb);
        for (size_t i = 0; i < nb; i++) {
            h->p1_tables.emplace_back();
            CK(h->p1_tables.back().ensure(tab_bytes));
            ptrs[i] = (double *)h->p1_tables.back().p;
            h->p1_host_ptrs[newL[b0 + i]] = ptrs[i];
        }
        CK(cudaMemcpyAsync(h->d_Ls.p, newL.data() + b0, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, h->st));
        CK(cudaMemcpyAsync(h->d_tabptrs.p, ptrs.data(), nb * sizeof(double *), cudaMemcpyHostToDevice, h->st));
        k_p1_tables<<<(unsigned)nb, P1_DCAP, 0, h->st>>>((const uint32_t *)h->d_Ls.p, (double *const *)h->d_tabptrs.p, (dd_t *)h->d_p1scratch.p);
        h->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(h->st));          // ptrs / newL staging is reused by the next batch
    }
    if (!newL.empty()) {
        // the cache is bounded: past P1_CACHE_MAX tables, those of lengths this input does not contain are dropped (1.3 MB each)
        if (h->p1_tables.size() > P1_CACHE_MAX) {
            CK(cudaStreamSynchronize(h->st));
            std::vector<DevBuf> keep;
            for (uint32_t L = 0; L < 8192; L++) {
                if (!h->p1_host_ptrs[L]) continue;
                const bool present = (presL[L >> 5] >> (L & 31)) & 1u;
                for (auto &t : h->p1_tables)
                    if (t.p == (void *)h->p1_host_ptrs[L]) {
                        if (present) { keep.push_back(t); t.p = nullptr; t.cap = 0; } else { t.release(); h->p1_host_ptrs[L] = nullptr; }
                        break;
                    }
            }
            h->p1_tables.swap(keep);
        }
        h->d_p1scratch.release();                      // the build scratch (2.6 MB per table of the batch) is not kept
        CK(cudaMemcpyAsync(h->d_p1ptrs.p, h->p1_host_ptrs.data(), 8192 * sizeof(double *), cudaMemcpyHostToDevice, h->st));
    }
    // cd caps (and the mult tables sized by them) depend on is_bcr: rebuild when it changes
    if (h->tables_is_bcr != is_bcr) {
        build_cdmax(h, is_bcr);
        CK(cudaMemcpyAsync(h->d_cdmax.p, h->cdmax_host.data(), (C3_MAX_TOTAL + 1) * sizeof(uint16_t), cudaMemcpyHostToDevice, h->st));
        MultInfo none; none.off = 0xFFFFFFFFu; none.d1p1 = 0;
        h->multinfo_host.assign(65536, none);
        h->multpool_host.clear();
        h->tables_is_bcr = is_bcr;
    }
    // mult tables for new (cl0, cl1)
    bool newmult = false;
    for (uint32_t cc = 0; cc < 65536; cc++) {
        if (!((presC[cc >> 5] >> (cc & 31)) & 1u)) continue;
        if (h->multinfo_host[cc].off != 0xFFFFFFFFu) continue;
        const int n0 = (int)(cc >> 8), n1 = (int)(cc & 255);
        int cdm = h->cdmax_host[n0 + n1];
        if (cdm == 0xFFFF) cdm = 0;
        const int D0 = std::min(n0, cdm), D1 = std::min(n1, cdm);
        MultInfo mi; mi.off = (uint32_t)h->multpool_host.size(); mi.d1p1 = (uint32_t)(D1 + 1);
        for (int a = 0; a <= D0; a++)
            for (int b = 0; b <= D1; b++) h->multpool_host.push_back(mult_value(h->params, a, n0, b, n1));
        h->multinfo_host[cc] = mi;
        newmult = true;
    }
    if (newmult) {
        CK(h->d_multpool.ensure(h->multpool_host.size() * sizeof(double) + 64));
        CK(cudaMemcpyAsync(h->d_multpool.p, h->multpool_host.data(), h->multpool_host.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
        CK(cudaMemcpyAsync(h->d_multinfo.p, h->multinfo_host.data(), 65536 * sizeof(MultInfo), cudaMemcpyHostToDevice, h->st));
        CK(cudaStreamSynchronize(h->st));          // multpool_host may reallocate on the next call
    }
    return 0;
}

static void mark(ejoin_handle *h, const char *name) {      // one event per pipeline stage (ejoin_set_stage_timing / EJOIN_DEBUG): off in timed runs,
    if (!h->stage_timing && !h->debug) return;              // an event between two kernels keeps the second from starting under the first's tail
    if (h->n_marks == h->mark_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); h->mark_ev.push_back(e); h->mark_name.push_back(name); }
    h->mark_name[h->n_marks] = name;
    cudaEventRecord(h->mark_ev[h->n_marks++], h->st);
}
static void print_marks(ejoin_handle *h) {
    h->mark_ms.assign(h->n_marks, 0.0f);
    for (size_t i = 1; i < h->n_marks; i++) cudaEventElapsedTime(&h->mark_ms[i], h->mark_ev[i - 1], h->mark_ev[i]);
    if (!h->debug || h->n_marks < 2) return;
    const DevCounters &k = h->counters;
    fprintf(stderr, "[ejoin] n=%u cand=%llu rest=%llu todo_r1=%llu todo_r2=%llu accept_r1=%llu accept=%llu pending=%u generic=%llu tier2_pairs=%llu kept=%llu\n", h->n,
            k.n_cand, k.n_rest, k.n_todo_r1, k.n_todo, k.n_accept_r1, k.n_accept, k.n_pending, k.n_generic, k.pairs_tier2, k.n_kept);
    for (size_t i = 1; i < h->n_marks; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->mark_ev[i - 1], h->mark_ev[i]);
        fprintf(stderr, "[ejoin]   %-28s %8.3f ms\n", h->mark_name[i], ms);
    }
}

// The device pipeline on the resident input: keys -> sort -> groups -> pack -> tier 1 ->
// tier 2 -> prune -> sort of the kept edges.  Ends with the counters on the host.
int device_pipeline(ejoin_handle *h, int is_bcr, uint32_t stride, uint32_t phase) {
    if (!h->uploaded) return fail(h, EJOIN_E_ARG, "nothing uploaded");
    CK(cudaSetDevice(h->device));
    const uint32_t n = h->n;
    const DevParams pr = dev_params(h->params, is_bcr);
    memset(&h->counters, 0, sizeof(h->counters));
    h->stats.ms_kernel_pack = h->stats.ms_kernel_tier1 = h->stats.ms_kernel_tier2 = h->stats.ms_kernel_other = 0;
    if (h->cand_cap == 0) h->cand_cap = std::max<unsigned long long>(1ull << 20, 12ull * n);
    if (h->rest_cap == 0) h->rest_cap = std::max<unsigned long long>(1ull << 20, 8ull * n);
    if (h->edge_cap == 0) h->edge_cap = std::max<unsigned long long>(1ull << 20, 4ull * n);
    // tier-1 segment length: shorter on small inputs (more, smaller work items: a rank's eighth of a join otherwise has fewer
    // items than the GPU has CTA slots and the sweep's time is the time of its longest item)
    uint32_t seg_min = n >= 1000000u ? SEG_MIN : 4u;       // measured on eighths of the 1.6 M-cell join: 8 -> 4 halves the sweep; 2 and 1 cost phase A more than they save
    if (const char *e = getenv("EJOIN_SEG_MIN")) seg_min = (uint32_t)std::max(1, atoi(e));
    uint32_t mma_min_tiles = MMA_MIN_TILES_DEFAULT;        // groups of at least this many tiles (and CDR3s of at most MMA_MAX_BASES bases) sweep on the tensor cores; 0 = never
    if (const char *e = getenv("EJOIN_MMA_MIN_TILES")) mma_min_tiles = (uint32_t)std::max(0, atoi(e));
    unsigned long long i8_cap_tiles = mma_min_tiles ? (unsigned long long)n / T1_TILE + 64 : 0;      // groups that find no slot left stay with the popcount sweep
    if (const char *e = getenv("EJOIN_MMA_CAP_TILES")) if (mma_min_tiles) i8_cap_tiles = (unsigned long long)std::max(1, atoi(e));
    if (h->xrow_cap == 0 || h->xrow_seg != seg_min) { h->xrow_cap = 1024 + (unsigned long long)(n / T1_TILE) * (16 / std::min(16u, seg_min)); h->xrow_seg = seg_min; }
    if (n == 0) return 0;
    bool tables_retry = false;
    for (int attempt = 0; attempt < 6; attempt++) {
        const unsigned long long ccap = h->cand_cap, rcap = h->rest_cap;
        const unsigned long long xcap = h->xrow_cap;
        const unsigned long long cap = std::max<unsigned long long>(h->edge_cap, (unsigned long long)n + 1024);
        // ---- size and carve the work arena ----
        size_t cub_bytes = 0, tmpb = 0;
        const size_t nsort = (size_t)std::max<unsigned long long>(n, cap);
        cub::DeviceRadixSort::SortPairs(nullptr, tmpb, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                        (unsigned long long *)nullptr, nsort);
        cub_bytes = std::max(cub_bytes, tmpb);
        cub::DeviceRunLengthEncode::Encode(nullptr, tmpb, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (uint32_t *)nullptr,
                                           (uint32_t *)nullptr, (int)n);
        cub_bytes = std::max(cub_bytes, tmpb);
        cub::DeviceScan::InclusiveScan(nullptr, tmpb, (uint32_t *)nullptr, (uint32_t *)nullptr, MaxOp(), (int)n);
        cub_bytes = std::max(cub_bytes, tmpb);
        cub::DeviceScan::ExclusiveSum(nullptr, tmpb, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
        cub_bytes = std::max(cub_bytes, tmpb);
        cub::DeviceScan::ExclusiveSum(nullptr, tmpb, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n);
        cub_bytes = std::max(cub_bytes, tmpb);
        const size_t t2_units = 3 * ((size_t)(h->tig_bytes / 128) + 2 * (size_t)n) + 16;
        const size_t c3_words = (size_t)n * 2 * T1_WMAX + 4 * (size_t)n + 64;
        Sizer sz;
        sz.take(8 * (size_t)n); sz.take(8 * (size_t)n); sz.take(4 * (size_t)n); sz.take(4 * (size_t)n);   // keys, keys_out, vals, perm
        sz.take(4 * (size_t)n); sz.take(4 * (size_t)n);                                                     // head_pos, bucket_start
        sz.take(8 * (size_t)n + 8); sz.take(4 * (size_t)n + 4); sz.take(64);                                // uniq, counts, num_runs
        sz.take(sizeof(GroupInfo) * ((size_t)n + 1));
        sz.take(8 * (size_t)n); sz.take(8 * (size_t)n); sz.take(4 * (size_t)n);                             // group-table scans
        sz.take(4 * (size_t)n); sz.take(4 * (size_t)n); sz.take(4 * (size_t)n);                             // t2sizes, t2off, gid
        sz.take(sizeof(EntryMeta) * (size_t)n); sz.take(4 * c3_words); sz.take(16 * t2_units);
        sz.take(4 * (size_t)n); sz.take(4 * (size_t)n); sz.take(8 * (size_t)n); sz.take(8 * (size_t)n);           // k_to_s, parent, f0 key/val
        sz.take(16 * ccap); sz.take(8 * rcap); sz.take(8 * cap); sz.take(8 * cap); sz.take(sizeof(ejoin_raw_edge_t) * cap);       // cand, rest, todo, generic, acc
        sz.take(4 * (size_t)n * PHASE_A_K); sz.take(4 * (size_t)n); sz.take((size_t)n); sz.take((size_t)n); sz.take(4 * (size_t)n);   // phase-A slots, survivor counts, needed, force, asc list
        sz.take((size_t)n + 1024); sz.take(8 * ((size_t)n + 1024));                                           // work items: cost class, sorted list (items <= n)
        sz.take(4 * (size_t)xcap * PHASE_A_K * T1_TILE); sz.take(4 * (size_t)xcap * T1_TILE);                   // extra slot rows (segments q >= 1)
        sz.take(8 * cap); sz.take(8 * cap); sz.take(8 * cap); sz.take(8 * cap);                             // edge keys/vals in+out
        sz.take(16 * (size_t)n); sz.take(8 * cap); sz.take(4 * (size_t)n);                                  // Boruvka: two bid tables, ends, component sizes
        sz.take(cub_bytes + 256);
        sz.take(i8_cap_tiles * MMA_TILE_BYTES + 256);                                                       // int8 tiles of the tensor-core groups
        sz.take(8 * (size_t)n + 64); sz.take(sizeof(ejoin_edge_t) * ((size_t)n + 1)); sz.take(4 * (size_t)n + 64);   // payload staging (taken later)
        sz.take(4 * (size_t)n + 64); sz.take(4 * (size_t)n + 64); sz.take(4 * (size_t)h->din.n_exact + 64);   // union-find labels (in, out), first entry per exact
        CK(h->work.buf.ensure(sz.off + 4096));
        h->work.reset();
        Arena &A = h->work;
        unsigned long long *keys = A.take<unsigned long long>(n), *keys_out = A.take<unsigned long long>(n);
        uint32_t *vals = A.take<uint32_t>(n), *perm = A.take<uint32_t>(n);
        uint32_t *head_pos = A.take<uint32_t>(n), *bucket_start = A.take<uint32_t>(n);
        unsigned long long *uniq = A.take<unsigned long long>((size_t)n + 1);
        uint32_t *counts = A.take<uint32_t>((size_t)n + 1), *num_runs = A.take<uint32_t>(16);
        GroupInfo *groups = A.take<GroupInfo>((size_t)n + 1);
        unsigned long long *sumA = A.take<unsigned long long>(n), *sumB = A.take<unsigned long long>(n);
        uint32_t *sumC = A.take<uint32_t>(n);
        uint32_t *t2sizes = A.take<uint32_t>(n), *t2off = A.take<uint32_t>(n), *gid = A.take<uint32_t>(n);
        EntryMeta *meta = A.take<EntryMeta>(n);
        uint32_t *c3 = A.take<uint32_t>(c3_words);
        uint4 *t2 = A.take<uint4>(t2_units);
        uint32_t *k_to_s = A.take<uint32_t>(n), *parent = A.take<uint32_t>(n);
        unsigned long long *f0_key = A.take<unsigned long long>(n), *f0_val = A.take<unsigned long long>(n);
        uint4 *cand = A.take<uint4>(ccap);
        uint2 *rest = A.take<uint2>(rcap), *todo = A.take<uint2>(cap), *generic_q = A.take<uint2>(cap);
        uint32_t *firstk = A.take<uint32_t>((size_t)n * PHASE_A_K), *surv_cnt = A.take<uint32_t>(n);
        uint8_t *needed = A.take<uint8_t>(n), *force = A.take<uint8_t>(n);
        uint32_t *asc_list = A.take<uint32_t>(n);
        const uint32_t item_cap = n + 1024;
        uint8_t *item_cls = A.take<uint8_t>(item_cap);
        uint2 *order = A.take<uint2>(item_cap);
        uint32_t *xk = A.take<uint32_t>((size_t)xcap * PHASE_A_K * T1_TILE), *xcnt = A.take<uint32_t>((size_t)xcap * T1_TILE);
        ejoin_raw_edge_t *acc = A.take<ejoin_raw_edge_t>(cap);
        unsigned long long *pk_in = A.take<unsigned long long>(cap), *pv_in = A.take<unsigned long long>(cap);
        unsigned long long *pk_out = A.take<unsigned long long>(cap), *pv_out = A.take<unsigned long long>(cap);
        unsigned long long *best = A.take<unsigned long long>(2 * (size_t)n);   // two bid tables (k_boruvka alternates)
        uint2 *ends = A.take<uint2>(cap);
        uint32_t *comp_size = A.take<uint32_t>(n);
        void *cub_tmp = A.take<char>(cub_bytes + 256);
        int8_t *i8_tiles = A.take<int8_t>(i8_cap_tiles * MMA_TILE_BYTES + 256);
        h->meta = meta; h->c3 = c3; h->t2 = t2; h->k_to_s = k_to_s; h->kept_keys = pk_out; h->kept_vals = pv_out;
        h->w_parent = parent; h->w_best = best; h->w_comp_size = comp_size; h->w_f0_key = f0_key; h->w_f0_val = f0_val;
        DevCounters *ctr = (DevCounters *)h->d_counters.p;
        const unsigned nb256 = (n + 255) / 256;
        cudaStream_t st = h->st;

        if (h->copy_pending) CK(cudaStreamWaitEvent(st, h->ev_early, 0));
        h->debug = getenv("EJOIN_DEBUG") != nullptr; h->n_marks = 0;
        CK(cudaEventRecord(h->ev[0], st));
        mark(h, "start");
        CK(cudaMemsetAsync(ctr, 0, sizeof(DevCounters), st));
        CK(cudaMemsetAsync(h->d_presL.p, 0, 8192 / 8, st));
        CK(cudaMemsetAsync(h->d_presC.p, 0, 65536 / 8, st));
        k_bucket_heads<<<nb256, 256, 0, st>>>(h->din, head_pos);
        size_t tb = cub_bytes;
        cub::DeviceScan::InclusiveScan(cub_tmp, tb, head_pos, bucket_start, MaxOp(), (int)n, st);
        k_entry_keys<<<std::min(nb256, (unsigned)h->sm_count * 4u), 256, 0, st>>>(h->din, bucket_start, keys, vals, (uint32_t *)h->d_presL.p, (uint32_t *)h->d_presC.p, ctr);
        h->launches += 2;
        CK(cudaGetLastError());
        // Tables (p1 per total length, multipliers per CDR3 length pair) are cached on the handle.  Once a handle has built
        // tables, the key pass is not waited for any more: the pipeline runs on, a pair that finds its table missing raises
        // ERR_TABLES_MISSING (decide_score), and the check after the forest builds what is missing and reruns the join.
        const bool optimistic = h->tables_built && h->tables_is_bcr == is_bcr && !tables_retry && getenv("EJOIN_SYNC_TABLES") == nullptr;
        DevCounters first;
        memset(&first, 0, sizeof(first));
        if (!optimistic) {
            int rc = prepare_tables(h, is_bcr);     // one sync
            if (rc) return rc;
            first = h->counters;
            h->tables_built = true;
            tables_retry = false;
        }
        mark(h, "keys (+tables first run)");
        tb = cub_bytes;
        int entry_bits = 17;                        // 16 bits of CDR3 lengths + the bits of a bucket's first index (invalid entries: all ones, behind every valid key)
        while (entry_bits < 48 && (1ull << (entry_bits - 16)) < (unsigned long long)n) entry_bits++;
        cub::DeviceRadixSort::SortPairs(cub_tmp, tb, keys, keys_out, vals, perm, (int)n, 0, entry_bits, st);
        mark(h, "sort entries");
        tb = cub_bytes;
        cub::DeviceRunLengthEncode::Encode(cub_tmp, tb, keys_out, uniq, counts, num_runs, (int)n, st);
        k_group_vals<<<nb256, 256, 0, st>>>(uniq, counts, num_runs, n, (const uint16_t *)h->d_cdmax.p, sumA, sumB, sumC, ctr, seg_min);
        tb = cub_bytes;
        cub::DeviceScan::ExclusiveSum(cub_tmp, tb, sumA, sumA, (int)n, st);
        tb = cub_bytes;
        cub::DeviceScan::ExclusiveSum(cub_tmp, tb, sumB, sumB, (int)n, st);
        tb = cub_bytes;
        cub::DeviceScan::ExclusiveSum(cub_tmp, tb, sumC, sumC, (int)n, st);
        k_group_build<<<nb256, 256, 0, st>>>(uniq, counts, num_runs, (const uint16_t *)h->d_cdmax.p, sumA, sumB, sumC, groups, ctr, seg_min, mma_min_tiles, i8_cap_tiles);
        h->launches += 2;                           // k_group_vals, k_group_build (own kernels only: CUB's sort / RLE / scan passes are never counted)
        mark(h, "rle + group table");
        k_item_hist<<<h->sm_count, 256, 0, st>>>(groups, ctr, item_cls, item_cap, stride, phase);
        k_item_scatter<<<h->sm_count, 256, 0, st>>>(groups, ctr, item_cls, order, item_cap, stride, phase);
        h->launches += 2;
        mark(h, "work items by cost");
        k_sorted_heads<<<nb256, 256, 0, st>>>(keys_out, n, t2sizes);          // t2sizes doubles as the flag buffer
        tb = cub_bytes;
        cub::DeviceScan::InclusiveSum(cub_tmp, tb, t2sizes, gid, (int)n, st);
        h->launches += 1;
        k_t2_sizes<<<nb256, 256, 0, st>>>(h->din, perm, ctr, t2sizes);
        tb = cub_bytes;
        cub::DeviceScan::ExclusiveSum(cub_tmp, tb, t2sizes, t2off, (int)n, st);
        h->launches += 1;                           // k_t2_sizes
        CK(cudaEventRecord(h->ev[1], st));
        mark(h, "heads/scans/t2 sizes");
        k_pack_c3<<<(unsigned)(((size_t)n * 4 + 255) / 256), 256, 0, st>>>(h->din, perm, gid, groups, ctr, c3);
        // stride mode: the gathered edges of the other ranks need this rank's rows too -> pack every entry
        k_init_rows<<<nb256, 256, 0, st>>>(parent, f0_key, surv_cnt, needed, force, n, stride > 1 ? 1 : 0);
        if (mma_min_tiles) { k_pack_i8<<<nb256, 256, 0, st>>>(c3, gid, groups, ctr, i8_tiles); h->launches += 1; }     // returns at once without a tensor-core group
        CK(cudaEventRecord(h->ev[2], st));
        mark(h, "k_pack_c3 + k_init_rows");
        // tier 1: needs only the c3 rows, so it runs while the V..J arena is still on its way
        int rows_occ = 8;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rows_occ, k_sweep, T1_TILE, 0);
        if (rows_occ < 1) rows_occ = 1;
        k_sweep<<<h->sm_count * rows_occ, T1_TILE, 0, st>>>(c3, groups, ctr, order, cand, ccap, firstk, surv_cnt, xk, xcnt, xcap, needed);
        mark(h, "k_sweep");
        if (mma_min_tiles) {
            // the large groups of short CDR3s on the tensor cores (the kernel returns at once when the group table flagged none)
            const size_t mma_smem = (1 + MMA_NA) * (size_t)MMA_TILE_BYTES;
            if (h->mma_grid == 0) {
                CK(cudaFuncSetAttribute(k_sweep_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_smem));
                CK(cudaFuncSetAttribute(k_sweep_mma, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
                // two CTAs per SM: 2 x 113 KB of shared memory, 2 x 256 of the 512 TMEM columns.  (Measured with the prototype: two per SM
                // sweep 1.67x faster than one, although the occupancy API answers 1 for this kernel; CTAs beyond what is resident
                // would only find the work list empty.)
                h->mma_grid = h->sm_count * 2;
            }
            k_sweep_mma<<<h->mma_grid, MMA_THREADS, mma_smem, st>>>(i8_tiles, groups, ctr, order, cand, ccap, firstk, surv_cnt, xk, xcnt, xcap, needed);
            h->launches += 1;
        }
        CK(cudaEventRecord(h->ev[3], st));
        mark(h, "k_sweep_mma");
        if (h->copy_pending) CK(cudaStreamWaitEvent(st, h->ev_small, 0));          // the late arrays (ids, offsets, bounds, germlines)
        if (h->recs_pending) {
            k_unpack_recs<<<nb256, 256, 0, st>>>(h->d_rec, n, h->rec_out);
            h->launches += 1;
            h->recs_pending = false;               // the parallel arrays stay valid for later resident runs
        }
        k_meta<<<nb256, 256, 0, st>>>(h->din, perm, gid, groups, t2off, ctr, meta, k_to_s, needed, pr.ref_v_trim, pr.ref_j_trim, asc_list);
        k_pack_ref<<<std::max(1u, h->din.n_refseq), 32, 0, st>>>(h->din);
        h->launches += 1;
        CK(cudaEventRecord(h->ev[7], st));
        mark(h, "k_meta");
        // 2-bit input: k_pack_t2 takes only the entries with an ASCII chain, which k_meta listed (~1 % of them; the other half
        // warps of the grid leave at once), k_pack_t2_2bit all the others
        const uint32_t *pack_list = h->din.tig2_src ? asc_list : nullptr;
        const unsigned pack_blocks = h->din.tig2_src ? (unsigned)h->sm_count * 4u : (unsigned)(((size_t)n * 16 + 255) / 256);
        if (h->copy_pending && h->n_chunks > 0) {
            // pack each chunk's (needed) entries as its bytes land (an entry belongs to the chunk that completes it)
            uint64_t prev = 0;
            for (int i = 0; i < h->n_chunks; i++) {
                CK(cudaStreamWaitEvent(st, h->ev_chunk[i == h->n_chunks - 1 ? 7 : i], 0));
                k_pack_t2<<<pack_blocks, 256, 0, st>>>(h->din, ctr, meta, t2, needed, prev, h->chunk_end[i], pack_list);
                prev = h->chunk_end[i];
            }
            h->launches += h->n_chunks - 1;
        } else {
            k_pack_t2<<<pack_blocks, 256, 0, st>>>(h->din, ctr, meta, t2, needed, 0ull, ~0ull, pack_list);
        }
        if (h->din.tig2_src) {       // entries with both chains 2 bits per base (the launches above did the ones with an ASCII chain)
            k_pack_t2_2bit<<<(unsigned)(((size_t)n * 8 + 255) / 256), 256, 0, st>>>(h->din, ctr, meta, t2, needed);
            h->launches += 1;
        }
        CK(cudaEventRecord(h->ev[6], st));
        mark(h, "k_pack_t2");
        DevTables tbs;
        tbs.meta = meta; tbs.c3 = c3; tbs.t2 = t2; tbs.p1tab = (const double *const *)h->d_p1ptrs.p;
        tbs.multinfo = (const MultInfo *)h->d_multinfo.p; tbs.multpool = (const double *)h->d_multpool.p;
        tbs.cdmax = (const uint16_t *)h->d_cdmax.p;
        SlotView sv;
        sv.firstk = firstk; sv.cnt = surv_cnt; sv.xk = xk; sv.xcnt = xcnt; sv.gid_incl = gid; sv.groups = groups; sv.stride = stride; sv.phase = phase;
        k_phase_a<<<(n + 127) / 128, 128, 0, st>>>(h->din, tbs, pr, ctr, sv, force, parent, f0_key, f0_val, cand, ccap);
        mark(h, "k_phase_a");
        h->launches += 6;                           // k_pack_c3, k_init_rows, k_sweep, k_meta, k_pack_t2, k_phase_a
        // phase B round 1: entries that gave up in phase A get all their survivors scored; the smallest
        // accepted partner becomes the tree edge (atomicMin on the parent)
        k_phaseb_split<<<h->sm_count * 8, 256, 0, st>>>(cand, ccap, parent, force, ctr, todo, cap, rest, rcap);
        mark(h, "k_phaseb_split");
        k_tier2<<<h->sm_count * 8, 128, 0, st>>>(h->din, tbs, pr, ctr, todo, cap, generic_q, acc, 0, parent);
        mark(h, "k_tier2 r1");
        k_tier2_generic<<<h->sm_count * 8, 128, 0, st>>>(h->din, tbs, pr, ctr, generic_q, cap, acc, 0, parent, 0);
        mark(h, "k_tier2_generic r1");
        k_forced_resolve<<<h->sm_count * 4, 256, 0, st>>>(acc, ctr, cap, k_to_s, 0, parent, f0_key, f0_val);
        k_root_all<<<nb256, 256, 0, st>>>(parent, n);
        mark(h, "resolve + root jumps");
        // round 2: the other recorded survivors, scored only across trees
        k_round2_begin<<<1, 1, 0, st>>>(ctr);
        k_forced_prune<<<h->sm_count * 4, 256, 0, st>>>(acc, ctr, cap, k_to_s, 0, parent);
        k_phaseb_cross<<<h->sm_count * 8, 256, 0, st>>>(rest, rcap, parent, ctr, todo, cap);
        mark(h, "round2 begin/prune/cross");
        k_tier2<<<h->sm_count * 8, 128, 0, st>>>(h->din, tbs, pr, ctr, todo, cap, generic_q, acc, 0, nullptr);
        mark(h, "k_tier2 r2");
        k_tier2_generic<<<h->sm_count * 8, 128, 0, st>>>(h->din, tbs, pr, ctr, generic_q, cap, acc, 0, nullptr, 1);
        h->launches += 10;                          // phase B: split, 2 x (k_tier2, k_tier2_generic), resolve, roots, round-2 begin, prune, cross
        CK(cudaEventRecord(h->ev[4], st));
        mark(h, "k_tier2_generic r2");
        const bool device_final = stride == 1;      // stride mode: the forest needs every rank's edges -> host replay after the gather
        h->device_final = device_final;
        if (device_final) {
            k_fill_u64<<<2 * nb256, 256, 0, st>>>(best, ~0ull, 2 * (size_t)n);
            {   // all Boruvka rounds in one cooperative launch (grid = what is resident at once)
                if (h->bor_grid == 0) {
                    int occ = 1;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_boruvka, 256, 0);
                    h->bor_grid = occ >= 1 ? h->sm_count : 1;          // one CTA per SM: the grid-wide barriers are the cost, not the edge loop
                }
                ejoin_raw_edge_t *a_acc = acc; DevCounters *a_ctr = ctr; unsigned long long a_cap = cap; const uint32_t *a_kts = k_to_s;
                uint32_t *a_par = parent; unsigned long long *a_best = best; uint2 *a_ends = ends; uint32_t a_n = n;
                void *args[] = { &a_acc, &a_ctr, &a_cap, &a_kts, &a_par, &a_best, &a_ends, &a_n };
                CK(cudaLaunchCooperativeKernel((void *)k_boruvka, dim3(h->bor_grid), dim3(256), args, 0, st));
            }
            h->launches += 2;
        }
        CK(cudaGetLastError());
        mark(h, "boruvka");
        // The collect kernels run before the host has looked at the counters (they are bounded by the capacities, so an
        // overflowed list cannot hurt them): one wait instead of two.
        if (device_final) {
            CK(cudaMemsetAsync(comp_size, 0, 4 * (size_t)n, st));
            k_comp_size<<<nb256, 256, 0, st>>>(parent, ctr, comp_size);
            k_final_collect<<<h->sm_count * 8, 256, 0, st>>>(h->din, f0_key, f0_val, n, acc, ctr, cap, k_to_s, 0, parent, comp_size, pr.easy, pk_in, pv_in);
            h->launches += 2;
        } else {
            k_collect<<<h->sm_count * 8, 256, 0, st>>>(f0_key, f0_val, n, acc, ctr, cap, pk_in, pv_in);
            h->launches += 1;
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&h->counters, ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
        mark(h, "comp size + collect");
        CK(cudaStreamSynchronize(st));
        if (optimistic) {
            // what prepare_tables checks at its own sync
            if (h->counters.err & ERR_CDR3_TOO_LONG) return fail(h, EJOIN_E_UNSUPPORTED, "a CDR3 exceeds 255 nt (or 256 nt for both chains)");
            if (h->counters.err & ERR_LEN_TOO_LARGE) return fail(h, EJOIN_E_UNSUPPORTED, "V..J lengths of one entry sum to 8192 or more");
            if (h->counters.err & ERR_BAD_EXACT) return fail(h, EJOIN_E_ARG, "exact_id outside [0, n_exact)");
            if (h->counters.err & ERR_TABLES_MISSING) { tables_retry = true; continue; }       // the next attempt takes the waiting path and builds them
        } else { h->counters.pairs_candidate = first.pairs_candidate; h->counters.n_buckets = first.n_buckets; }
        if (h->counters.err & ERR_CDR3_BAD_CHAR) return fail(h, EJOIN_E_UNSUPPORTED, "a CDR3 contains a character outside ACGT");
        if (h->counters.err & ERR_BAD_REFSEQ) return fail(h, EJOIN_E_ARG, "vs_seq/js_seq index outside the germline table");
        if (h->counters.err & ERR_ARENA_RANGE) return fail(h, EJOIN_E_ARG, "an arena offset lies outside the slice of its shard (arenas must be laid out in entry order) or outside the arena");
        if (h->counters.err & ERR_K_TOO_LARGE) return fail(h, EJOIN_E_UNSUPPORTED, "a pair has 1024 or more mutations (p1 table width)");
        if (h->counters.err & ERR_GROUP_TOO_LARGE) return fail(h, EJOIN_E_UNSUPPORTED, "more than 8.3 M entries share all four lengths (one comparison group)");
        {
            // a buffer was too small: the counters keep counting past the capacity, so they tell the size to retry with
            // (a list that overflowed may have starved the lists behind it, hence the loop)
            const DevCounters &k = h->counters;
            const unsigned long long need_e = std::max(std::max(k.n_todo, k.n_todo_r1), std::max(k.n_generic_q, (unsigned long long)n + k.n_accept));
            bool retry = false;
            if (k.n_cand > ccap) { h->cand_cap = k.n_cand + k.n_cand / 8 + 1024; retry = true; }
            if (k.n_rest > rcap) { h->rest_cap = k.n_rest + k.n_rest / 8 + 1024; retry = true; }
            if (need_e > cap) { h->edge_cap = need_e + need_e / 4 + 1024; retry = true; }
            if (k.n_xrows > xcap) { h->xrow_cap = k.n_xrows + 64; retry = true; }
            if (retry) continue;
        }
        const unsigned long long nk = h->counters.n_kept;
        if (nk) {
            // keys are (k1 << 32 | k2) with k < n: only the bits that can differ are sorted
            int kb = 1;
            while (kb < 32 && (1ull << kb) < (unsigned long long)n) kb++;
            tb = cub_bytes;
            cub::DeviceRadixSort::SortPairs(cub_tmp, tb, pk_in, pk_out, pv_in, pv_out, (int)nk, 0, 32 + kb, st);
        }
        CK(cudaEventRecord(h->ev[5], st));
        mark(h, "sort edges");
        CK(cudaStreamSynchronize(st));
        print_marks(h);
        float ms;
        float ms2;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]); cudaEventElapsedTime(&ms2, h->ev[3], h->ev[6]); h->stats.ms_kernel_pack = ms + ms2;   // meta + c3, then t2 rows
        cudaEventElapsedTime(&ms, h->ev[7], h->ev[6]); h->stats.ms_kernel_pack_t2 = ms;
        cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->stats.ms_kernel_tier1 = ms;                                                        // the sweep
        cudaEventElapsedTime(&ms, h->ev[6], h->ev[4]); h->stats.ms_kernel_tier2 = ms;                                                        // phase A + phase B
        float tot;
        cudaEventElapsedTime(&tot, h->ev[0], h->ev[5]);
        h->stats.ms_device = tot;
        h->stats.ms_kernel_other = tot - h->stats.ms_kernel_pack - h->stats.ms_kernel_tier1 - h->stats.ms_kernel_tier2;
        if (h->copy_pending) {
            float up = 0;
            cudaEventElapsedTime(&up, h->ev_copy0, h->n_chunks > 0 ? h->ev_chunk[7] : h->ev_small);
            h->stats.ms_h2d = up;                 // H2D on the copy stream, overlapped with the stages before the join
            h->copy_pending = false;
        }
        return 0;
    }
    return fail(h, EJOIN_E_NOMEM, "candidate buffer kept overflowing");
}

void fill_stats(ejoin_handle *h, ejoin_stats_t *s) {
    const DevCounters &c = h->counters;
    s->n_buckets = c.n_buckets; s->n_groups = c.n_groups; s->pairs_candidate = c.pairs_candidate; s->pairs_compared = c.pairs_compared;
    s->pairs_tier2 = c.pairs_tier2; s->pairs_accepted = c.n_accept; s->pairs_generic = c.n_generic; s->edges_after_prune = c.n_kept;
    s->gpu_launches = h->launches; s->h2d_bytes = h->h2d_bytes;
    if (h->tigs_in_place) { unsigned long long ib = 0; for (int i = 0; i < 32; i++) ib += c.inplace_bytes[i]; s->h2d_bytes += ib; }   // bytes k_pack_t2 pulled over PCIe
    s->alg_bytes_tier1 = c.t1_alg_bytes; s->alg_bytes_tier2 = c.t2_alg_bytes; s->mma_pairs = c.mma_pairs; s->mma_macs = c.mma_macs;
    { unsigned long long pb = 0; for (int i = 0; i < 32; i++) pb += c.pack_bytes[i]; s->alg_bytes_pack_t2 = pb; }
    s->ms_kernel_pack_t2 = h->stats.ms_kernel_pack_t2;
    s->alg_bytes_pack = h->tig_bytes + c.n_valid * sizeof(EntryMeta) + c.c3_words * 4 + c.t2_units * 16;
    s->ms_kernel_pack = h->stats.ms_kernel_pack; s->ms_kernel_tier1 = h->stats.ms_kernel_tier1; s->ms_kernel_tier2 = h->stats.ms_kernel_tier2;
    s->ms_kernel_other = h->stats.ms_kernel_other; s->ms_device = h->stats.ms_device;
}

struct UF {
    std::vector<uint32_t> p;
    explicit UF(size_t n) : p(n) { for (size_t i = 0; i < n; i++) p[i] = (uint32_t)i; }
    uint32_t find(uint32_t a) {
        uint32_t r = a;
        while (p[r] != r) r = p[r];
        while (p[a] != r) { uint32_t nx = p[a]; p[a] = r; a = nx; }
        return r;
    }
    bool join(uint32_t a, uint32_t b) {          // the smaller index stays the root: labels = min member
        a = find(a); b = find(b);
        if (a == b) return false;
        if (a < b) p[b] = a; else p[a] = b;
        return true;
    }
};

// Ordered replay over the accepted edges sorted by (k1,k2): join_core's skip-if-same-class
// (A.3/A.8), the two-cell rule (A.6) and the emit (A.7).  Returns the emitted (k1,k2).
// Work is O(#edges): the union-find lives in the handle and only the touched entries are reset.
void replay(ejoin_handle *h, const ejoin_pack_t *pk, const unsigned long long *keys, const unsigned long long *vals, size_t ne,
            std::vector<uint2> &final_pairs) {
    std::vector<uint32_t> &p = h->uf_parent;
    if (p.size() < pk->n_info) {
        size_t old = p.size();
        p.resize(pk->n_info);
        for (size_t i = old; i < p.size(); i++) p[i] = (uint32_t)i;
    }
    auto find = [&](uint32_t a) {
        uint32_t r = a;
        while (p[r] != r) r = p[r];
        while (p[a] != r) { uint32_t nx = p[a]; p[a] = r; a = nx; }
        return r;
    };
    std::vector<size_t> &forest = h->forest_tmp;
    forest.clear();
    for (size_t i = 0; i < ne; i++) {
        uint32_t a = find((uint32_t)(keys[i] >> 32)), b = find((uint32_t)keys[i]);
        if (a == b) continue;
        if (a < b) p[b] = a; else p[a] = b;
        forest.push_back(i);
    }
    // orbit size = forest edges of the orbit + 1; count edges per root (roots are touched entries)
    std::vector<uint32_t> &cnt = h->cnt_tmp;
    if (cnt.size() < pk->n_info) cnt.resize(pk->n_info, 0);
    for (size_t i : forest) cnt[find((uint32_t)(keys[i] >> 32))]++;
    final_pairs.clear();
    final_pairs.reserve(forest.size());
    for (size_t i : forest) {
        uint32_t k1 = (uint32_t)(keys[i] >> 32), k2 = (uint32_t)keys[i];
        if (cnt[find(k1)] == 1) {                          // an orbit with exactly two entries
            uint32_t cells = pk->ncells[pk->exact_id[k1]] + pk->ncells[pk->exact_id[k2]];
            uint32_t cd = (uint32_t)(vals[i] & 0xFFFF), d = (uint32_t)((vals[i] >> 16) & 0xFFFF);
            if (cells == 2 && !h->params.easy && !(cd <= d)) continue;     // help1.rs:349-353
        }
        final_pairs.push_back(make_uint2(k1, k2));
    }
    // reset what was touched
    for (size_t i : forest) { uint32_t k1 = (uint32_t)(keys[i] >> 32), k2 = (uint32_t)keys[i]; cnt[find(k1)] = 0; }
    for (size_t i : forest) { uint32_t k1 = (uint32_t)(keys[i] >> 32), k2 = (uint32_t)keys[i]; p[k1] = k1; p[k2] = k2; }
}

// finish_join on the host (used by ejoin_partition and when the handle does not hold the whole input)
void host_partition(const ejoin_pack_t *pk, const uint32_t *k1, const uint32_t *k2, size_t ne, uint32_t *class_of, uint32_t *n_classes) {
    const uint32_t n = pk->n_info;
    UF uf(n);
    for (size_t i = 0; i < ne; i++) uf.join(k1[i], k2[i]);
    std::vector<int64_t> first(pk->n_exact, -1);
    for (uint32_t k = 0; k < n; k++) {
        uint32_t x = pk->exact_id[k];
        if (x >= pk->n_exact) continue;
        if (first[x] < 0) first[x] = k; else uf.join((uint32_t)first[x], k);
    }
    uint32_t ncls = 0;
    for (uint32_t k = 0; k < n; k++) { class_of[k] = uf.find(k); ncls += class_of[k] == k; }
    if (n_classes) *n_classes = ncls;
}

// Emitted edges -> full PotentialJoin records (k_payload) and, when the handle holds the whole
// input, the finish_join partition by the GPU union-find; both land in the handle's pinned
// result buffer.
int emit_result(ejoin_handle *h, const ejoin_pack_t *pk, const std::vector<uint2> &pairs, const unsigned long long *d_keys, size_t ne_dev,
                size_t pin_used, ejoin_result_t *out) {
    const size_t ne = d_keys ? ne_dev : pairs.size();
    const uint32_t n = h->n;
    const bool whole = h->rt.n == 1 && h->rt.lo[0] == 0 && h->n == pk->n_info;
    const DevParams pr = dev_params(h->params, (int)pk->is_bcr);
    char *pin = (char *)h->pin;
    size_t off = (pin_used + 255) & ~(size_t)255;
    ejoin_edge_t *h_edges = (ejoin_edge_t *)(pin + off);
    off += sizeof(ejoin_edge_t) * (ne ? ne : 1);
    off = (off + 255) & ~(size_t)255;
    uint32_t *h_labels = (uint32_t *)(pin + off);
    off += sizeof(uint32_t) * (size_t)(n ? n : 1);
    if (off > h->pin_cap) return fail(h, EJOIN_E_NOMEM, "pinned result buffer too small");
    out->n_edges = ne; out->edges = h_edges; out->owner = h;
    out->class_of = nullptr; out->n_info = 0; out->n_classes = 0;
    if (n == 0) return 0;                                  // nothing uploaded: empty result
    Arena &A = h->work;
    const size_t save = A.off;
    uint2 *d_pairs = A.take<uint2>(ne ? ne : 1);
    ejoin_edge_t *d_out = A.take<ejoin_edge_t>(ne ? ne : 1);
    uint32_t *d_slow = A.take<uint32_t>(ne ? ne : 1);
    uint32_t *d_label = A.take<uint32_t>(n ? n : 1), *d_label_out = A.take<uint32_t>(n ? n : 1);
    uint32_t *d_first = A.take<uint32_t>(pk->n_exact ? pk->n_exact : 1);
    if (A.off > A.buf.cap) { A.off = save; return fail(h, EJOIN_E_NOMEM, "work arena too small for the epilogue"); }
    DevCounters *ctr = (DevCounters *)h->d_counters.p;
    const double t_dev0 = now_ms();
    if (ne) {
        if (d_keys) k_keys_to_pairs<<<(unsigned)((ne + 255) / 256), 256, 0, h->st>>>(d_keys, (uint32_t)ne, h->rt, d_pairs);
        else CK(cudaMemcpyAsync(d_pairs, pairs.data(), ne * sizeof(uint2), cudaMemcpyHostToDevice, h->st));
        DevTables tbs;
        tbs.meta = h->meta; tbs.c3 = h->c3; tbs.t2 = h->t2; tbs.p1tab = (const double *const *)h->d_p1ptrs.p;
        tbs.multinfo = (const MultInfo *)h->d_multinfo.p; tbs.multpool = (const double *)h->d_multpool.p; tbs.cdmax = (const uint16_t *)h->d_cdmax.p;
        CK(cudaMemsetAsync(&ctr->n_payload_slow, 0, sizeof(unsigned int), h->st));
        k_payload_fast<<<h->sm_count * 8, 128, 0, h->st>>>(h->din, tbs, pr, ctr, h->k_to_s, d_pairs, (uint32_t)ne, h->rt, h->n, d_out, d_slow);
        k_payload<<<h->sm_count * 4, 128, 0, h->st>>>(h->din, tbs, pr, ctr, h->k_to_s, d_pairs, d_slow, h->rt, h->n, d_out);
        h->launches += d_keys ? 3 : 2;               // (k_keys_to_pairs,) k_payload_fast, k_payload
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_edges, d_out, ne * sizeof(ejoin_edge_t), cudaMemcpyDeviceToHost, h->st));
    }
    if (whole && n) {
        const unsigned nbx = (std::max(n, pk->n_exact) + 255) / 256, nbn = (std::max<uint32_t>(n, (uint32_t)ne) + 255) / 256;
        CK(cudaMemsetAsync(&ctr->n_classes, 0, sizeof(unsigned long long), h->st));
        k_uf_init<<<nbx, 256, 0, h->st>>>(d_label, n, d_first, pk->n_exact);
        k_uf_first<<<(n + 255) / 256, 256, 0, h->st>>>(h->din.exact_id, n, pk->n_exact, d_first);
        k_uf_union<<<nbn, 256, 0, h->st>>>(d_label, d_pairs, (uint32_t)ne, 0, h->din.exact_id, n, pk->n_exact, d_first);
        k_uf_flatten<<<(n + 255) / 256, 256, 0, h->st>>>(d_label, n, ctr, d_label_out);
        h->launches += 4;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_labels, d_label_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
        CK(cudaMemcpyAsync(&h->counters.n_classes, &ctr->n_classes, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
    }
    CK(cudaStreamSynchronize(h->st));
    if (getenv("EJOIN_DEBUG")) fprintf(stderr, "[ejoin] epilogue: %zu edges, device part %.3f ms\n", ne, now_ms() - t_dev0);
    A.off = save;
    if (whole) { out->class_of = h_labels; out->n_info = n; out->n_classes = (uint32_t)h->counters.n_classes; }
    h->stats.d2h_bytes += ne * sizeof(ejoin_edge_t) + (whole ? (size_t)n * 4 : 0);
    return 0;
}

}  // namespace

extern "C" int ejoin_upload(ejoin_handle_t *h, const ejoin_pack_t *pk) {
    if (!h) return EJOIN_E_ARG;
    int rc = upload_range(h, pk, 0, pk ? pk->n_info : 0, /*allow_in_place=*/false);   // resident means resident
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->st_copy));
    h->copy_pending = false;
    h->last_stride = 1; h->last_phase = 0;
    h->is_bcr = (int)pk->is_bcr;
    return 0;
}

static bool shard_plan(const ejoin_pack_t *pk, int rank, int world, Ranges &out);

extern "C" int ejoin_upload_shard(ejoin_handle_t *h, const ejoin_pack_t *pk, int rank, int world) {
    if (!h || !pk || rank < 0 || rank >= world) return EJOIN_E_ARG;
    Ranges rg;
    const bool stride_mode = shard_plan(pk, rank, world, rg);
    int rc = upload_ranges(h, pk, rg, /*allow_in_place=*/false);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->st_copy));
    h->copy_pending = false;
    h->last_stride = stride_mode ? (uint32_t)world : 1u; h->last_phase = stride_mode ? (uint32_t)rank : 0u;
    h->is_bcr = (int)pk->is_bcr;
    return 0;
}

extern "C" int ejoin_run_resident(ejoin_handle_t *h, ejoin_stats_t *stats) {
    if (!h) return EJOIN_E_ARG;
    // an in-place run leaves din.tig_src pointing at the caller's arena, which the caller owns only during that call
    if (h->uploaded && h->tigs_in_place) return fail(h, EJOIN_E_ARG, "ejoin_run_resident needs ejoin_upload / ejoin_upload_shard first (the last run read the caller's arena in place)");
    h->launches = 0;
    int rc = device_pipeline(h, h->is_bcr, h->last_stride, h->last_phase);
    if (rc) return rc;
    if (stats) { memset(stats, 0, sizeof(*stats)); fill_stats(h, stats); }
    return 0;
}

static int finish_common(ejoin_handle *h, const ejoin_pack_t *pk, const unsigned long long *keys, const unsigned long long *vals, size_t nk,
                         size_t pin_used, ejoin_result_t *out) {
    double t0 = now_ms();
    std::vector<uint2> fin;
    replay(h, pk, keys, vals, nk, fin);
    if (getenv("EJOIN_DEBUG")) fprintf(stderr, "[ejoin] replay: %zu sorted edges -> %zu emitted, %.3f ms\n", nk, fin.size(), now_ms() - t0);
    int rc = emit_result(h, pk, fin, nullptr, 0, pin_used, out);
    if (rc) return rc;
    h->stats.ms_host_post = now_ms() - t0;
    return 0;
}

extern "C" int ejoin_run(ejoin_handle_t *h, const ejoin_pack_t *pk, ejoin_result_t *out) {
    if (!h || !out) return EJOIN_E_ARG;
    memset(out, 0, sizeof(*out));
    memset(&h->stats, 0, sizeof(h->stats));
    h->launches = 0;
    const double t0 = now_ms();
    int rc = upload_range(h, pk, 0, pk ? pk->n_info : 0);
    if (rc) return rc;
    h->is_bcr = (int)pk->is_bcr;
    const double t1 = now_ms();                 // copies are in flight on the copy stream; the pipeline waits on their events
    h->last_stride = 1; h->last_phase = 0;
    rc = device_pipeline(h, (int)pk->is_bcr, 1, 0);
    if (rc) return rc;
    const double t2 = now_ms();
    const size_t nk = (size_t)h->counters.n_kept;
    rc = ensure_pin(h, sizeof(ejoin_edge_t) * (nk + 1) + 4 * ((size_t)h->n + 1) + 1024);
    if (rc) return rc;
    h->stats.d2h_bytes = sizeof(DevCounters);
    const double t3 = now_ms();
    {   // the sorted keys on the device already are raw_joins: payload + partition, no host replay
        static const std::vector<uint2> none;
        rc = emit_result(h, pk, none, h->kept_keys, nk, 0, out);
        if (rc) return rc;
    }
    const double t4 = now_ms();
    fill_stats(h, &out->stats);
    (void)t1;
    out->stats.ms_h2d = h->stats.ms_h2d; out->stats.ms_d2h = t3 - t2; out->stats.ms_host_post = t4 - t3; out->stats.ms_total = t4 - t0;
    out->stats.d2h_bytes = h->stats.d2h_bytes;
    out->stats.gpu_launches = h->launches;
    return 0;
}

extern "C" int ejoin_set_stage_timing(ejoin_handle_t *h, int on) {
    if (!h) return EJOIN_E_ARG;
    h->stage_timing = on != 0;
    if (!on) { h->mark_ms.clear(); h->n_marks = 0; }
    return EJOIN_OK;
}

extern "C" int ejoin_stage_time(const ejoin_handle_t *h, int index, const char **name, double *ms) {
    if (!h || index < 0 || (size_t)index + 1 >= h->mark_ms.size() || (size_t)index + 1 >= h->n_marks) return EJOIN_E_ARG;
    if (name) *name = h->mark_name[index + 1];
    if (ms) *ms = h->mark_ms[index + 1];
    return EJOIN_OK;
}

extern "C" void ejoin_result_free(ejoin_result_t *r) {
    if (!r) return;
    if (!r->owner) { free(r->edges); free(r->class_of); }     // owner != NULL: the handle's pinned buffer
    memset(r, 0, sizeof(*r));
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU decomposition: cost-balanced bucket ranges, dealt to the ranks in turn.
// ---------------------------------------------------------------------------------------------
#define SHARD_CHUNKS 4          // bucket ranges per rank
// bucket starts (runs of equal lens) and their contiguous, cost-balanced owners; one pass over the lens arrays
static void plan_buckets(const ejoin_pack_t *pk, int world, std::vector<uint32_t> &bs, std::vector<uint32_t> &owner) {
    bs.clear();
    const uint16_t *l0 = pk->len[0], *l1 = pk->len[1];
    // `info` is sorted with lens first (CloneInfo's Ord), so a bucket is found by galloping + bisection from its head instead of a
    // scan of all entries (a host scan of 1.3 M entries costs a millisecond per call).  Every boundary reported is a true run
    // boundary (len differs across it) whatever the input, which is all the sharding needs: a rank's range is a union of runs.
    const uint32_t n_all = pk->n_info;
    auto same = [&](uint32_t a, uint32_t b) { return l0[a] == l0[b] && l1[a] == l1[b]; };
    for (uint32_t k = 0; k < n_all;) {
        bs.push_back(k);
        uint32_t lo = k, step = 1;
        while (lo + step < n_all && same(k, lo + step)) { lo += step; step <<= 1; }
        uint32_t hi = std::min<uint64_t>((uint64_t)lo + step, n_all);      // same(k, lo) holds, hi == n_all or !same(k, hi)
        while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (same(k, mid)) lo = mid; else hi = mid; }
        k = hi;
    }
    const uint32_t nb = (uint32_t)bs.size();
    bs.push_back(pk->n_info);
    owner.assign(nb, 0);
    // cost = candidate pairs + a linear term for packing.  The bucket list is cut into world * SHARD_CHUNKS contiguous
    // chunks of equal prefix cost and the chunks are dealt to the ranks in turn: what the pair count cannot see — how many
    // CDR3 survivors a bucket sends to the full scoring, which varies tenfold along the length axis (one contiguous eighth
    // of the 1.6 M-cell repertoire held 55 % of all scored pairs) — is then spread over every rank.
    int per_rank = SHARD_CHUNKS;
    if (const char *e = getenv("EJOIN_SHARD_CHUNKS")) per_rank = std::max(1, std::min(EJOIN_MAX_RANGES, atoi(e)));
    const int chunks = world * per_rank;
    double total = 0;
    for (uint32_t b = 0; b < nb; b++) { double m = bs[b + 1] - bs[b]; total += m * (m - 1) / 2 + 64.0 * m; }
    double acc = 0;
    int ch = 0;
    for (uint32_t b = 0; b < nb; b++) {
        const double m = bs[b + 1] - bs[b], cost = m * (m - 1) / 2 + 64.0 * m;
        // move to the next chunk when this bucket's midpoint passes the chunk's boundary
        while (ch < chunks - 1 && acc + cost / 2 > total * (ch + 1) / chunks) ch++;
        owner[b] = (uint32_t)(ch % world);
        acc += cost;
    }
}

extern "C" int ejoin_plan_shards(const ejoin_pack_t *pk, int world, uint32_t *owner, uint32_t *bstart, uint32_t cap, uint32_t *nb_out) {
    if (!pk || world < 1 || !nb_out) return EJOIN_E_ARG;
    if (pk->abi_version != EJOIN_ABI_VERSION || (pk->n_info && (!pk->len[0] || !pk->len[1]))) return EJOIN_E_ARG;
    std::vector<uint32_t> bs, own;
    plan_buckets(pk, world, bs, own);
    const uint32_t nb = (uint32_t)own.size();
    *nb_out = nb;
    if (!owner || !bstart) return 0;
    if (cap < nb) return EJOIN_E_ARG;
    memcpy(owner, own.data(), sizeof(uint32_t) * nb);
    memcpy(bstart, bs.data(), sizeof(uint32_t) * (nb + 1));
    return 0;
}

// this rank's share: a few ascending entry ranges, or (stride mode) everything and every world-th tile row
static bool shard_plan(const ejoin_pack_t *pk, int rank, int world, Ranges &out) {
    std::vector<uint32_t> owner, bstart;
    plan_buckets(pk, world, bstart, owner);
    const uint32_t nb = (uint32_t)owner.size();
    // giant-bucket case: if one bucket holds most of the work, every rank takes all entries and
    // an interleaved share of the tile rows instead (stride mode)
    double total = 0, biggest = 0;
    for (uint32_t b = 0; b < nb; b++) { double m = bstart[b + 1] - bstart[b]; double c = m * (m - 1) / 2; total += c; biggest = std::max(biggest, c); }
    const bool stride_mode = world > 1 && biggest > 1.25 * total / world;
    out.clear();
    if (stride_mode || world <= 1) { out.emplace_back(0u, pk->n_info); return stride_mode; }
    for (uint32_t b = 0; b < nb; b++)
        if ((int)owner[b] == rank) {
            if (!out.empty() && out.back().second == bstart[b]) out.back().second = bstart[b + 1];
            else out.emplace_back(bstart[b], bstart[b + 1]);
        }
    return stride_mode;
}

static int run_shard_core(ejoin_handle *h, const ejoin_pack_t *pk, int rank, int world, int want_final, int *mode_out, ejoin_result_t *final,
                          ejoin_raw_edge_t **edges, uint64_t *n_edges, ejoin_stats_t *stats) {
    if (edges) *edges = nullptr;
    if (n_edges) *n_edges = 0;
    memset(&h->stats, 0, sizeof(h->stats));
    h->launches = 0;
    const double t0 = now_ms();
    Ranges rg;
    const bool stride_mode = shard_plan(pk, rank, world, rg);
    if (mode_out) *mode_out = stride_mode ? 1 : 0;
    int rc = upload_ranges(h, pk, rg);
    if (rc) return rc;
    h->is_bcr = (int)pk->is_bcr;
    const double t1 = now_ms();
    h->last_stride = stride_mode ? (uint32_t)world : 1u; h->last_phase = stride_mode ? (uint32_t)rank : 0u;
    rc = device_pipeline(h, (int)pk->is_bcr, h->last_stride, h->last_phase);
    if (rc) return rc;
    const double t2 = now_ms();
    const size_t nk = (size_t)h->counters.n_kept;
    size_t d2h = sizeof(DevCounters);
    if (want_final && !stride_mode) {
        // the shard's buckets are complete on this device: emitted joins with payload, no host replay
        rc = ensure_pin(h, sizeof(ejoin_edge_t) * (nk + 1) + 4 * ((size_t)h->n + 1) + 1024);
        if (rc) return rc;
        static const std::vector<uint2> none;
        rc = emit_result(h, pk, none, h->kept_keys, nk, 0, final);
        if (rc) return rc;
        d2h += h->stats.d2h_bytes;
    } else {
        rc = ensure_pin(h, 16 * nk + 64);
        if (rc) return rc;
        unsigned long long *hk = (unsigned long long *)h->pin, *hv = hk + nk;
        if (nk) {
            CK(cudaMemcpyAsync(hk, h->kept_keys, 8 * nk, cudaMemcpyDeviceToHost, h->st));
            CK(cudaMemcpyAsync(hv, h->kept_vals, 8 * nk, cudaMemcpyDeviceToHost, h->st));
            CK(cudaStreamSynchronize(h->st));
        }
        ejoin_raw_edge_t *e = (ejoin_raw_edge_t *)malloc(sizeof(ejoin_raw_edge_t) * (nk ? nk : 1));
        for (size_t i = 0; i < nk; i++) {
            e[i].k1 = host_to_global(h->rt, (uint32_t)(hk[i] >> 32)); e[i].k2 = host_to_global(h->rt, (uint32_t)hk[i]);
            e[i].cd = (uint16_t)(hv[i] & 0xFFFF); e[i].d = (uint16_t)((hv[i] >> 16) & 0xFFFF); e[i].flags = (uint32_t)(hv[i] >> 32);
        }
        if (edges) *edges = e; else free(e);
        if (n_edges) *n_edges = nk;
        d2h += 16 * nk;
    }
    const double t3 = now_ms();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        fill_stats(h, stats);
        (void)t1;
        stats->ms_h2d = h->stats.ms_h2d; stats->ms_d2h = 0; stats->ms_host_post = t3 - t2; stats->ms_total = t3 - t0; stats->d2h_bytes = d2h;
        stats->gpu_launches = h->launches;
    }
    return 0;
}

extern "C" int ejoin_run_shard(ejoin_handle_t *h, const ejoin_pack_t *pk, int rank, int world, ejoin_raw_edge_t **edges, uint64_t *n_edges,
                               ejoin_stats_t *stats) {
    if (!h || !pk || !edges || !n_edges || rank < 0 || rank >= world) return EJOIN_E_ARG;
    return run_shard_core(h, pk, rank, world, 0, nullptr, nullptr, edges, n_edges, stats);
}

extern "C" int ejoin_run_shard_final(ejoin_handle_t *h, const ejoin_pack_t *pk, int rank, int world, int *mode, ejoin_result_t *final,
                                     ejoin_raw_edge_t **raw, uint64_t *n_raw, ejoin_stats_t *stats) {
    if (!h || !pk || !mode || !final || !raw || !n_raw || rank < 0 || rank >= world) return EJOIN_E_ARG;
    memset(final, 0, sizeof(*final));
    return run_shard_core(h, pk, rank, world, 1, mode, final, raw, n_raw, stats);
}

extern "C" void ejoin_raw_free(ejoin_raw_edge_t *e) { free(e); }

// ---------------------------------------------------------------------------------------------
// The exchange step inside the library: every rank scores its bucket ranges and builds the
// PotentialJoin records of its emitted joins on its own GPU; the records travel device to device to rank 0
// (ncclSend / ncclRecv in one group, straight out of the work arena); rank 0 runs finish_join (GPU union-find
// over all info entries) on the gathered buffer and makes the one device-to-host copy of the result.
// ---------------------------------------------------------------------------------------------
#define NCK(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + nc->GetErrorString(r_);                           \
            return EJOIN_E_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

extern "C" int ejoin_comm_id(uint8_t *id) {
    NcclApi *nc = nccl_api();
    if (!nc || !id) return nc ? EJOIN_E_ARG : EJOIN_E_UNSUPPORTED;
    ncclUniqueId u;
    if (nc->GetUniqueId(&u) != ncclSuccess) return EJOIN_E_CUDA;
    static_assert(sizeof(u) == EJOIN_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof(u));
    return EJOIN_OK;
}

extern "C" int ejoin_comm_init(ejoin_handle_t *h, const uint8_t *id, int rank, int world) {
    if (!h || !id || world < 1 || rank < 0 || rank >= world) return EJOIN_E_ARG;
    NcclApi *nc = nccl_api();
    if (!nc) return fail(h, EJOIN_E_UNSUPPORTED, "libnccl.so.2 not found");
    CK(cudaSetDevice(h->device));
    if (h->comm) { if (nc->CommAbort) nc->CommAbort(h->comm); else nc->CommDestroy(h->comm); h->comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NCK(nc->CommInitRank(&h->comm, world, u, rank));
    h->comm_rank = rank; h->comm_world = world;
    CK(h->d_xcounts.ensure(8 * (size_t)(world + 1)));
    return EJOIN_OK;
}

extern "C" int ejoin_run_sharded(ejoin_handle_t *h, const ejoin_pack_t *pk, ejoin_result_t *out) {
    if (!h || !pk || !out) return EJOIN_E_ARG;
    memset(out, 0, sizeof(*out));
    if (!h->comm) return fail(h, EJOIN_E_ARG, "ejoin_run_sharded: ejoin_comm_init has not been called on this handle");
    NcclApi *nc = nccl_api();
    const int rank = h->comm_rank, world = h->comm_world, root = 0;
    memset(&h->stats, 0, sizeof(h->stats));
    h->launches = 0;
    const double t0 = now_ms();
    Ranges rg;
    const bool stride_mode = shard_plan(pk, rank, world, rg);
    int rc = upload_ranges(h, pk, rg);
    if (rc) return rc;
    const uint32_t n_all = pk->n_info;
    const bool whole = rg.size() == 1 && rg[0].first == 0 && rg[0].second == n_all;
    if (rank == root && !whole && n_all) {
        // finish_join runs over every info entry: rank 0 also takes the whole exact_id array (4 bytes per entry)
        CK(h->d_exact_all.ensure(4 * (size_t)n_all));
        CK(cudaMemcpyAsync(h->d_exact_all.p, pk->exact_id, 4 * (size_t)n_all, cudaMemcpyHostToDevice, h->st_copy));
        CK(cudaEventRecord(h->ev[6], h->st_copy));          // ev[6] is re-recorded by the pipeline; the wait below precedes that
        h->h2d_bytes += 4 * (size_t)n_all;
    }
    h->is_bcr = (int)pk->is_bcr;
    h->last_stride = stride_mode ? (uint32_t)world : 1u; h->last_phase = stride_mode ? (uint32_t)rank : 0u;
    if (rank == root && !whole && n_all) CK(cudaStreamWaitEvent(h->st, h->ev[6], 0));
    rc = device_pipeline(h, (int)pk->is_bcr, h->last_stride, h->last_phase);
    if (rc) return rc;
    const double t1 = now_ms();
    const size_t nk = (size_t)h->counters.n_kept;
    DevCounters *ctr = (DevCounters *)h->d_counters.p;
    cudaStream_t st = h->st;
    if (h->n == 0) CK(cudaMemsetAsync(ctr, 0, sizeof(DevCounters), st));      // a rank without entries ran no pipeline: its device counters are stale
    std::vector<unsigned long long> counts(world, 0);
    auto exchange_counts = [&]() -> int {
        unsigned long long *dc = (unsigned long long *)h->d_xcounts.p;
        NCK(nc->AllGather(&ctr->n_kept, dc, 1, ncclUint64, h->comm, st));
        CK(cudaMemcpyAsync(counts.data(), dc, 8 * (size_t)world, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return 0;
    };
    h->stats.d2h_bytes = sizeof(DevCounters);
    if (stride_mode) {
        // a giant bucket shared by the ranks: the accepted edges (16 bytes each) go to rank 0, which replays them
        rc = exchange_counts();
        if (rc) return rc;
        size_t total = 0, off_me = 0;
        for (int r = 0; r < world; r++) { if (r < rank) off_me += counts[r]; total += counts[r]; }
        (void)off_me;
        unsigned long long *gk = nullptr, *gv = nullptr, *sk_in = nullptr, *sk_out = nullptr, *sv_in = nullptr, *sv_out = nullptr;
        void *gacc = nullptr, *g_cub = nullptr;
        uint2 *gends = nullptr;
        size_t g_cub_bytes = 0;
        if (rank == root) {
            const size_t te = total + 1;
            cub::DeviceRadixSort::SortPairs(nullptr, g_cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                            (unsigned long long *)nullptr, (int)te);
            CK(h->d_gather.ensure(6 * 8 * te + sizeof(ejoin_raw_edge_t) * te + 8 * te + g_cub_bytes + 4096));
            char *b = (char *)h->d_gather.p;
            auto carve = [&](size_t bytes) { char *r = b; b += (bytes + 255) & ~(size_t)255; return r; };
            gk = (unsigned long long *)carve(8 * te); gv = (unsigned long long *)carve(8 * te);
            sk_in = (unsigned long long *)carve(8 * te); sk_out = (unsigned long long *)carve(8 * te);
            sv_in = (unsigned long long *)carve(8 * te); sv_out = (unsigned long long *)carve(8 * te);
            gacc = carve(sizeof(ejoin_raw_edge_t) * te); gends = (uint2 *)carve(8 * te);
            g_cub = carve(g_cub_bytes);
        }
        NCK(nc->GroupStart());
        if (rank == root) {
            size_t off = 0;
            for (int r = 0; r < world; r++) {
                if (r != root && counts[r]) { NCK(nc->Recv(gk + off, counts[r], ncclUint64, r, h->comm, st)); NCK(nc->Recv(gv + off, counts[r], ncclUint64, r, h->comm, st)); }
                off += counts[r];
            }
        } else if (nk) {
            NCK(nc->Send(h->kept_keys, nk, ncclUint64, root, h->comm, st));
            NCK(nc->Send(h->kept_vals, nk, ncclUint64, root, h->comm, st));
        }
        NCK(nc->GroupEnd());
        if (rank == root) {
            size_t off = 0;
            for (int r = 0; r < root; r++) off += counts[r];
            if (nk) {
                CK(cudaMemcpyAsync(gk + off, h->kept_keys, 8 * nk, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync(gv + off, h->kept_vals, 8 * nk, cudaMemcpyDeviceToDevice, st));
            }
            // The forest over the gathered edges, on this GPU (every entry is resident here in stride mode): Boruvka from singletons —
            // the gathered set holds every rank's exact tree edges plus the accepted edges it could not prove redundant, so its
            // minimum spanning forest under the (k1,k2) order is join_core's `pot` (forest) —, then the two-cell rule,
            // the sort, the payload and finish_join exactly as on one GPU.  No edge visits the host.
            const uint32_t n = h->n;
            const unsigned nb256 = (n + 255) / 256;
            ejoin_raw_edge_t *acc = (ejoin_raw_edge_t *)gacc;
            if (total) k_kv_to_acc<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gk, gv, total, acc);
            k_iota_parent<<<nb256, 256, 0, st>>>(h->w_parent, h->w_f0_key, n);
            k_set_accept<<<1, 1, 0, st>>>(ctr, total);
            k_fill_u64<<<2 * nb256, 256, 0, st>>>(h->w_best, ~0ull, 2 * (size_t)n);
            {
                if (h->bor_grid == 0) { int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_boruvka, 256, 0); h->bor_grid = occ >= 1 ? h->sm_count : 1; }
                ejoin_raw_edge_t *a_acc = acc; DevCounters *a_ctr = ctr; unsigned long long a_cap = total; const uint32_t *a_kts = h->k_to_s;
                uint32_t *a_par = h->w_parent; unsigned long long *a_best = h->w_best; uint2 *a_ends = gends; uint32_t a_n = n;
                void *args[] = { &a_acc, &a_ctr, &a_cap, &a_kts, &a_par, &a_best, &a_ends, &a_n };
                CK(cudaLaunchCooperativeKernel((void *)k_boruvka, dim3(h->bor_grid), dim3(256), args, 0, st));
            }
            CK(cudaMemsetAsync(h->w_comp_size, 0, 4 * (size_t)n, st));
            k_comp_size<<<nb256, 256, 0, st>>>(h->w_parent, (const DevCounters *)h->d_counters.p, h->w_comp_size);
            k_final_collect<<<h->sm_count * 8, 256, 0, st>>>(h->din, h->w_f0_key, h->w_f0_val, n, acc, ctr, total, h->k_to_s, 0, h->w_parent, h->w_comp_size,
                                                             h->params.easy, sk_in, sv_in);
            h->launches += 7;
            CK(cudaGetLastError());
            unsigned long long nfin = 0;
            CK(cudaMemcpyAsync(&nfin, &ctr->n_kept, sizeof(nfin), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (nfin) {
                int kb = 1;
                while (kb < 32 && (1ull << kb) < (unsigned long long)n) kb++;
                size_t tb = g_cub_bytes;
                cub::DeviceRadixSort::SortPairs(g_cub, tb, sk_in, sk_out, sv_in, sv_out, (int)nfin, 0, 32 + kb, st);
            }
            rc = ensure_pin(h, sizeof(ejoin_edge_t) * (nfin + 1) + 4 * ((size_t)n + 1) + 1024);
            if (rc) return rc;
            static const std::vector<uint2> none;
            rc = emit_result(h, pk, none, sk_out, (size_t)nfin, 0, out);
            if (rc) return rc;
        } else CK(cudaStreamSynchronize(st));
    } else {
        // contiguous bucket ranges: the rank's emitted joins are final; build their records here, ship them to rank 0
        Arena &A = h->work;
        const size_t save = A.off;
        uint2 *d_pairs = nullptr;
        ejoin_edge_t *d_out = nullptr;
        uint32_t *d_slow = nullptr;
        if (nk) {
            d_pairs = A.take<uint2>(nk); d_out = A.take<ejoin_edge_t>(nk); d_slow = A.take<uint32_t>(nk);
            if (A.off > A.buf.cap) { A.off = save; return fail(h, EJOIN_E_NOMEM, "work arena too small for the epilogue"); }
        }
        if (nk) {
            const DevParams pr = dev_params(h->params, (int)pk->is_bcr);
            DevTables tbs;
            tbs.meta = h->meta; tbs.c3 = h->c3; tbs.t2 = h->t2; tbs.p1tab = (const double *const *)h->d_p1ptrs.p;
            tbs.multinfo = (const MultInfo *)h->d_multinfo.p; tbs.multpool = (const double *)h->d_multpool.p; tbs.cdmax = (const uint16_t *)h->d_cdmax.p;
            k_keys_to_pairs<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(h->kept_keys, (uint32_t)nk, h->rt, d_pairs);
            CK(cudaMemsetAsync(&ctr->n_payload_slow, 0, sizeof(unsigned int), st));
            k_payload_fast<<<h->sm_count * 8, 128, 0, st>>>(h->din, tbs, pr, ctr, h->k_to_s, d_pairs, (uint32_t)nk, h->rt, h->n, d_out, d_slow);
            k_payload<<<h->sm_count * 4, 128, 0, st>>>(h->din, tbs, pr, ctr, h->k_to_s, d_pairs, d_slow, h->rt, h->n, d_out);
            h->launches += 3;
            CK(cudaGetLastError());
        }
        rc = exchange_counts();
        if (rc) { A.off = save; return rc; }
        size_t total = 0;
        for (int r = 0; r < world; r++) total += counts[r];
        ejoin_edge_t *g_edges = nullptr, *g_sorted = nullptr;
        unsigned long long *sk_in = nullptr, *sk_out = nullptr;
        uint32_t *si_in = nullptr, *si_out = nullptr;
        void *g_cub = nullptr;
        size_t g_cub_bytes = 0;
        char *g_rest = nullptr;
        if (rank == root) {
            cub::DeviceRadixSort::SortPairs(nullptr, g_cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (int)std::max<size_t>(total, 1));
            const size_t te = total + 1;
            CK(h->d_gather.ensure(2 * sizeof(ejoin_edge_t) * te + 2 * 8 * te + 2 * 4 * te + g_cub_bytes + 8 * te + 8 * (size_t)n_all +
                                  4 * (size_t)std::max(1u, pk->n_exact) + 4096));
            char *b = (char *)h->d_gather.p;
            auto carve = [&](size_t bytes) { char *r = b; b += (bytes + 255) & ~(size_t)255; return r; };
            g_edges = (ejoin_edge_t *)carve(sizeof(ejoin_edge_t) * te); g_sorted = (ejoin_edge_t *)carve(sizeof(ejoin_edge_t) * te);
            sk_in = (unsigned long long *)carve(8 * te); sk_out = (unsigned long long *)carve(8 * te);
            si_in = (uint32_t *)carve(4 * te); si_out = (uint32_t *)carve(4 * te);
            g_cub = carve(g_cub_bytes);
            g_rest = b;
        }
        NCK(nc->GroupStart());
        if (rank == root) {
            size_t off = 0;
            for (int r = 0; r < world; r++) {
                if (r != root && counts[r]) NCK(nc->Recv(g_edges + off, counts[r] * sizeof(ejoin_edge_t), ncclUint8, r, h->comm, st));
                off += counts[r];
            }
        } else if (nk) NCK(nc->Send(d_out, nk * sizeof(ejoin_edge_t), ncclUint8, root, h->comm, st));
        NCK(nc->GroupEnd());
        if (rank == root) {
            size_t off = 0;
            for (int r = 0; r < root; r++) off += counts[r];
            if (nk) CK(cudaMemcpyAsync(g_edges + off, d_out, nk * sizeof(ejoin_edge_t), cudaMemcpyDeviceToDevice, st));
            if (total && world > 1) {
                // every rank's list is in raw_joins order; the ranks' bucket ranges interleave, so one sort by (k1,k2) orders the union
                int kb = 1;
                while (kb < 32 && (1ull << kb) < (unsigned long long)n_all) kb++;
                k_edge_keys<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g_edges, (uint32_t)total, sk_in, si_in);
                size_t tb = g_cub_bytes;
                cub::DeviceRadixSort::SortPairs(g_cub, tb, sk_in, sk_out, si_in, si_out, (int)total, 0, 32 + kb, st);
                const unsigned long long nthreads = (unsigned long long)total * (sizeof(ejoin_edge_t) / 8);
                k_gather_edges<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(g_edges, si_out, (uint32_t)total, g_sorted);
                h->launches += 2;
                g_edges = g_sorted;
            }
            // finish_join on the gathered records
            char *base = g_rest;
            base = (char *)(((uintptr_t)base + 15) & ~(uintptr_t)15);
            uint2 *g_pairs = (uint2 *)base; base += 8 * (total + 1);
            uint32_t *d_label = (uint32_t *)base; base += 4 * (size_t)n_all;
            uint32_t *d_label_out = (uint32_t *)base; base += 4 * (size_t)n_all;
            uint32_t *d_first = (uint32_t *)base;
            const uint32_t *d_exact = whole ? h->din.exact_id : (const uint32_t *)h->d_exact_all.p;
            rc = ensure_pin(h, sizeof(ejoin_edge_t) * (total + 1) + 4 * ((size_t)n_all + 1) + 1024);
            if (rc) { A.off = save; return rc; }
            ejoin_edge_t *h_edges = (ejoin_edge_t *)h->pin;
            uint32_t *h_labels = (uint32_t *)((char *)h->pin + ((sizeof(ejoin_edge_t) * (total + 1) + 255) & ~(size_t)255));
            if (total) {
                k_edge_pairs<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g_edges, (uint32_t)total, g_pairs);
                CK(cudaMemcpyAsync(h_edges, g_edges, total * sizeof(ejoin_edge_t), cudaMemcpyDeviceToHost, st));
            }
            if (n_all) {
                const uint32_t nx = pk->n_exact;
                CK(cudaMemsetAsync(&ctr->n_classes, 0, sizeof(unsigned long long), st));
                k_uf_init<<<(std::max(n_all, nx) + 255) / 256, 256, 0, st>>>(d_label, n_all, d_first, nx);
                k_uf_first<<<(n_all + 255) / 256, 256, 0, st>>>(d_exact, n_all, nx, d_first);
                k_uf_union<<<(std::max<uint32_t>(n_all, (uint32_t)total) + 255) / 256, 256, 0, st>>>(d_label, g_pairs, (uint32_t)total, 0, d_exact, n_all, nx, d_first);
                k_uf_flatten<<<(n_all + 255) / 256, 256, 0, st>>>(d_label, n_all, ctr, d_label_out);
                h->launches += 5;
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(h_labels, d_label_out, 4 * (size_t)n_all, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(&h->counters.n_classes, &ctr->n_classes, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            }
            CK(cudaStreamSynchronize(st));
            out->n_edges = total; out->edges = h_edges; out->owner = h;
            out->class_of = n_all ? h_labels : nullptr; out->n_info = n_all; out->n_classes = (uint32_t)h->counters.n_classes;
            h->stats.d2h_bytes += total * sizeof(ejoin_edge_t) + 4 * (size_t)n_all;
        } else {
            CK(cudaStreamSynchronize(st));
            out->owner = h;
        }
        A.off = save;
    }
    const double t2 = now_ms();
    fill_stats(h, &out->stats);
    out->stats.ms_h2d = h->stats.ms_h2d; out->stats.ms_d2h = 0; out->stats.ms_host_post = t2 - t1; out->stats.ms_total = t2 - t0;
    out->stats.d2h_bytes = h->stats.d2h_bytes;
    out->stats.gpu_launches = h->launches;
    return 0;
}

// Ordered replay + payload + partition from gathered raw edges.  The handle must still hold
// the upload that covers every entry the edges touch (its own shard, or everything in
// stride mode).
extern "C" int ejoin_finish(ejoin_handle_t *h, const ejoin_pack_t *pk, const ejoin_raw_edge_t *edges, uint64_t n_edges, ejoin_result_t *out) {
    if (!h || !pk || !out || (n_edges && !edges)) return EJOIN_E_ARG;
    memset(out, 0, sizeof(*out));
    {
        int rc0 = check_pack(h, pk);
        if (rc0) return rc0;
        if (!h->uploaded && n_edges) return fail(h, EJOIN_E_ARG, "ejoin_finish: the handle holds no upload");
        // every endpoint must lie in the range this handle holds ([k_base, k_base + n): everything in stride mode,
        // the rank's own buckets otherwise): the payload kernels index the resident rows by k - k_base
        for (uint64_t i = 0; i < n_edges; i++) {
            uint32_t l1, l2;
            if (edges[i].k1 >= edges[i].k2 || edges[i].k2 >= pk->n_info || !host_to_local(h->rt, h->n, edges[i].k1, &l1) || !host_to_local(h->rt, h->n, edges[i].k2, &l2))
                return fail(h, EJOIN_E_ARG, "ejoin_finish: an edge endpoint lies outside the entries this handle holds "
                                            "(gather-then-finish needs stride mode or a handle holding the whole input)");
        }
    }
    std::vector<std::pair<unsigned long long, unsigned long long>> kv(n_edges);
    for (uint64_t i = 0; i < n_edges; i++) {
        kv[i].first = ((unsigned long long)edges[i].k1 << 32) | edges[i].k2;
        kv[i].second = ((unsigned long long)edges[i].flags << 32) | ((unsigned long long)edges[i].d << 16) | edges[i].cd;
    }
    std::sort(kv.begin(), kv.end());
    kv.erase(std::unique(kv.begin(), kv.end(), [](auto &a, auto &b) { return a.first == b.first; }), kv.end());
    const size_t nk = kv.size();
    int rc = ensure_pin(h, 16 * nk + sizeof(ejoin_edge_t) * (nk + 1) + 4 * ((size_t)h->n + 1) + 1024);
    if (rc) return rc;
    unsigned long long *keys = (unsigned long long *)h->pin, *vals = keys + nk;
    for (size_t i = 0; i < nk; i++) { keys[i] = kv[i].first; vals[i] = kv[i].second; }
    rc = finish_common(h, pk, keys, vals, nk, 16 * nk, out);
    if (rc) return rc;
    fill_stats(h, &out->stats);
    out->stats.ms_host_post = h->stats.ms_host_post;
    return 0;
}

extern "C" int ejoin_partition(const ejoin_pack_t *pk, const uint32_t *k1, const uint32_t *k2, uint64_t ne, uint32_t *class_of, uint32_t *n_classes) {
    if (!pk || !class_of || (ne && (!k1 || !k2))) return EJOIN_E_ARG;
    for (uint64_t i = 0; i < ne; i++) if (k1[i] >= pk->n_info || k2[i] >= pk->n_info) return EJOIN_E_ARG;
    host_partition(pk, k1, k2, ne, class_of, n_classes);
    return EJOIN_OK;
}

extern "C" int ejoin_partition_gpu(ejoin_handle_t *h, const ejoin_pack_t *pk, const uint32_t *k1, const uint32_t *k2, uint64_t ne, uint32_t *class_of,
                                   uint32_t *n_classes) {
    if (!h || !pk || !class_of || (ne && (!k1 || !k2))) return EJOIN_E_ARG;
    CK(cudaSetDevice(h->device));
    const uint32_t n = pk->n_info, nx = pk->n_exact;
    if (n == 0) { if (n_classes) *n_classes = 0; return EJOIN_OK; }
    // its own small buffer: the work arena may hold a resident shard
    CK(h->d_part.ensure(4 * (size_t)n * 3 + 4 * (size_t)std::max(1u, nx) + 8 * (ne + 1) + 1024));
    char *base = (char *)h->d_part.p;
    uint32_t *d_exact = (uint32_t *)base; base += 4 * (size_t)n;
    uint32_t *d_label = (uint32_t *)base; base += 4 * (size_t)n;
    uint32_t *d_label_out = (uint32_t *)base; base += 4 * (size_t)n;
    uint32_t *d_first = (uint32_t *)base; base += 4 * (size_t)std::max(1u, nx);
    base = (char *)(((uintptr_t)base + 15) & ~(uintptr_t)15);
    uint2 *d_pairs = (uint2 *)base;
    // its own pinned staging: h->pin may hold a live result (one live result per handle)
    const size_t need2 = ((8 * (ne + 1) + 255) & ~(size_t)255) + 4 * (size_t)n + 1024;
    if (need2 > h->pin2_cap) {
        if (h->pin2) cudaFreeHost(h->pin2);
        h->pin2 = nullptr; h->pin2_cap = 0;
        CK(cudaMallocHost(&h->pin2, need2 + need2 / 2));
        h->pin2_cap = need2 + need2 / 2;
    }
    uint2 *hp = (uint2 *)h->pin2;
    for (uint64_t i = 0; i < ne; i++) { if (k1[i] >= n || k2[i] >= n) return fail(h, EJOIN_E_ARG, "edge endpoint out of range"); hp[i] = make_uint2(k1[i], k2[i]); }
    uint32_t *hl = (uint32_t *)((char *)h->pin2 + ((8 * (ne + 1) + 255) & ~(size_t)255));
    DevCounters *ctr = (DevCounters *)h->d_counters.p;
    CK(cudaMemcpyAsync(d_exact, pk->exact_id, 4 * (size_t)n, cudaMemcpyHostToDevice, h->st));
    if (ne) CK(cudaMemcpyAsync(d_pairs, hp, 8 * ne, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemsetAsync(&ctr->n_classes, 0, sizeof(unsigned long long), h->st));
    k_uf_init<<<(std::max(n, nx) + 255) / 256, 256, 0, h->st>>>(d_label, n, d_first, nx);
    k_uf_first<<<(n + 255) / 256, 256, 0, h->st>>>(d_exact, n, nx, d_first);
    k_uf_union<<<(std::max<uint32_t>(n, (uint32_t)ne) + 255) / 256, 256, 0, h->st>>>(d_label, d_pairs, (uint32_t)ne, 0, d_exact, n, nx, d_first);
    k_uf_flatten<<<(n + 255) / 256, 256, 0, h->st>>>(d_label, n, ctr, d_label_out);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hl, d_label_out, 4 * (size_t)n, cudaMemcpyDeviceToHost, h->st));
    unsigned long long ncls = 0;
    CK(cudaMemcpyAsync(&ncls, &ctr->n_classes, sizeof(ncls), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    memcpy(class_of, hl, 4 * (size_t)n);
    if (n_classes) *n_classes = (uint32_t)ncls;
    return EJOIN_OK;
}

// ---------------------------------------------------------------------------------------------
// ejoin_pack_compact: ASCII / parallel-array pack -> compact pack (2-bit V..J bases in the DnaString
// layout, one 64-byte record per entry), in pinned host memory.  Host code only; a caller that holds
// CloneInfo.tigsp writes this form directly.
// ---------------------------------------------------------------------------------------------
#include <thread>
namespace {
struct CompactOwner {
    ejoin_pack_t pk;                       // first member: the pointer handed to the caller
    uint64_t *tig2 = nullptr; ejoin_entry_rec_t *rec = nullptr; uint8_t *ascii = nullptr;
    uint64_t *c3w = nullptr; uint32_t *c3off = nullptr;
    bool pinned = true;
};
// pinned when a CUDA device is there; plain aligned memory otherwise (the conversion itself is host code and is tested
// on machines without a GPU; the join entry points still refuse to run there)
void *compact_alloc(CompactOwner *o, size_t bytes) {
    void *p = nullptr;
    if (o->pinned) {
        if (cudaMallocHost(&p, bytes) == cudaSuccess) return p;
        cudaGetLastError();
        if (o->tig2) return nullptr;       // do not mix the two kinds
        o->pinned = false;
    }
    if (posix_memalign(&p, 4096, (bytes + 4095) & ~(size_t)4095) != 0) return nullptr;
    return p;
}
struct Code256 { uint8_t t[256]; Code256() { memset(t, 0xFF, 256); t['A'] = 0; t['C'] = 1; t['G'] = 2; t['T'] = 3; } };
const Code256 g_code;
// DnaString word -> (high-bit plane << 32) | low-bit plane, base i at bit 31-i of each half (EJOIN_F_TIGS_PLANES)
inline uint64_t even_bits32(uint64_t v) {          // bits 0,2,4,..,62 of v -> bits 0..31
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    return (v | (v >> 16)) & 0xFFFFFFFFull;
}
inline uint64_t planes_of_word(uint64_t w) { return (even_bits32(w >> 1) << 32) | even_bits32(w); }

template <typename F> void parallel_ranges(uint32_t n, int threads, F f) {
    threads = std::max(1, std::min(threads, (int)((n + 4095) / 4096)));
    if (threads == 1) { f(0u, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        const uint32_t a = (uint32_t)((uint64_t)n * t / threads), b = (uint32_t)((uint64_t)n * (t + 1) / threads);
        th.emplace_back([=] { f(a, b); });
    }
    for (auto &x : th) x.join();
}
}  // namespace

extern "C" int ejoin_pack_compact(const ejoin_pack_t *in, int threads, ejoin_pack_t **out) {
    if (!in || !out) return EJOIN_E_ARG;
    *out = nullptr;
    if (in->abi_version != EJOIN_ABI_VERSION || (in->flags & (EJOIN_F_TIGS_2BIT | EJOIN_F_ENTRY_RECS | EJOIN_F_CDR3_2BIT))) return EJOIN_E_ARG;
    if (pack_null_field(in)) return EJOIN_E_ARG;
    const uint32_t n = in->n_info;
    if (threads < 1) threads = 1;
    const bool as_planes = getenv("EJOIN_COMPACT_DNASTRING") == nullptr;      // tests: keep the words exactly as a DnaString holds them
    // pass 1: which chains stay ASCII (deletion marker or a byte outside ACGT), sizes
    std::vector<uint8_t> asc[2];
    for (int c = 0; c < 2; c++) asc[c].assign(n, 0);
    bool id_overflow = false;
    parallel_ranges(n, threads, [&](uint32_t a, uint32_t b) {
        bool ov = false;
        for (uint32_t k = a; k < b; k++)
            for (int c = 0; c < 2; c++) {
                bool x = in->has_del[c][k] != 0;
                if (!x) {
                    const uint8_t *p = in->tig_arena + in->tig_off[c][k];
                    uint8_t bad = 0;
                    for (uint32_t i = 0, L = in->len[c][k]; i < L; i++) bad |= g_code.t[p[i]];
                    x = (bad & 0x80) != 0;
                }
                asc[c][k] = x;
                auto big = [](uint32_t v) { return v != EJOIN_NONE && v >= 0xFFFFu; };
                ov |= big(in->v_ref_id[c][k]) || big(in->j_ref_id[c][k]) || big(in->dref_id[c][k]) || big(in->vs_seq[c][k]) || big(in->js_seq[c][k]) ||
                      big(in->vname_id[c][k]) || big(in->vuni_seq[c][k]) || big(in->utr_seq[c][k]) || (c == 0 && big(in->c_ref_id_light[k]));
            }
        if (ov) id_overflow = true;
    });
    if (id_overflow) return EJOIN_E_UNSUPPORTED;
    std::vector<uint64_t> woff[2], aoff[2];
    for (int c = 0; c < 2; c++) { woff[c].assign(n, 0); aoff[c].assign(n, 0); }
    uint64_t words = 0, abytes = 0;
    for (uint32_t k = 0; k < n; k++) {
        // every entry starts on a 64-byte boundary: SM reads of host memory move whole 64-byte pieces, and an entry read in place
        // then costs 192 bytes of PCIe traffic instead of up to 256 (measured: 41.7 against 32.6 GB/s of useful bytes)
        words = (words + 7) & ~7ull;
        for (int c = 0; c < 2; c++) {
            if (asc[c][k]) { aoff[c][k] = abytes; abytes += in->len[c][k]; }
            else { woff[c][k] = words; words += (in->len[c][k] + 31) / 32; }
        }
    }
    std::vector<uint32_t> c3off_v(n ? n : 1);
    uint64_t c3words = 0;
    for (uint32_t k = 0; k < n; k++) { c3off_v[k] = (uint32_t)c3words; c3words += ((uint64_t)in->cdr3_len[0][k] + in->cdr3_len[1][k] + 31) / 32; }
    if (words >= 0xFFFFFFFFull || abytes >= 0xFFFFFFFFull || c3words >= 0xFFFFFFFFull) return EJOIN_E_UNSUPPORTED;
    CompactOwner *o = new CompactOwner();
    o->pk = *in;
    o->tig2 = (uint64_t *)compact_alloc(o, 8 * (words + 16));
    if (o->tig2) o->rec = (ejoin_entry_rec_t *)compact_alloc(o, sizeof(ejoin_entry_rec_t) * ((size_t)n + 1));
    if (o->rec) o->ascii = (uint8_t *)compact_alloc(o, abytes + 128);
    if (o->ascii) o->c3w = (uint64_t *)compact_alloc(o, 8 * (c3words + 16));
    if (o->c3w) o->c3off = (uint32_t *)compact_alloc(o, 4 * ((size_t)n + 16));
    if (!o->tig2 || !o->rec || !o->ascii || !o->c3w || !o->c3off) { ejoin_pack_compact_free(&o->pk); return EJOIN_E_NOMEM; }
    memcpy(o->c3off, c3off_v.data(), 4 * (size_t)n);
    bool bad_cdr3 = false;
    memset(o->tig2, 0, 8 * (words + 16));                 // alignment gaps included
    memset(o->ascii + abytes, 0, 128);
    parallel_ranges(n, threads, [&](uint32_t a, uint32_t b) {
        for (uint32_t k = a; k < b; k++) {
            ejoin_entry_rec_t r;
            memset(&r, 0, sizeof(r));
            auto n16 = [](uint32_t v) { return v == EJOIN_NONE ? (uint16_t)0xFFFF : (uint16_t)v; };
            for (int c = 0; c < 2; c++) {
                const uint8_t *p = in->tig_arena + in->tig_off[c][k];
                const uint32_t L = in->len[c][k];
                if (asc[c][k]) {
                    memcpy(o->ascii + aoff[c][k], p, L);
                    r.tig_off[c] = (uint32_t)aoff[c][k]; r.has_del[c] = 1;
                } else {
                    uint64_t *w = o->tig2 + woff[c][k];
                    for (uint32_t i0 = 0; i0 < L; i0 += 32) {
                        uint64_t x = 0;
                        const uint32_t m = std::min(32u, L - i0);
                        for (uint32_t i = 0; i < m; i++) x |= (uint64_t)g_code.t[p[i0 + i]] << (62 - 2 * i);
                        w[i0 >> 5] = as_planes ? planes_of_word(x) : x;
                    }
                    r.tig_off[c] = (uint32_t)woff[c][k]; r.has_del[c] = 0;
                }
                r.cdr3_start[c] = in->cdr3_start[c][k];
                r.v_ref_id[c] = n16(in->v_ref_id[c][k]); r.j_ref_id[c] = n16(in->j_ref_id[c][k]); r.dref_id[c] = n16(in->dref_id[c][k]);
                r.vs_seq[c] = n16(in->vs_seq[c][k]); r.js_seq[c] = n16(in->js_seq[c][k]); r.vname_id[c] = n16(in->vname_id[c][k]);
                r.vuni_seq[c] = n16(in->vuni_seq[c][k]); r.utr_seq[c] = n16(in->utr_seq[c][k]);
            }
            {   // both CDR3s concatenated, 2 bits per base, DnaString bit order
                uint64_t *w = o->c3w + c3off_v[k];
                const uint32_t cl0 = in->cdr3_len[0][k], tot = cl0 + in->cdr3_len[1][k];
                const uint8_t *q0 = in->cdr3_arena + in->cdr3_off[0][k], *q1 = in->cdr3_arena + in->cdr3_off[1][k];
                uint8_t bad = 0;
                for (uint32_t i0 = 0; i0 < tot; i0 += 32) {
                    uint64_t x = 0;
                    const uint32_t m = std::min(32u, tot - i0);
                    for (uint32_t i = 0; i < m; i++) {
                        const uint32_t p = i0 + i;
                        const uint8_t code = g_code.t[p < cl0 ? q0[p] : q1[p - cl0]];
                        bad |= code;
                        x |= (uint64_t)(code & 3) << (62 - 2 * i);
                    }
                    w[i0 >> 5] = x;
                }
                if (bad & 0x80) bad_cdr3 = true;
            }
            r.c_ref_id_light = n16(in->c_ref_id_light[k]); r.hcomp = in->hcomp[k];
            r.fr1_start = in->fr1_start[k]; r.cdr1_start = in->cdr1_start[k]; r.fr2_start = in->fr2_start[k];
            r.cdr2_start = in->cdr2_start[k]; r.fr3_start = in->fr3_start[k];
            o->rec[k] = r;
        }
    });
    if (bad_cdr3) { ejoin_pack_compact_free(&o->pk); return EJOIN_E_UNSUPPORTED; }      // a CDR3 character outside ACGT (documented limit)
    ejoin_pack_t &pk = o->pk;
    pk.flags = (in->flags & ~EJOIN_F_TIGS_IN_PLACE) | EJOIN_F_TIGS_2BIT | (as_planes ? EJOIN_F_TIGS_PLANES : 0u) | EJOIN_F_ENTRY_RECS | EJOIN_F_CDR3_2BIT |
               (o->pinned ? EJOIN_F_TIGS_IN_PLACE : 0u);
    pk.cdr3_2bit = o->c3w; pk.cdr3_2bit_words = c3words; pk.cdr3_2bit_off = o->c3off;
    pk.cdr3_arena = nullptr; pk.cdr3_arena_bytes = 0; pk.cdr3_off[0] = pk.cdr3_off[1] = nullptr;
    pk.tig2_arena = o->tig2; pk.tig2_arena_words = words; pk.entry_rec = o->rec;
    pk.tig_arena = o->ascii; pk.tig_arena_bytes = abytes;
    for (int c = 0; c < 2; c++) {
        pk.tig_off[c] = nullptr; pk.has_del[c] = nullptr; pk.cdr3_start[c] = nullptr; pk.v_ref_id[c] = nullptr; pk.j_ref_id[c] = nullptr;
        pk.dref_id[c] = nullptr; pk.vs_seq[c] = nullptr; pk.js_seq[c] = nullptr; pk.vname_id[c] = nullptr; pk.vuni_seq[c] = nullptr; pk.utr_seq[c] = nullptr;
    }
    pk.c_ref_id_light = nullptr; pk.hcomp = nullptr;
    pk.fr1_start = pk.cdr1_start = pk.fr2_start = pk.cdr2_start = pk.fr3_start = nullptr;
    *out = &o->pk;
    return EJOIN_OK;
}

extern "C" void ejoin_pack_compact_free(ejoin_pack_t *p) {
    if (!p) return;
    CompactOwner *o = reinterpret_cast<CompactOwner *>(p);
    for (void *q : { (void *)o->tig2, (void *)o->rec, (void *)o->ascii, (void *)o->c3w, (void *)o->c3off })
        if (q) { if (o->pinned) cudaFreeHost(q); else free(q); }
    delete o;
}
