// libbioslam_b200: C ABI (include/bioslam_b200.h) of the B200-native Levenberg-Marquardt path.
// Host-side orchestration only; all arithmetic runs in the kernels of kernels_*.cuh.  There is NO CPU fallback: every
// entry point that computes launches CUDA kernels on the problem's device and reports CUDA errors as B200_CUDA_ERROR.
#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "kernels_preint.cuh"
#include "kernels_solve.cuh"
#include "kernels_local.cuh"
#include "kernels_sweep_pipe.cuh"
#include "kernels_jointangles.cuh"

#ifndef B200_USE_EMU
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the entry points are resolved with dlsym, so the library loads without NCCL
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges around linearize / solve levels / retract (no cost without a profiler)
struct NvtxRange { explicit NvtxRange(const char* n) { nvtxRangePushA(n); } ~NvtxRange() { nvtxRangePop(); } };
#define B200_NVTX_CAT2(a, b) a##b
#define B200_NVTX_CAT(a, b) B200_NVTX_CAT2(a, b)
#define B200_NVTX(name) NvtxRange B200_NVTX_CAT(nvtx_range_, __LINE__)(name)
#else
#define B200_NVTX(name)
#endif

using namespace b200d;

static thread_local std::string g_last_error;
static void set_err(const std::string& s) { g_last_error = s; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e__));                          \
            return B200_CUDA_ERROR;                                                                \
        }                                                                                          \
    } while (0)

static const int kValDim[4] = {12, 3, 6, 3};
static const int kTanDim[4] = {6, 3, 6, 2};

struct FInfo { int nvars; int vt[B200_MAX_FACTOR_VARS]; int nconst; };
static bool finfo(int type, FInfo& f) {
    const int P = B200_VAR_POSE3, V = B200_VAR_VEC3, B = B200_VAR_BIAS6, U = B200_VAR_UNIT3;
    switch (type) {
    case B200_F_IMU_COMBINED: f = {6, {P, V, P, V, B, B}, 190}; return true;
    case B200_F_IMU_STATICBIAS: f = {5, {P, V, P, V, B}, 115}; return true;
    case B200_F_ANGVEL: f = {2, {V, B}, 3}; return true;
    case B200_F_HINGE_VEC: case B200_F_HINGE_NORM: f = {5, {P, V, P, V, U}, 0}; return true;
    case B200_F_JOINTPOS_VEC: case B200_F_JOINTPOS_NORM: f = {4, {P, P, V, V}, 0}; return true;
    case B200_F_JOINTVEL_VEC: case B200_F_JOINTVEL_NORM: f = {8, {P, V, V, V, P, V, V, V}, 0}; return true;
    case B200_F_SEGLEN_MAG: case B200_F_SEGLEN_MAX: case B200_F_SEGLEN_MIN: f = {2, {V, V}, 0}; return true;
    case B200_F_SEGLEN_DISCREPANCY: f = {4, {V, V, V, V}, 0}; return true;
    case B200_F_ANGLE_AXIS_SEG: f = {3, {U, V, V}, 0}; return true;
    case B200_F_POSE3_TRANSLATION_PRIOR: case B200_F_POSE3_COMPASS_PRIOR: case B200_F_PRIOR_POSE3: f = {1, {P}, 0}; return true;
    case B200_F_PRIOR_VEC3: f = {1, {V}, 0}; return true;
    case B200_F_PRIOR_BIAS6: f = {1, {B}, 0}; return true;
    case B200_F_PRIOR_UNIT3: f = {1, {U}, 0}; return true;
    case B200_F_MAG_POSE3: f = {1, {P}, 3}; return true;
    default: return false;
    }
}

#define B200_MAX_SPLIT 11   // CTAs per chunk sweep on the upper levels (arrow rows split in whole 16-row tiles: 165 rows = 11 tiles)
enum { SC_LIN_ERR = 0, SC_MODEL_ERR = 1, SC_NEW_ERR = 2, SC_CUR_ERR = 3, SC_FAIL = 4, SC_COUNT = 8 };

// Levenberg-Marquardt state on the device (north_star subsystem 4: "the trust-region / damping update on device"): written by
// lm_judge_kernel after every lambda trial, read by the host once per iterate().
struct LmDev {
    double lambda;       // damping of the LM state (what the next iterate() starts from)
    double lambda_try;   // damping of the trial being solved (sb.lam points here)
    double err;          // cost at the current values
    double new_err;      // cost of the last accepted trial
    int iterations, total_trials, inner, status;
    int done;            // the tryLambda loop of this iterate() has ended: queued trials return at once (sb.skip points here)
    int accepted;        // the try buffers hold the accepted step
    int pad0, pad1;
};

#ifndef B200_USE_EMU
// NCCL, resolved at run time from the libnccl.so.2 the process already has (torch's bundled copy in the Python host) or the
// system one.  In-library collectives (b200_nccl_init) remove the host round trip of the b200_comm_set callbacks.
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommSplit) CommSplit = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};
static NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
#define B200_NCCL_SYM(n) api.n = reinterpret_cast<decltype(api.n)>(dlsym(h, "nccl" #n))
    B200_NCCL_SYM(GetUniqueId); B200_NCCL_SYM(CommInitRank); B200_NCCL_SYM(CommSplit); B200_NCCL_SYM(CommDestroy); B200_NCCL_SYM(AllGather);
    B200_NCCL_SYM(AllReduce); B200_NCCL_SYM(GroupStart); B200_NCCL_SYM(GroupEnd); B200_NCCL_SYM(GetErrorString);
#undef B200_NCCL_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommSplit && api.CommDestroy && api.AllGather && api.AllReduce && api.GroupStart && api.GroupEnd;
    return api;
}
struct NcclCtx { ncclComm_t comm = nullptr; int nranks = 1, rank = 0; };
static int nccl_allgather_cb(void* ctx, void* buf, uint64_t bytes_per_rank, void* stream) {
    NcclCtx* c = static_cast<NcclCtx*>(ctx);
    return nccl_api().AllGather(static_cast<char*>(buf) + (size_t)c->rank * bytes_per_rank, buf, bytes_per_rank / sizeof(double), ncclFloat64, c->comm,
                                static_cast<cudaStream_t>(stream)) == ncclSuccess ? 0 : 1;
}
static int nccl_allreduce_cb(void* ctx, double* buf, int32_t n, void* stream) {
    NcclCtx* c = static_cast<NcclCtx*>(ctx);
    return nccl_api().AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, c->comm, static_cast<cudaStream_t>(stream)) == ncclSuccess ? 0 : 1;
}
#endif

struct b200_problem {
    int device = 0;
    cudaStream_t stream = nullptr;
    int K = 0, ntv = 0, nsv = 0, cstride = 0, n_slots = 0;
    std::vector<int> tvtype, svtype;
    std::vector<b200_factor_slot> slots;
    std::vector<DevSlot> hslots;
    Dims dm{};   // assembly / internal tangent layout [local | chain]
    Dims ds{};   // solver layout: == dm, or the chain dofs only when the keyframe-local dofs are pre-eliminated (kernels_local.cuh)
    bool use_le = false;
    LocalElim le{};
    size_t smem_le = 0;
    double *d_Dc = nullptr, *d_Ac = nullptr, *d_Sc = nullptr, *d_V = nullptr, *d_Wl = nullptr;
    double* d_red_partial = nullptr; int* d_red_ticket = nullptr;   // scratch of reduce_sum_kernel (per problem: problems run concurrently)
    int* d_le_fresh = nullptr;   // device flag: the reduced blocks do not belong to the current linearization yet (kernels_local.cuh)
    double* d_delta = nullptr;   // full step [K][ntp] in internal order (== sb.delta without pre-elimination)
    LmDev* d_lm = nullptr;
    LmDev* h_lm = nullptr;       // pinned mirror
    int tvd = 0, svd = 0;
    std::vector<int> tv_valoff, tv_tan_int, tv_tan_ext, sv_valoff, sv_tan;
    std::vector<int> int2ext;  // internal tangent index -> external (or -1 for the padding variable)
    // device buffers
    double *d_tv = nullptr, *d_tv_try = nullptr, *d_sv = nullptr, *d_sv_try = nullptr, *d_cst = nullptr, *d_stage = nullptr;
    size_t stage_doubles = 0;
    DevSlot* d_slots = nullptr;
    DevVar *d_tvars = nullptr, *d_svars = nullptr;
    double *d_J = nullptr, *d_r = nullptr, *d_e = nullptr, *d_e2 = nullptr;
    long long nJ = 0, nr = 0, ne = 0, max_rows = 0;
    double* d_e2r = nullptr;     // per factor ROW: 0.5 (J delta + r)^2 of the trial step (model_error_rows_kernel)
    int max_inst = 0;
    double *d_D = nullptr, *d_O = nullptr, *d_A = nullptr, *d_S = nullptr;
    SolveBuffers sb{};
    struct HostLevel {
        int n_nodes = 0, P = 1, c_lo = 0, c_hi = 1, nsplit = 1;
        std::vector<int> begin, st, sep_st;     // partition over nodes; storage index per node; per separator
        int *d_begin = nullptr, *d_st = nullptr, *d_sep_st = nullptr;
        double *F = nullptr, *S = nullptr, *Dred = nullptr, *Ored = nullptr, *Ared = nullptr, *Sred = nullptr;
        Level dev{};
    };
    std::vector<HostLevel> levels;             // levels[0]: keyframes, levels[1]: its separators (optional)
    int P = 1, P2 = 1;                         // chunks of level 0 / level 1
    Level top{};                               // top level: one chunk over the remaining separators (or all keyframes)
    int* d_top_begin = nullptr;
    int *d_orow = nullptr, *d_ocol = nullptr, *d_ozs = nullptr, *d_ogrp = nullptr;
    double *d_Ftop = nullptr, *d_Stop = nullptr;
    size_t smem_schur = 0, smem_backsub = 0;
    int schur_ldy = 0, schur_kslab = 0, schur_tile_ldy = 0;
    size_t smem_schur_tile = 0;
    bool schur_tile_ok = false;
    int Wk = 0;                                // keyframes per rank window
    double *d_stageS = nullptr, *d_stageF = nullptr, *d_pack = nullptr;
    size_t pack_doubles = 0;
    size_t smem_sweep = 0, smem_asm[2] = {0, 0};
    AsmPlan asm_plan{};
    AsmItem* d_asm_items = nullptr; AsmTerm* d_asm_terms = nullptr; AsmBlock* d_asm_blocks[2] = {nullptr, nullptr};
    AsmElem* d_asm_elems[2] = {nullptr, nullptr};
    double* h_scal = nullptr;  // pinned: SC_COUNT doubles + status
    // LM state
    b200_lm_params lm{};
    double err = 0, lambda = 0;
    int iterations = 0, total_trials = 0;
    bool have_lin = false, err_stale = false, lm_started = false;
    long long launches = 0, host_syncs = 0;
    int spec_trials = 2;           // lambda trials queued per host synchronisation (device-resident ladder)
    bool use_graph = true;         // launch every trial as an instantiated CUDA graph (single rank)
    int parity = 0;                // which of the two value buffers is current (one trial graph per parity)
#ifndef B200_USE_EMU
    cudaGraphExec_t trial_graph[2] = {nullptr, nullptr};
#endif
    long long graph_launches = 0;  // kernels inside one trial graph
    cudaEvent_t ev[16] = {};
    cudaEvent_t tev[4][12] = {};   // per queued trial: phase events (TE_*)
    double phase_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // lin, solve, retract+error, sweep ms, sweep launches, top ms, top launches, backsub ms,
                                                                 // local elimination, level-0 Schur + boundary exchange + reduce, level 1, levels >= 2
    std::vector<float> pending;
    bool time_phases = false;
    // sharding (time axis): this rank owns chunks [c_lo, c_hi) == keyframes [k_lo, k_hi)
    int world = 1, rank = 0, c_lo = 0, c_hi = 1, k_lo = 0, k_hi = 0;
    b200_allgather_fn ag_fn = nullptr;
    b200_allreduce_fn ar_fn = nullptr;
    void* comm_ctx = nullptr;
    // lambda speculation (b200_spec_set): spec_G groups of `world` ranks try lambda, lambda*factor, ... concurrently
    int spec_G = 1, spec_g = 0, spec_rec = 0;
    b200_allgather_fn spec_ag = nullptr;
    void* spec_ctx = nullptr;
    double *d_spec_delta = nullptr, *d_spec_rec = nullptr, *h_spec_rec = nullptr;
#ifndef B200_USE_EMU
    NcclCtx nccl_all, nccl_grp;    // in-library NCCL (b200_nccl_init): all ranks / the ranks of this lambda group
    bool own_nccl = false;
#endif
    double comm_ms = 0;            // device time of the collectives (phase timing on)
    long long comm_calls = 0;
    cudaEvent_t cev[2] = {};
};

// B200_POISON=1 (debugging): fill every allocation with 0xFF bytes (NaN doubles / -1 ints) so that a read of memory the
// library never wrote shows up in the results instead of depending on what the allocator handed out.
template <class T>
static cudaError_t dalloc(T** p, size_t n) {
    static const bool poison = std::getenv("B200_POISON") != nullptr;
    cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess && poison) e = cudaMemset(*p, 0xFF, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess && poison) e = cudaDeviceSynchronize();
    return e;
}

// shared-memory row stride == 4 mod 8 doubles: the 8 x 4 / 4 x 8 operand fragments of mma.m8n8k4.f64 (lane -> row g = lane >> 2,
// element t = lane & 3) then fall into 16 distinct 8-byte bank pairs per half warp, for the A role, the NN B role and the NT B role
static int choose_ld(int ntp) { int ld = ntp; while ((ld & 7) != 4) ld++; return ld; }

// assembly plan: see kernels_assemble.cuh
static b200_status build_asm_plan(b200_problem* p, size_t smem_cap) {
    const Dims& dm = p->dm;
    std::vector<AsmItem> items;
    std::vector<std::vector<int>> item_of(p->n_slots, std::vector<int>(2, -1));
    int off = 0;
    for (int si = 0; si < p->n_slots; si++) {
        const DevSlot& s = p->hslots[si];
        bool k1 = false;
        for (int v = 0; v < s.nvars; v++) k1 |= (s.var_kind[v] == B200_AT_K1);
        for (int sel = 0; sel < (k1 ? 2 : 1); sel++) {
            item_of[si][sel] = (int)items.size();
            items.push_back({si, sel, off});
            off += s.rows * (s.cols + 1);
        }
    }
    if (items.size() > 256) { set_err("more than 256 factor instances per keyframe"); return B200_NOT_IMPLEMENTED; }
    struct Key { int half, which, row, col; bool operator<(const Key& o) const { return std::tie(half, which, row, col) < std::tie(o.half, o.which, o.row, o.col); } };
    struct Blk { int da, db; std::vector<AsmTerm> terms; };
    std::map<Key, Blk> blocks;
    auto add = [&](int half, int which, int row, int col, int da, int db, AsmTerm t) {
        Blk& b = blocks[Key{half, which, row, col}];
        b.da = da; b.db = db; b.terms.push_back(t);
    };
    for (int si = 0; si < p->n_slots; si++) {
        const DevSlot& s = p->hslots[si];
        for (int sel = 0; sel < 2; sel++) {
            const int it = item_of[si][sel];
            if (it < 0) continue;
            for (int a = 0; a < s.nvars; a++) {
                const int kda = s.var_kind[a];
                const bool a_static = kda == B200_AT_STATIC;
                const bool a_atk = (kda == B200_AT_K && sel == 0) || (kda == B200_AT_K1 && sel == 1);
                const bool a_atk1 = (kda == B200_AT_K1 && sel == 0);
                const int ta = s.tan_off[a], da = s.tan_dim[a], ca = s.col_off[a];
                if (a_atk) add(1, 1, dm.ns, ta, 1, da, AsmTerm{it, ca, -1, 2, 0, 0, 0, 0});                       // Ax rhs row: -g_k
                if (a_static && sel == 0) {
                    add(0, 1, ta, dm.ns, da, 1, AsmTerm{it, ca, -1, 1, 0, 0, 0, 0});                               // Sx rhs column
                    add(0, 1, dm.ns, ta, 1, da, AsmTerm{it, ca, -1, 2, 0, 0, 0, 0});                               // Sx rhs row
                }
                for (int b = 0; b < s.nvars; b++) {
                    const int kdb = s.var_kind[b];
                    const bool b_static = kdb == B200_AT_STATIC;
                    const bool b_atk = (kdb == B200_AT_K && sel == 0) || (kdb == B200_AT_K1 && sel == 1);
                    const int tb = s.tan_off[b], db = s.tan_dim[b], cb = s.col_off[b];
                    // D_k: the solver reads the LOWER triangle only (cp_lower_async): blocks strictly above the diagonal are not built
                    if (a_atk && b_atk) { if (ta >= tb) add(0, 0, ta, tb, da, db, AsmTerm{it, ca, cb, 0, 0, 0, 0, 0}); }
                    else if (a_static && b_static && sel == 0) add(0, 1, ta, tb, da, db, AsmTerm{it, ca, cb, 0, 0, 0, 0, 0});
                    if (a_atk1 && b_atk) add(1, 0, ta - dm.nl, tb - dm.nl, da, db, AsmTerm{it, ca, cb, 0, 0, 0, 0, 0});
                    else if (a_static && b_atk) add(1, 1, ta, tb, da, db, AsmTerm{it, ca, cb, 0, 0, 0, 0, 0});
                }
            }
        }
    }
    std::vector<AsmTerm> terms;
    std::vector<AsmBlock> hb[2];
    for (auto& kv : blocks) {
        AsmBlock b{kv.first.which, kv.first.row, kv.first.col, kv.second.da, kv.second.db, (int)terms.size(), 0};
        for (AsmTerm t : kv.second.terms) {
            const AsmItem& im = items[t.item];
            t.stage_off = im.stage_off; t.rows = p->hslots[im.slot].rows; t.cols = p->hslots[im.slot].cols; t.pad = 0;
            terms.push_back(t);
        }
        b.term_end = (int)terms.size();
        if (b.da * b.db > 64) { set_err("internal: destination block larger than 64 entries"); return B200_NOT_IMPLEMENTED; }
        hb[kv.first.half].push_back(b);
    }
    p->smem_asm[0] = sizeof(double) * (size_t)off;   // staged Jacobians only: the dense blocks are written straight to global memory
    p->smem_asm[1] = sizeof(double) * (size_t)off;
    if (std::max(p->smem_asm[0], p->smem_asm[1]) > smem_cap) { set_err("time block + staged Jacobians do not fit in shared memory"); return B200_NOT_IMPLEMENTED; }
    CK(dalloc(&p->d_asm_items, items.size())); CK(dalloc(&p->d_asm_terms, terms.size()));
    CK(cudaMemcpyAsync(p->d_asm_items, items.data(), sizeof(AsmItem) * items.size(), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->d_asm_terms, terms.data(), sizeof(AsmTerm) * terms.size(), cudaMemcpyHostToDevice, p->stream));
    for (int h = 0; h < 2; h++) {
        CK(dalloc(&p->d_asm_blocks[h], hb[h].size()));
        CK(cudaMemcpyAsync(p->d_asm_blocks[h], hb[h].data(), sizeof(AsmBlock) * hb[h].size(), cudaMemcpyHostToDevice, p->stream));
        // entry list, block after block (row-major inside a block)
        std::vector<AsmElem> he;
        for (size_t b = 0; b < hb[h].size(); b++)
            for (int ia = 0; ia < hb[h][b].da; ia++)
                for (int ib = 0; ib < hb[h][b].db; ib++) he.push_back(AsmElem{(int)b, (short)ia, (short)ib});
        CK(dalloc(&p->d_asm_elems[h], he.size()));
        CK(cudaMemcpyAsync(p->d_asm_elems[h], he.data(), sizeof(AsmElem) * he.size(), cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));   // `he` is a local
        p->asm_plan.elems[h] = p->d_asm_elems[h]; p->asm_plan.n_elems[h] = (int)he.size();
    }
    CK(cudaStreamSynchronize(p->stream));
    AsmPlan& pl = p->asm_plan;
    pl.items = p->d_asm_items; pl.n_items = (int)items.size(); pl.terms = p->d_asm_terms;
    for (int h = 0; h < 2; h++) { pl.blocks[h] = p->d_asm_blocks[h]; pl.n_blocks[h] = (int)hb[h].size(); }
    pl.stage_doubles = off;
    return B200_OK;
}

static void free_levels(b200_problem* p) {
    for (auto& l : p->levels) {
        void* q[] = {l.d_begin, l.d_st, l.d_sep_st, l.F, l.S, l.Dred, l.Ored, l.Ared, l.Sred};
        for (void* x : q) if (x) cudaFree(x);
    }
    p->levels.clear();
}

extern "C" {

static void free_graphs(b200_problem* p);

const char* b200_last_error(void) { return g_last_error.c_str(); }

void b200_problem_destroy(b200_problem* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    void* ptrs[] = {p->d_tv, p->d_tv_try, p->d_sv, p->d_sv_try, p->d_cst, p->d_stage, p->d_slots, p->d_tvars, p->d_svars, p->d_J, p->d_r,
                    p->d_e, p->d_e2, p->d_D, p->d_O, p->d_A, p->d_S, p->sb.Wst, p->sb.Yst, p->sb.ybuf, p->d_Ftop, p->d_Stop, p->sb.delta, p->sb.dstat,
                    p->d_top_begin, p->d_orow, p->d_ocol, p->d_ozs, p->d_ogrp,
                    p->sb.scal, p->sb.status, p->d_stageS, p->d_stageF, p->d_pack, p->d_asm_items, p->d_asm_terms, p->d_asm_blocks[0], p->d_asm_blocks[1],
                    p->d_asm_elems[0], p->d_asm_elems[1]};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (p->use_le && p->d_delta) cudaFree(p->d_delta);
    for (void* q : {(void*)p->d_Dc, (void*)p->d_Ac, (void*)p->d_Sc, (void*)p->d_V, (void*)p->d_Wl, (void*)p->d_lm, (void*)p->d_le_fresh, (void*)p->d_red_partial, (void*)p->d_red_ticket, (void*)p->d_e2r}) if (q) cudaFree(q);
    if (p->h_lm) cudaFreeHost(p->h_lm);
    if (p->d_spec_delta) cudaFree(p->d_spec_delta);
    if (p->d_spec_rec) cudaFree(p->d_spec_rec);
    if (p->h_spec_rec) cudaFreeHost(p->h_spec_rec);
    free_levels(p);
    if (p->h_scal) cudaFreeHost(p->h_scal);
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    for (auto& q : p->tev) for (auto& e : q) if (e) cudaEventDestroy(e);
    for (auto& e : p->cev) if (e) cudaEventDestroy(e);
#ifndef B200_USE_EMU
    if (p->own_nccl) {
        if (p->nccl_grp.comm && p->nccl_grp.comm != p->nccl_all.comm) nccl_api().CommDestroy(p->nccl_grp.comm);
        if (p->nccl_all.comm) nccl_api().CommDestroy(p->nccl_all.comm);
    }
#endif
    free_graphs(p);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

static b200_status set_partition(b200_problem* p, int P1, int P2);

b200_status b200_problem_create(const b200_problem_desc* d, int32_t device, b200_problem** out) {
    if (!d || !out || d->K < 1 || d->n_time_vars < 1 || d->n_slots < 1) { set_err("bad description"); return B200_BAD_ARG; }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_err("no such CUDA device"); return B200_BAD_ARG; }
    CK(cudaSetDevice(device));
    b200_problem* p = new b200_problem();
    p->device = device;
    p->K = d->K; p->ntv = d->n_time_vars; p->nsv = d->n_static_vars; p->cstride = d->const_stride; p->n_slots = d->n_slots;
    p->tvtype.assign(d->time_var_type, d->time_var_type + d->n_time_vars);
    if (d->n_static_vars) p->svtype.assign(d->static_var_type, d->static_var_type + d->n_static_vars);
    p->slots.assign(d->slots, d->slots + d->n_slots);
    auto fail = [&](const std::string& m, b200_status st) { set_err(m); b200_problem_destroy(p); return st; };
    // ---- classify time variables: chained (appear in a factor spanning k and k+1) vs keyframe-local
    std::vector<char> chained(p->ntv, 0);
    for (const auto& s : p->slots) {
        FInfo fi;
        if (!finfo(s.type, fi) || s.nvars != fi.nvars) return fail("slot: unknown type or wrong variable count", B200_BAD_ARG);
        bool spans = false, any_time = false;
        for (int v = 0; v < s.nvars; v++) {
            const int kind = s.var_kind[v], idx = s.var_index[v];
            if (kind == B200_AT_STATIC) { if (idx < 0 || idx >= p->nsv || p->svtype[idx] != fi.vt[v]) return fail("slot: bad static variable", B200_BAD_ARG); }
            else if (kind == B200_AT_K || kind == B200_AT_K1) {
                if (idx < 0 || idx >= p->ntv || p->tvtype[idx] != fi.vt[v]) return fail("slot: bad time variable", B200_BAD_ARG);
                any_time = true;
                if (kind == B200_AT_K1) spans = true;
            } else return fail("slot: bad var_kind", B200_BAD_ARG);
        }
        if (s.k_begin < 0 || s.k_end > d->K - (spans ? 1 : 0) || s.k_begin >= s.k_end) return fail("slot: bad keyframe range", B200_BAD_ARG);
        if (!any_time && !(s.k_begin == 0 && s.k_end == 1)) return fail("static-only slot must have k range [0,1)", B200_BAD_ARG);
        if ((fi.nconst > 0) != (s.const_offset >= 0)) return fail("slot: const_offset mismatch", B200_BAD_ARG);
        if (fi.nconst > 0 && s.const_offset + fi.nconst > d->const_stride) return fail("slot: constants out of range", B200_BAD_ARG);
        if (spans) for (int v = 0; v < s.nvars; v++) if (s.var_kind[v] != B200_AT_STATIC) chained[s.var_index[v]] = 1;
    }
    // ---- layouts
    p->tv_valoff.resize(p->ntv); p->tv_tan_ext.resize(p->ntv); p->tv_tan_int.resize(p->ntv);
    int nt = 0;
    for (int v = 0; v < p->ntv; v++) { p->tv_valoff[v] = p->tvd; p->tv_tan_ext[v] = nt; p->tvd += kValDim[p->tvtype[v]]; nt += kTanDim[p->tvtype[v]]; }
    int nl = 0;
    for (int v = 0; v < p->ntv; v++) if (!chained[v]) { p->tv_tan_int[v] = nl; nl += kTanDim[p->tvtype[v]]; }
    int pos = nl;
    for (int v = 0; v < p->ntv; v++) if (chained[v]) { p->tv_tan_int[v] = pos; pos += kTanDim[p->tvtype[v]]; }
    p->sv_valoff.resize(p->nsv); p->sv_tan.resize(p->nsv);
    int ns = 0;
    for (int v = 0; v < p->nsv; v++) { p->sv_valoff[v] = p->svd; p->sv_tan[v] = ns; p->svd += kValDim[p->svtype[v]]; ns += kTanDim[p->svtype[v]]; }
    Dims& dm = p->dm;
    dm.K = p->K; dm.Kld = (p->K + 3) & ~3LL;
    dm.nt = nt; dm.ntp = nt + (nt & 1); dm.nl = nl; dm.nc = dm.ntp - nl; dm.ns = ns; dm.ns1 = ns + 1;
    if (dm.ntp > 256 || dm.ns1 + dm.nc + 1 > 512) return fail("time block larger than 256 or arrow larger than 512 tangent dims is not supported", B200_NOT_IMPLEMENTED);
    // ---- solver layout: with keyframe-local dofs (and a chain to solve) they are eliminated per keyframe before the sweep, the
    // chain solver then sees the time-chained dofs only (padded to an even count)
    p->use_le = nl > 0 && nl <= B200_LE_MAXNL && dm.nc >= 2 && dm.nc + dm.ns1 <= B200_LE_THREADS - 32 && !std::getenv("B200_NO_LOCAL_ELIM");   // (one thread per panel row)
    Dims& ds = p->ds;
    ds = dm;
    if (p->use_le) { ds.nt = dm.nc; ds.ntp = (dm.nc + 1) & ~1; ds.nl = 0; ds.nc = ds.ntp; }
    dm.nco = ds.nc;   // O_k is assembled with the solver's chain dimension as its stride
    ds.nco = ds.nc;
    ds.ld = choose_ld(ds.ntp); ds.hmax = ds.ns1 + ds.nc;
    dm.ld = ds.ld; dm.hmax = ds.hmax;
    p->int2ext.assign(dm.ntp, -1);
    for (int v = 0; v < p->ntv; v++) for (int j = 0; j < kTanDim[p->tvtype[v]]; j++) p->int2ext[p->tv_tan_int[v] + j] = p->tv_tan_ext[v] + j;
    // ---- device slots
    p->hslots.resize(p->n_slots);
    for (int si = 0; si < p->n_slots; si++) {
        const auto& s = p->slots[si];
        DevSlot& dsl = p->hslots[si];
        FInfo fi; finfo(s.type, fi);
        std::memset(&dsl, 0, sizeof(dsl));
        dsl.type = s.type; dsl.k_begin = s.k_begin; dsl.k_end = s.k_end; dsl.n_inst = s.k_end - s.k_begin;
        dsl.nvars = s.nvars; dsl.rows = factor_rows(s.type); dsl.const_off = s.const_offset;
        int col = 0;
        for (int v = 0; v < s.nvars; v++) {
            dsl.var_kind[v] = s.var_kind[v]; dsl.var_type[v] = fi.vt[v];
            const int idx = s.var_index[v];
            if (s.var_kind[v] == B200_AT_STATIC) { dsl.val_off[v] = p->sv_valoff[idx]; dsl.tan_off[v] = p->sv_tan[idx]; }
            else { dsl.val_off[v] = p->tv_valoff[idx]; dsl.tan_off[v] = p->tv_tan_int[idx]; }
            dsl.col_off[v] = col; dsl.tan_dim[v] = kTanDim[fi.vt[v]]; col += dsl.tan_dim[v];
        }
        dsl.cols = col;
        dsl.J_off = p->nJ; dsl.r_off = p->nr; dsl.e_off = p->ne;
        p->nJ += (long long)dsl.rows * dsl.cols * dsl.n_inst; p->nr += (long long)dsl.rows * dsl.n_inst; p->ne += dsl.n_inst;
        p->max_rows = std::max(p->max_rows, (long long)dsl.rows * dsl.n_inst);
        p->max_inst = std::max(p->max_inst, dsl.n_inst);
        std::memcpy(dsl.params, s.params, sizeof(dsl.params));
    }
    // ---- allocate
    const long long Kld = dm.Kld;
    CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    for (auto& e : p->ev) CK(cudaEventCreate(&e));
    for (auto& q : p->tev) for (auto& e : q) CK(cudaEventCreate(&e));
    for (auto& e : p->cev) CK(cudaEventCreate(&e));
    if (const char* env = std::getenv("B200_NO_GRAPH")) p->use_graph = std::atoi(env) == 0;
    if (const char* env = std::getenv("B200_SPEC_TRIALS")) p->spec_trials = std::max(1, std::atoi(env));
    CK(dalloc(&p->d_tv, (size_t)p->tvd * Kld)); CK(dalloc(&p->d_tv_try, (size_t)p->tvd * Kld));
    CK(dalloc(&p->d_sv, (size_t)p->svd)); CK(dalloc(&p->d_sv_try, (size_t)p->svd));
    CK(dalloc(&p->d_cst, (size_t)std::max(p->cstride, 1) * Kld));
    p->stage_doubles = (size_t)p->K * std::max(std::max(p->tvd, p->cstride), dm.ntp);
    CK(dalloc(&p->d_stage, p->stage_doubles));
    CK(dalloc(&p->d_slots, (size_t)p->n_slots)); CK(dalloc(&p->d_tvars, (size_t)p->ntv)); CK(dalloc(&p->d_svars, (size_t)p->nsv));
    CK(dalloc(&p->d_J, (size_t)p->nJ)); CK(dalloc(&p->d_r, (size_t)p->nr)); CK(dalloc(&p->d_e, (size_t)p->ne)); CK(dalloc(&p->d_e2, (size_t)p->ne)); CK(dalloc(&p->d_e2r, (size_t)p->nr));
    const size_t K = p->K;
    CK(dalloc(&p->d_D, K * dm.ntp * dm.ntp)); CK(dalloc(&p->d_O, K * dm.nco * dm.nco));
    CK(dalloc(&p->d_A, K * dm.ns1 * dm.ntp)); CK(dalloc(&p->d_S, K * dm.ns1 * dm.ns1));
    // assemble_kernel only rewrites the structurally non-zero sub-blocks: everything else is zeroed here, once
    CK(cudaMemsetAsync(p->d_D, 0, sizeof(double) * K * dm.ntp * dm.ntp, p->stream)); CK(cudaMemsetAsync(p->d_O, 0, sizeof(double) * K * dm.nco * dm.nco, p->stream));
    CK(cudaMemsetAsync(p->d_A, 0, sizeof(double) * K * dm.ns1 * dm.ntp, p->stream)); CK(cudaMemsetAsync(p->d_S, 0, sizeof(double) * K * dm.ns1 * dm.ns1, p->stream));
    CK(dalloc(&p->sb.Wst, K * ds.ntp * ds.ntp)); CK(dalloc(&p->sb.Yst, K * (size_t)ds.hmax * ds.ntp)); CK(dalloc(&p->sb.ybuf, K * ds.ntp));
    CK(dalloc(&p->sb.delta, (K + 1024) * ds.ntp)); CK(dalloc(&p->sb.dstat, (size_t)ns + 1)); CK(dalloc(&p->sb.scal, (size_t)SC_COUNT + 8));   // [SC_COUNT..): rank-local scalars that are NOT all-reduced
    CK(dalloc(&p->sb.status, (size_t)1));
    CK(dalloc(&p->d_lm, (size_t)1));
    CK(dalloc(&p->d_red_partial, (size_t)B200_REDUCE_BLOCKS)); CK(dalloc(&p->d_red_ticket, (size_t)1));
    CK(cudaMemsetAsync(p->d_red_ticket, 0, sizeof(int), p->stream));
    CK(cudaMemsetAsync(p->d_lm, 0, sizeof(LmDev), p->stream));
    CK(cudaMallocHost((void**)&p->h_lm, sizeof(LmDev)));
    std::memset(p->h_lm, 0, sizeof(LmDev));
    p->sb.lam = &p->d_lm->lambda_try; p->sb.lam_in_sweep = p->use_le ? 0 : 1; p->sb.skip = nullptr;
    if (p->use_le) {
        // reduced blocks (what the chain solver reads), the local factors, and the full step in internal order.  Only the entries
        // local_elim_kernel writes ever change: the padding column / the upper triangle stay zero from here on
        CK(dalloc(&p->d_Dc, K * ds.ntp * ds.ntp)); CK(dalloc(&p->d_Ac, K * ds.ns1 * ds.ntp)); CK(dalloc(&p->d_Sc, K * ds.ns1 * ds.ns1));
        CK(cudaMemsetAsync(p->d_Dc, 0, sizeof(double) * K * ds.ntp * ds.ntp, p->stream));
        CK(cudaMemsetAsync(p->d_Ac, 0, sizeof(double) * K * ds.ns1 * ds.ntp, p->stream));
        CK(cudaMemsetAsync(p->d_Sc, 0, sizeof(double) * K * ds.ns1 * ds.ns1, p->stream));
        CK(dalloc(&p->d_V, K * (size_t)(dm.nc + dm.ns1) * nl)); CK(dalloc(&p->d_Wl, K * (size_t)nl * nl));
        CK(dalloc(&p->d_le_fresh, (size_t)1));
        CK(cudaMemsetAsync(p->d_le_fresh, 1, sizeof(int), p->stream));
        CK(dalloc(&p->d_delta, (K + 1024) * dm.ntp));
        CK(cudaMemsetAsync(p->d_delta, 0, sizeof(double) * (K + 1024) * dm.ntp, p->stream));
        LocalElim& le = p->le;
        le.nl = nl; le.nc = dm.nc; le.ncp = ds.ntp; le.ns1 = dm.ns1; le.ntp = dm.ntp; le.ldl = le_ldl(nl);
        le.D = p->d_D; le.A = p->d_A; le.S = p->d_S; le.Dc = p->d_Dc; le.Ac = p->d_Ac; le.Sc = p->d_Sc; le.V = p->d_V; le.Wl = p->d_Wl;
        p->smem_le = sizeof(double) * (size_t)le_smem_doubles(nl, dm.nc, dm.ns1);
    } else p->d_delta = p->sb.delta;
    CK(dalloc(&p->d_Ftop, (size_t)B200_MAX_SPLIT * 2 * (ds.nc + ds.hmax) * ds.nc)); CK(dalloc(&p->d_Stop, (size_t)ds.hmax * ds.hmax));
    CK(dalloc(&p->d_top_begin, (size_t)2));
    {
        // structure of the time-coupling block O in SOLVER coordinates: which chain dofs of keyframe k+1 (rows) meet which chain
        // dofs of k (columns); dm.nl = first chain dof in the internal order, ds.nl = first chain column of the solver's time block
        const int nc = ds.nc, nl_ = ds.nl;
        std::vector<int> orow(2 * nc), ocol(2 * nc), ozs(nc);
        for (int i = 0; i < nc; i++) { orow[2 * i] = nc; orow[2 * i + 1] = 0; ocol[2 * i] = nc; ocol[2 * i + 1] = 0; }
        for (const DevSlot& s : p->hslots)
            for (int a = 0; a < s.nvars; a++) {
                if (s.var_kind[a] != B200_AT_K1) continue;
                for (int b = 0; b < s.nvars; b++) {
                    if (s.var_kind[b] != B200_AT_K) continue;
                    const int ra = s.tan_off[a] - dm.nl, cb = s.tan_off[b] - dm.nl;
                    for (int i = ra; i < ra + s.tan_dim[a]; i++) { orow[2 * i] = std::min(orow[2 * i], cb); orow[2 * i + 1] = std::max(orow[2 * i + 1], cb + s.tan_dim[b]); }
                    for (int j = cb; j < cb + s.tan_dim[b]; j++) { ocol[2 * j] = std::min(ocol[2 * j], ra); ocol[2 * j + 1] = std::max(ocol[2 * j + 1], ra + s.tan_dim[a]); }
                }
            }
        int ow = 2;
        for (int i = 0; i < nc; i++) {
            if (orow[2 * i + 1] <= orow[2 * i]) { orow[2 * i] = orow[2 * i + 1] = 0; }
            if (ocol[2 * i + 1] <= ocol[2 * i]) { ocol[2 * i] = ocol[2 * i + 1] = 0; }
            ozs[i] = (nl_ + orow[2 * i]) & ~1;
            ow = std::max(ow, (nl_ + orow[2 * i + 1] - ozs[i] + 1) & ~1);
        }
        ow = std::min(ow, ds.ntp);
        for (int i = 0; i < nc; i++) if (ozs[i] + ow > ds.ntp) ozs[i] = ds.ntp - ow;
        ds.ow = ow; ds.owp = ow; while ((ds.owp & 7) != 4) ds.owp++;   // packed-row stride == 4 mod 8 (conflict-free B-operand loads)
        std::vector<int> ogrp;
        for (int b0 = 0; b0 < nc;) {   // runs of <= 16 consecutive rows with the same window, dealt to 4-row groups with stride
            int b1 = b0 + 1;
            while (b1 < nc && b1 - b0 < 16 && ozs[b1] == ozs[b0]) b1++;
            const int str = (b1 - b0 + 3) / 4;
            for (int c = 0; c < str; c++) {
                int cnt = 0;
                while (cnt < 4 && b0 + c + str * cnt < b1) cnt++;
                ogrp.push_back(b0 + c); ogrp.push_back(str); ogrp.push_back(cnt);
            }
            b0 = b1;
        }
        ds.n_ogrp = (int)ogrp.size() / 3;
        CK(dalloc(&p->d_ogrp, ogrp.size()));
        CK(cudaMemcpyAsync(p->d_ogrp, ogrp.data(), sizeof(int) * ogrp.size(), cudaMemcpyHostToDevice, p->stream));
        ds.ogrp = p->d_ogrp;
        CK(dalloc(&p->d_orow, (size_t)2 * nc)); CK(dalloc(&p->d_ocol, (size_t)2 * nc)); CK(dalloc(&p->d_ozs, (size_t)nc));
        CK(cudaMemcpyAsync(p->d_orow, orow.data(), sizeof(int) * 2 * nc, cudaMemcpyHostToDevice, p->stream));
        CK(cudaMemcpyAsync(p->d_ocol, ocol.data(), sizeof(int) * 2 * nc, cudaMemcpyHostToDevice, p->stream));
        CK(cudaMemcpyAsync(p->d_ozs, ozs.data(), sizeof(int) * nc, cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        ds.orow = p->d_orow; ds.ocol = p->d_ocol; ds.ozs = p->d_ozs;
        dm.ow = ds.ow; dm.owp = ds.owp; dm.orow = ds.orow; dm.ocol = ds.ocol; dm.ozs = ds.ozs; dm.ogrp = ds.ogrp; dm.n_ogrp = ds.n_ogrp;
    }
    CK(cudaMallocHost((void**)&p->h_scal, (SC_COUNT + 2) * sizeof(double)));
    // uploads
    // every upload goes through p->stream (a non-blocking stream is not ordered against legacy-stream copies)
    CK(cudaMemcpyAsync(p->d_slots, p->hslots.data(), sizeof(DevSlot) * p->n_slots, cudaMemcpyHostToDevice, p->stream));
    std::vector<DevVar> tvars(p->ntv), svars(std::max(p->nsv, 1));
    for (int v = 0; v < p->ntv; v++) tvars[v] = {p->tvtype[v], p->tv_valoff[v], p->tv_tan_int[v]};
    for (int v = 0; v < p->nsv; v++) svars[v] = {p->svtype[v], p->sv_valoff[v], p->sv_tan[v]};
    CK(cudaMemcpyAsync(p->d_tvars, tvars.data(), sizeof(DevVar) * p->ntv, cudaMemcpyHostToDevice, p->stream));
    if (p->nsv) CK(cudaMemcpyAsync(p->d_svars, svars.data(), sizeof(DevVar) * p->nsv, cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaMemsetAsync(p->d_cst, 0, sizeof(double) * std::max(p->cstride, 1) * Kld, p->stream));  // same stream as the transpose below (a legacy-stream memset would race with it)
    if (p->cstride > 0 && d->consts) {
        CK(cudaMemcpyAsync(p->d_stage, d->consts, sizeof(double) * K * p->cstride, cudaMemcpyHostToDevice, p->stream));
        const long long n = (long long)K * p->cstride;
        B200_LAUNCH(transpose_in_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, p->stream, p->d_stage, p->K, p->cstride, Kld, p->d_cst);
        p->launches++;
        CK(cudaStreamSynchronize(p->stream));
    }
    // ---- shared memory budgets
    size_t static_smem = 0;
    {
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, top_finish_kernel)); static_smem = std::max(static_smem, (size_t)fa.sharedSizeBytes);
        CK(cudaFuncGetAttributes(&fa, chain_sweep_kernel<false>)); static_smem = std::max(static_smem, (size_t)fa.sharedSizeBytes);
        CK(cudaFuncGetAttributes(&fa, chain_sweep_kernel<true>)); static_smem = std::max(static_smem, (size_t)fa.sharedSizeBytes);
        CK(cudaFuncGetAttributes(&fa, chain_backsub_kernel)); static_smem = std::max(static_smem, (size_t)fa.sharedSizeBytes);
        CK(cudaFuncGetAttributes(&fa, chain_schur_kernel)); static_smem = std::max(static_smem, (size_t)fa.sharedSizeBytes);
        CK(cudaFuncGetAttributes(&fa, assemble_kernel)); static_smem = std::max(static_smem, (size_t)fa.sharedSizeBytes);
    }
    const size_t smem_cap = (227 * 1024 - static_smem - 1024) & ~(size_t)15;  // dynamic budget next to the kernels' static arrays
    {
        b200_status ast = build_asm_plan(p, smem_cap);
        if (ast != B200_OK) { b200_problem_destroy(p); return ast; }
    }
    if (p->use_le && p->smem_le > smem_cap) return fail("local block does not fit in shared memory", B200_NOT_IMPLEMENTED);
    {
        // sweep: pivot block + packed O rows resident, the panel streams through in chunks of cr rows
        const long long cap_d = (long long)(smem_cap / sizeof(double));
        const long long fixed = (long long)ds.ntp * ds.ld + sweep_op_doubles(ds.nc, ds.owp);
        long long rows_fit = (cap_d - fixed) / ds.ld;
        if (rows_fit < 16) return fail("time block does not fit in shared memory", B200_NOT_IMPLEMENTED);
        ds.cr = (int)std::min<long long>(64, rows_fit & ~15LL);
        p->smem_sweep = sizeof(double) * (size_t)sweep_smem_doubles(ds.ntp, ds.ld, ds.cr, ds.nc, ds.owp);
        // keyframe-pipelined level-0 sweep: needs the solver layout without local dofs and its (larger) shared-memory carve-up
        // (row tiles of <= 128 columns fit the per-warp slice of Tensor Memory)
        ds.pipe = (ds.nl == 0 && ds.ntp >= 64 && ds.ntp <= 128 && (ds.ow & 1) == 0 && !std::getenv("B200_NO_PIPE")) ? 1 : 0;
        if (ds.pipe) {
            const size_t need = sizeof(double) * (size_t)sweep_pipe_smem_doubles(ds.ntp, ds.ld, ds.nc, ds.owp);
            if (need <= smem_cap) p->smem_sweep = std::max(p->smem_sweep, need); else ds.pipe = 0;
            if (ds.pipe && std::getenv("B200_NO_DENSE_STREAM")) ds.pipe = 2;   // (development switch: level 0 pipelined, dense levels on the round-1 kernel)
        }
        // back-substitution / top level: factor block + vectors; the static system (ns1 x lds) reuses the block area
        ds.bs_os = (sizeof(double) * (size_t)backsub_smem_doubles(ds.ntp, ds.ld, ds.nc) + 16 <= smem_cap && !std::getenv("B200_NO_BS_OSTAGE")) ? 1 : 0;
        const size_t bs = (size_t)backsub_smem_doubles(ds.ntp, ds.ld, ds.bs_os ? ds.nc : 0);
        if ((size_t)static_lds(ds.ns) * (ds.ns1 + 1) > (size_t)ds.ntp * ds.ld + 4 * 256)
            return fail("static block larger than the time block is not supported", B200_NOT_IMPLEMENTED);
        p->smem_backsub = (sizeof(double) * bs + 15) & ~(size_t)15;
        if (p->smem_backsub > smem_cap) return fail("solver working set does not fit in shared memory", B200_NOT_IMPLEMENTED);
        ds.sm_doubles = (int)(p->smem_sweep / sizeof(double));
        // Schur kernel: the stored arrow panel of one node (hmax rows), whole rows if they fit, else column slabs
        int ldy = choose_ld(ds.ntp), kslab = ds.ntp;
        if ((long long)ds.hmax * ldy > cap_d) {
            kslab = (int)((cap_d / ds.hmax) & ~7LL) - 4;
            if (kslab < 4) return fail("arrow panel does not fit in shared memory", B200_NOT_IMPLEMENTED);
            ldy = kslab;   // == 4 mod 8
        }
        p->schur_ldy = ldy; p->schur_kslab = kslab;
        p->smem_schur = sizeof(double) * (size_t)ds.hmax * ldy;
        p->schur_tile_ldy = choose_ld(ds.ntp);
        p->smem_schur_tile = sizeof(double) * (size_t)schur_tile_smem_doubles(p->schur_tile_ldy);
        p->schur_tile_ok = 64LL * p->schur_tile_ldy >= 4 * 32 * 33 && p->smem_schur_tile <= smem_cap && !std::getenv("B200_NO_SCHUR_TILE");
        dm.cr = ds.cr; dm.sm_doubles = ds.sm_doubles;
    }
#ifndef B200_USE_EMU
    CK(cudaFuncSetAttribute(assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(p->smem_asm[0], p->smem_asm[1])));
    CK(cudaFuncSetAttribute(chain_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_sweep));
    CK(cudaFuncSetAttribute(chain_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_sweep));
    CK(cudaFuncSetAttribute(chain_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_schur));
    if (p->schur_tile_ok) CK(cudaFuncSetAttribute(chain_schur_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_schur_tile));
    CK(cudaFuncSetAttribute(top_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_backsub));
    CK(cudaFuncSetAttribute(chain_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_backsub));
    if (p->use_le) CK(cudaFuncSetAttribute(local_elim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_le));
#endif
    p->k_lo = 0; p->k_hi = p->K;
    int P1 = 0, P2 = 0;   // 0: choose from the cost model
    if (const char* env = std::getenv("B200_CHUNKS")) P1 = std::atoi(env);
    if (const char* env = std::getenv("B200_GROUPS")) P2 = std::atoi(env);
    b200_status st = set_partition(p, P1, P2);
    if (st != B200_OK) { b200_problem_destroy(p); return st; }
    *out = p;
    return B200_OK;
}

// Upper levels (single rank): the m separators left by a level form a chain again.  Either one CTA sweeps them sequentially
// (cost m) or they are cut into P chunks whose interior nodes are eliminated in parallel (every step carries the left
// separator and a dense coupling block: ~2 units, plus a fixed per-level cost for the Schur / reduce / back-substitution
// launches) and the P-1 chunk separators recurse.  Dynamic programme over m; appends the chosen chunk counts to Ps.
static double plan_upper_levels(int m, std::vector<int>* Ps) {
    const double step_dense = 2.0, level_fixed = 1.5, step_top = 1.0;
    std::vector<double> T(m + 1, 0.0);
    std::vector<int> choice(m + 1, 1);
    for (int mm = 0; mm <= m; mm++) {
        T[mm] = step_top * mm;
        for (int P = 2; P <= mm / 2; P++) {
            int steps = 0;
            for (int j = 0; j < P; j++) {
                const int b0 = (int)((long long)j * mm / P), b1 = (j + 1 < P) ? (int)((long long)(j + 1) * mm / P) : mm;
                steps = std::max(steps, (j + 1 < P) ? b1 - b0 - 1 : b1 - b0);
            }
            const double t = step_dense * steps + level_fixed + T[P - 1];
            if (t < T[mm]) { T[mm] = t; choice[mm] = P; }
        }
    }
    const double total = T[m];
    if (Ps) while (m > 0 && choice[m] > 1) { Ps->push_back(choice[m]); m = choice[m] - 1; }
    return total;
}

// cost of a whole solve in the same units: level 0 (sparse coupling; every interior keyframe costs one sweep step plus its
// share of the Schur and back-substitution kernels, ~1.45 units) in waves of one chunk per SM, plus the upper levels
static double partition_cost(int K, int P1, int sms) {
    if (P1 <= 1) return (double)K;
    const double waves = std::ceil((double)P1 / sms);
    return waves * 1.45 * ((double)K / P1 - 1.0) + 1.5 + plan_upper_levels(P1 - 1, nullptr);
}

// Build the nested partition.  P1 chunks over the keyframes (every rank owns P1/world of them, inside its window of
// ceil(K/world) keyframes); P2 > 1: the P1-1 separators are themselves cut into P2 groups (rank-aligned) and only the
// P2-1 group separators are eliminated sequentially by the single top-level CTA.  Chunks hold >= 2 nodes.
static b200_status set_partition(b200_problem* p, int P1, int P2) {
    const Dims& dm = p->ds;   // solver layout
    const int K = p->K, W = p->world, sms = 148 * p->world;   // one chunk CTA per SM and wave, on every rank
    p->Wk = (K + W - 1) / W;
    if ((long long)(W - 1) * p->Wk >= K || K - (W - 1) * p->Wk < 3) { set_err("too few keyframes for this many ranks"); return B200_BAD_ARG; }
    const int max_p1 = std::max(1, K / 3);
    const bool auto_groups = P2 <= 0;
    if (P1 <= 0) {  // search the cost model (the upper levels recurse: plan_upper_levels)
        double best = 1e300; int bp1 = 1;
        for (int c1 = W; c1 <= std::min(max_p1, 2 * sms); c1 += W) {
            if (c1 > 1 && W > 1 && (c1 / W) < 3) continue;
            const double t = partition_cost(K, c1, sms);
            if (t < best) { best = t; bp1 = c1; }
        }
        P1 = bp1;
    }
    P1 = std::max(1, std::min(P1, max_p1));
    if (W > 1) { P1 = std::max(P1, 2 * W); P1 = (P1 + W - 1) / W * W; while (P1 > W && p->Wk / (P1 / W) < 3) P1 -= W; }
    if (P2 <= 0) P2 = 1;
    if (auto_groups && P1 > 3) {   // recursive levels above: take the first level the dynamic programme would choose
        std::vector<int> tmp;
        plan_upper_levels(P1 - 1, &tmp);
        P2 = tmp.empty() ? 1 : tmp[0];
    }
    if (W > 1) { P2 = std::max(P2, W); P2 = (P2 + W - 1) / W * W; while (P2 > W && (P1 / W - 1) / (P2 / W) < 2) P2 -= W; if ((P1 / W - 1) / (P2 / W) < 2) { set_err("partition too fine for this many ranks"); return B200_BAD_ARG; } }
    else { if (P1 < 4) P2 = 1; while (P2 > 1 && (P1 - 1) / P2 < 2) P2--; }
    free_levels(p);
    free_graphs(p);
    // chunk counts per level: Ps[0] over the K keyframes, Ps[l] over the Ps[l-1]-1 separators of the level below
    std::vector<int> Ps;
    if (P1 > 1) {
        Ps.push_back(P1);
        if (P2 > 1) Ps.push_back(P2);
        // single rank: all levels from the dynamic programme; sharded: level 1 is rank-aligned (P2 above), the levels above it
        // are computed redundantly on every rank (their inputs are all-gathered) and recurse the same way
        if (auto_groups && P2 > 1) plan_upper_levels(P2 - 1, &Ps);
    }
    p->P = P1; p->P2 = Ps.size() > 1 ? Ps[1] : 1; P2 = p->P2;
    const int nlev = (int)Ps.size();
    p->levels.resize(std::max(nlev, 1));
    // ---- level 0: keyframes
    {
        auto& l = p->levels[0];
        l.n_nodes = K; l.P = P1;
        l.begin.resize(P1 + 1);
        const int per = P1 / W;
        for (int r = 0; r < W; r++) {
            const int k0 = r * p->Wk, k1 = std::min(K, (r + 1) * p->Wk);
            for (int j = 0; j < per; j++) l.begin[r * per + j] = k0 + (int)((long long)j * (k1 - k0) / per);
        }
        l.begin[P1] = K;
        l.c_lo = p->rank * per; l.c_hi = l.c_lo + per;
        l.sep_st.resize(std::max(P1 - 1, 1));
        for (int c = 0; c + 1 < P1; c++) l.sep_st[c] = l.begin[c + 1] - 1;
        p->k_lo = l.begin[l.c_lo]; p->k_hi = l.begin[l.c_hi];
    }
    // ---- level li >= 1: the separators of level li-1 (level 1 is rank-aligned when the time axis is sharded)
    for (int li = 1; li < nlev; li++) {
        auto& lp = p->levels[li - 1];
        auto& l = p->levels[li];
        const int Pl = Ps[li];
        l.n_nodes = lp.P - 1; l.P = Pl;
        l.st = lp.sep_st; l.st.resize(l.n_nodes);
        l.begin.resize(Pl + 1);
        if (li == 1 && W > 1) {
            const int per1 = P1 / W, gpr = Pl / W;
            for (int r = 0; r < W; r++) {
                const int n0 = r * per1, n1 = std::min(P1 - 1, (r + 1) * per1);   // this rank's separators
                for (int j = 0; j < gpr; j++) l.begin[r * gpr + j] = n0 + (int)((long long)j * (n1 - n0) / gpr);
            }
            l.c_lo = p->rank * gpr; l.c_hi = l.c_lo + gpr;
        } else {
            for (int j = 0; j < Pl; j++) l.begin[j] = (int)((long long)j * l.n_nodes / Pl);
            l.c_lo = 0; l.c_hi = Pl;
        }
        l.begin[Pl] = l.n_nodes;
        l.sep_st.resize(std::max(Pl - 1, 1));
        for (int g = 0; g + 1 < Pl; g++) l.sep_st[g] = l.st[l.begin[g + 1] - 1];
    }
    const size_t Fsz = (size_t)(dm.nc + dm.hmax) * dm.nc, Ssz = (size_t)dm.hmax * dm.hmax;
    for (int li = 0; li < nlev; li++) {
        auto& l = p->levels[li];
        for (int c = 0; c < l.P; c++) {
            const int n = l.begin[c + 1] - l.begin[c];
            if (n < 2) { set_err("internal: partition produced a chunk with fewer than 2 nodes"); return B200_BAD_ARG; }
        }
        const size_t P = l.P;
        CK(dalloc(&l.d_begin, P + 1)); CK(dalloc(&l.d_sep_st, P)); CK(dalloc(&l.d_st, (size_t)std::max(l.n_nodes, 1)));
        CK(cudaMemcpyAsync(l.d_begin, l.begin.data(), sizeof(int) * (P + 1), cudaMemcpyHostToDevice, p->stream));
        CK(cudaMemcpyAsync(l.d_sep_st, l.sep_st.data(), sizeof(int) * l.sep_st.size(), cudaMemcpyHostToDevice, p->stream));
        if (!l.st.empty()) CK(cudaMemcpyAsync(l.d_st, l.st.data(), sizeof(int) * l.st.size(), cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        // upper levels on one GPU: few chunks, many idle SMs -> several CTAs per chunk sweep (row split, see Level::nsplit)
        l.nsplit = 1;
        const int max_split = std::max(1, std::min(B200_MAX_SPLIT, (dm.hmax + 15) / 16));   // (at most one CTA per 16-row tile of the arrow rows)
        if (li > 0 && (W == 1 || li >= 2)) { while (l.nsplit < max_split && l.P * (l.nsplit + 1) <= 148) l.nsplit++; }
        CK(dalloc(&l.F, (P + 1) * l.nsplit * 2 * Fsz)); CK(dalloc(&l.S, (P + 1) * Ssz));
        CK(dalloc(&l.Dred, P * dm.ntp * dm.ntp)); CK(dalloc(&l.Ored, P * dm.nc * dm.nc));
        CK(dalloc(&l.Ared, P * dm.ns1 * dm.ntp)); CK(dalloc(&l.Sred, P * dm.ns1 * dm.ns1));
        Level& d = l.dev;
        if (li == 0) {   // keyframes: the assembled blocks, or the reduced ones when the local dofs are pre-eliminated
            d.Dsrc = p->use_le ? p->d_Dc : p->d_D; d.Osrc = p->d_O; d.Asrc = p->use_le ? p->d_Ac : p->d_A; d.Ssrc = p->use_le ? p->d_Sc : p->d_S; d.st = nullptr;
        }
        else { auto& lp = p->levels[li - 1]; d.Dsrc = lp.Dred; d.Osrc = lp.Ored; d.Asrc = lp.Ared; d.Ssrc = lp.Sred; d.st = l.d_st; }
        d.begin = l.d_begin; d.P = l.P; d.F = l.F; d.S = l.S;
        d.Dred = l.Dred; d.Ored = l.Ored; d.Ared = l.Ared; d.Sred = l.Sred; d.sep_st = l.d_sep_st; d.dense_o = (li > 0); d.nsplit = l.nsplit;
    }
    if (nlev == 0) { auto& l = p->levels[0]; l.n_nodes = K; l.P = 1; p->k_lo = 0; p->k_hi = K; }
    {   // top level: a single chunk (no separators) over whatever the levels below left
        Level& t = p->top;
        t = Level{};
        int tn = K;
        if (nlev == 0) { t.Dsrc = p->use_le ? p->d_Dc : p->d_D; t.Osrc = p->d_O; t.Asrc = p->use_le ? p->d_Ac : p->d_A; t.Ssrc = p->use_le ? p->d_Sc : p->d_S; t.st = nullptr; }
        else { auto& l = p->levels[nlev - 1]; t.Dsrc = l.Dred; t.Osrc = l.Ored; t.Asrc = l.Ared; t.Ssrc = l.Sred; t.st = l.d_sep_st; tn = l.P - 1; }
        const int tb[2] = {0, tn};
        CK(cudaMemcpyAsync(p->d_top_begin, tb, sizeof(tb), cudaMemcpyHostToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        t.begin = p->d_top_begin; t.P = 1; t.F = p->d_Ftop; t.S = p->d_Stop; t.dense_o = (nlev > 0); t.nsplit = 1;
        if (t.dense_o && dm.pipe == 1) t.nsplit = std::max(1, std::min(B200_MAX_SPLIT, (dm.hmax + 15) / 16));   // the single top-level chunk: its arrow rows over several CTAs as well
    }
    if (W > 1) {
        if (p->d_stageS) cudaFree(p->d_stageS);
        if (p->d_stageF) cudaFree(p->d_stageF);
        if (p->d_pack) cudaFree(p->d_pack);
        p->d_stageS = p->d_stageF = p->d_pack = nullptr;
        CK(dalloc(&p->d_stageS, (size_t)W * Ssz)); CK(dalloc(&p->d_stageF, (size_t)W * 2 * Fsz));
        p->pack_doubles = (size_t)dm.ntp * dm.ntp + (size_t)dm.ns1 * dm.ntp + (size_t)dm.ns1 * dm.ns1;
        CK(dalloc(&p->d_pack, (size_t)P2 * p->pack_doubles));
    }
    p->have_lin = false;
    return B200_OK;
}

b200_status b200_set_partition(b200_problem* p, int32_t n_chunks, int32_t n_groups) {
    if (!p) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    return set_partition(p, n_chunks, n_groups);
}
// n_chunks > 0 with the separators eliminated by one sequential top-level sweep (two-level scheme)
b200_status b200_set_chunks(b200_problem* p, int32_t n_chunks) { return b200_set_partition(p, n_chunks, n_chunks > 0 ? 1 : 0); }
int32_t b200_get_chunks(const b200_problem* p) { return p ? p->P : 0; }
int32_t b200_get_groups(const b200_problem* p) { return p ? p->P2 : 0; }
int32_t b200_get_levels(const b200_problem* p, int32_t* sizes, int32_t max_levels) {
    if (!p || p->P <= 1) return 0;
    const int n = (int)p->levels.size();
    if (sizes) for (int i = 0; i < n && i < max_levels; i++) sizes[i] = p->levels[i].P;
    return n;
}

int32_t b200_time_value_dim(const b200_problem* p) { return p->tvd; }
int32_t b200_static_value_dim(const b200_problem* p) { return p->svd; }
int32_t b200_time_tangent_dim(const b200_problem* p) { return p->dm.nt; }
int32_t b200_static_tangent_dim(const b200_problem* p) { return p->dm.ns; }
int64_t b200_launch_count(const b200_problem* p) { return p->launches; }

b200_status b200_set_values(b200_problem* p, const double* tv, const double* sv) {
    if (!p || !tv || (p->svd && !sv)) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    const long long n = (long long)p->K * p->tvd;
    CK(cudaMemcpyAsync(p->d_stage, tv, sizeof(double) * n, cudaMemcpyHostToDevice, p->stream));
    B200_LAUNCH(transpose_in_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, p->stream, p->d_stage, p->K, p->tvd, p->dm.Kld, p->d_tv);
    p->launches++;
    if (p->svd) CK(cudaMemcpyAsync(p->d_sv, sv, sizeof(double) * p->svd, cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    p->have_lin = false;
    p->err_stale = true;   // the LM state's cost refers to the previous values: b200_lm_iterate re-evaluates it
    return B200_OK;
}

b200_status b200_get_values(b200_problem* p, double* tv, double* sv) {
    if (!p || !tv || (p->svd && !sv)) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    const long long n = (long long)p->K * p->tvd;
    B200_LAUNCH(transpose_out_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, p->stream, p->d_tv, p->K, p->tvd, p->dm.Kld, p->d_stage);
    p->launches++;
    CK(cudaMemcpyAsync(tv, p->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream));
    if (p->svd) CK(cudaMemcpyAsync(sv, p->d_sv, sizeof(double) * p->svd, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ pipeline stages (async)
static void launch_error(b200_problem* p, const double* tv, const double* sv, double* ebuf, int scal_idx, const int* skip = nullptr) {
    dim3 grid((p->max_inst + 127) / 128, p->n_slots);
    B200_LAUNCH(linearize_kernel<false>, grid, dim3(128), 0, p->stream, p->d_slots, tv, sv, p->d_cst, p->dm.Kld, (double*)nullptr,
                (double*)nullptr, ebuf, p->k_lo, p->k_hi, p->k_lo, skip);
    B200_LAUNCH(reduce_sum_kernel, dim3(B200_REDUCE_BLOCKS), dim3(B200_REDUCE_THREADS), 0, p->stream, ebuf, p->ne, p->sb.scal + scal_idx, skip, p->d_red_partial, p->d_red_ticket);
    p->launches += 2;
}

static void launch_linearize(b200_problem* p) {
    B200_NVTX("linearize + assemble");
    dim3 grid((p->max_inst + 127) / 128, p->n_slots);
    // Jacobians are needed one keyframe to the left of the window (the IMU factor k_lo-1 -> k_lo feeds D[k_lo])
    B200_LAUNCH(linearize_kernel<true>, grid, dim3(128), 0, p->stream, p->d_slots, p->d_tv, p->d_sv, p->d_cst, p->dm.Kld, p->d_J, p->d_r,
                p->d_e, std::max(p->k_lo - 1, 0), p->k_hi, p->k_lo, (const int*)nullptr);
    p->launches++;
    // one extra keyframe on the left: its O block (coupling k_lo-1 -> k_lo) is the left-separator coupling of this window
    const int ka = std::max(p->k_lo - 1, 0);
    B200_LAUNCH(assemble_kernel, dim3(p->k_hi - ka), dim3(512), std::max(p->smem_asm[0], p->smem_asm[1]), p->stream, p->dm,
                p->d_slots, p->asm_plan, p->d_J, p->d_r, ka, 0, p->K, p->d_D, p->d_O, p->d_A, p->d_S);
    p->launches++;
    if (p->d_le_fresh) cudaMemsetAsync(p->d_le_fresh, 1, sizeof(int), p->stream);   // new linearization: the next local elimination writes everything
}

// Collectives.  A "group" is a run of exchanges issued together (one ncclGroupStart / ncclGroupEnd with the in-library NCCL).
// With phase timing on, every group is bracketed by CUDA events and waited for, so that its device time can be attributed
// (b200_comm_ms); the production path never synchronises here.
static b200_status comm_begin(b200_problem* p) {
    if (p->time_phases) CK(cudaEventRecord(p->cev[0], p->stream));
#ifndef B200_USE_EMU
    if (p->own_nccl) nccl_api().GroupStart();
#endif
    return B200_OK;
}
static b200_status comm_end(b200_problem* p) {
#ifndef B200_USE_EMU
    if (p->own_nccl && nccl_api().GroupEnd() != ncclSuccess) { set_err("ncclGroupEnd failed"); return B200_NCCL_ERROR; }
#endif
    p->comm_calls++;
    if (p->time_phases) {
        CK(cudaEventRecord(p->cev[1], p->stream));
        CK(cudaEventSynchronize(p->cev[1]));
        float ms = 0; cudaEventElapsedTime(&ms, p->cev[0], p->cev[1]); p->comm_ms += ms;
    }
    return B200_OK;
}
static b200_status allgather(b200_problem* p, void* buf, size_t bytes_per_rank) {
    if (!p->ag_fn) { set_err("world_size > 1 but no collective registered (b200_comm_set / b200_nccl_init)"); return B200_NCCL_ERROR; }
    if (p->ag_fn(p->comm_ctx, buf, (uint64_t)bytes_per_rank, p->stream)) { set_err("allgather failed"); return B200_NCCL_ERROR; }
    return B200_OK;
}

// the reduced blocks now belong to the current linearization (a skipped trial leaves the flag alone: it wrote nothing)
__global__ void clear_flag_kernel(int* flag, const int* __restrict__ skip) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && !(skip && *skip)) *flag = 0;
}

// pack / unpack the level-1 source blocks of the group separators (D, A, S of the reduced system) for the all-gather
__global__ void pack_groups_kernel(Dims dm, Level lv1, const double* Dred, const double* Ared, const double* Sred, double* pack, long long stride,
                                   int g_first, int unpack, int skip_lo, int skip_hi) {
    const int g = g_first + blockIdx.x;
    if (g >= lv1.P - 1) return;
    if (unpack && g >= skip_lo && g < skip_hi) return;   // own groups: already in place
    const long long q = lv1.begin[g + 1] - 1;
    const long long nD = (long long)dm.ntp * dm.ntp, nA = (long long)dm.ns1 * dm.ntp, nS = (long long)dm.ns1 * dm.ns1;
    double* pk = pack + (long long)g * stride;
    double* D = const_cast<double*>(Dred) + q * nD; double* A = const_cast<double*>(Ared) + q * nA; double* S = const_cast<double*>(Sred) + q * nS;
    for (long long i = threadIdx.x; i < nD + nA + nS; i += blockDim.x) {
        double* x = (i < nD) ? D + i : (i < nD + nA ? A + (i - nD) : S + (i - nD - nA));
        if (unpack) *x = pk[i]; else pk[i] = *x;
    }
}

// one level: sweep (CTA per chunk) + arrow Schur complement (CTA per chunk)
static void launch_level_factor(b200_problem* p, const Level& lv, int c_lo, int c_hi) {
    const Dims& dm = p->ds;
    const dim3 sgrid(c_hi - c_lo, lv.nsplit > 0 ? lv.nsplit : 1);
    if (lv.dense_o) B200_LAUNCH(chain_sweep_kernel<true>, sgrid, dim3(B200_SOLVE_THREADS), p->smem_sweep, p->stream, dm, lv, p->sb, c_lo);
    else B200_LAUNCH(chain_sweep_kernel<false>, sgrid, dim3(B200_SOLVE_THREADS), p->smem_sweep, p->stream, dm, lv, p->sb, c_lo);
    // few chunks (upper levels): split the output tiles of a chunk's Schur complement over several CTAs
    const int nrt = (dm.hmax + 31) >> 5, ntmax = nrt * (nrt + 1) / 2;   // 32 x 32 tiles of the lower triangle of the complement (21 for 165 rows)
    if (p->schur_tile_ok && (c_hi - c_lo) * ntmax <= 8 * 148) {
        // few chunks: one output tile per CTA, three CTAs per SM (kernels_solve.cuh: chain_schur_tile_kernel)
        B200_LAUNCH(chain_schur_tile_kernel, dim3(c_hi - c_lo, ntmax), dim3(B200_SCHUR_TILE_THREADS), p->smem_schur_tile, p->stream, dm, lv, p->sb, c_lo,
                    p->schur_tile_ldy);
    } else {
        const int nsplit = std::max(1, std::min(148 / std::max(1, c_hi - c_lo), ntmax));
        B200_LAUNCH(chain_schur_kernel, dim3(c_hi - c_lo, nsplit), dim3(B200_SCHUR_THREADS), p->smem_schur, p->stream, dm, lv, p->sb, c_lo, p->schur_ldy,
                    p->schur_kslab);
    }
    p->launches += 2;
}
static void launch_level_backsub(b200_problem* p, const b200_problem::HostLevel& l) {
    const Dims& dm = p->ds;
    int max_nodes = 1;
    for (int c = l.c_lo; c < l.c_hi; c++) max_nodes = std::max(max_nodes, l.begin[c + 1] - l.begin[c]);
    B200_LAUNCH(chain_backsub_prepass_kernel, dim3(max_nodes, l.c_hi - l.c_lo), dim3(B200_PREPASS_THREADS), 0, p->stream, dm, l.dev, p->sb, l.c_lo);
    B200_LAUNCH(chain_backsub_kernel, dim3(l.c_hi - l.c_lo), dim3(B200_SOLVE_THREADS), p->smem_backsub, p->stream, dm, l.dev, p->sb, l.c_lo);
    p->launches += 2;
}

// (H + lambda I) delta = -g with lambda = *sb.lam (device scalar): local elimination, nested partitioned factorisation,
// back-substitution, all on p->stream.  tev != nullptr: phase events of this trial slot (direct launches only, not under capture).
enum { TE_SOLVE0 = 0, TE_SWEEP0, TE_SWEEP1, TE_TOP0, TE_TOP1, TE_BACK1, TE_SOLVE1, TE_TRIAL1, TE_L0DONE, TE_L1DONE, TE_COUNT };
static b200_status launch_solve(b200_problem* p, cudaEvent_t* tev) {
    B200_NVTX("solve");
    CK(cudaMemsetAsync(p->sb.status, 0, sizeof(int), p->stream));
    const Dims& dm = p->ds;
    const int W = p->world, r = p->rank;
    const size_t Fsz = (size_t)(dm.nc + dm.hmax) * dm.nc, Ssz = (size_t)dm.hmax * dm.hmax;
    const int nlev = (p->P > 1) ? (int)p->levels.size() : 0;
    if (p->use_le) {
        B200_LAUNCH(local_elim_kernel, dim3(p->k_hi - p->k_lo), dim3(B200_LE_THREADS), p->smem_le, p->stream, p->le, p->k_lo, p->sb.lam, p->sb.status, p->sb.skip,
                    (const int*)p->d_le_fresh);
        B200_LAUNCH(clear_flag_kernel, dim3(1), dim3(32), 0, p->stream, p->d_le_fresh, p->sb.skip);
        p->launches += 2;
    }
    if (tev) CK(cudaEventRecord(tev[TE_SWEEP0], p->stream));
    if (nlev >= 1) {
        B200_NVTX("solve: level 0 (keyframes)");
        auto& l0 = p->levels[0];
        B200_LAUNCH(chain_sweep_kernel<false>, dim3(l0.c_hi - l0.c_lo), dim3(B200_SOLVE_THREADS), p->smem_sweep, p->stream, dm, l0.dev, p->sb, l0.c_lo);
        p->launches++;
        if (tev) CK(cudaEventRecord(tev[TE_SWEEP1], p->stream));
        B200_LAUNCH(chain_schur_kernel, dim3(l0.c_hi - l0.c_lo), dim3(B200_SCHUR_THREADS), p->smem_schur, p->stream, dm, l0.dev, p->sb, l0.c_lo,
                    p->schur_ldy, p->schur_kslab);
        p->launches++;
        int red_lo = l0.c_lo, red_hi = std::min(l0.c_hi, l0.P - 1);
        if (W > 1) {
            // the separator that closes this rank's window needs the Schur complement / fill of the NEXT rank's first chunk
            CK(cudaMemcpyAsync(p->d_stageS + (size_t)r * Ssz, l0.S + (size_t)l0.c_lo * Ssz, sizeof(double) * Ssz, cudaMemcpyDeviceToDevice, p->stream));
            CK(cudaMemcpyAsync(p->d_stageF + (size_t)r * 2 * Fsz, l0.F + (size_t)l0.c_lo * 2 * Fsz, sizeof(double) * 2 * Fsz, cudaMemcpyDeviceToDevice, p->stream));
            b200_status st = comm_begin(p); if (st != B200_OK) return st;
            st = allgather(p, p->d_stageS, sizeof(double) * Ssz); if (st != B200_OK) return st;
            st = allgather(p, p->d_stageF, sizeof(double) * 2 * Fsz); if (st != B200_OK) return st;
            st = comm_end(p); if (st != B200_OK) return st;
            if (r + 1 < W) {
                CK(cudaMemcpyAsync(l0.S + (size_t)l0.c_hi * Ssz, p->d_stageS + (size_t)(r + 1) * Ssz, sizeof(double) * Ssz, cudaMemcpyDeviceToDevice, p->stream));
                CK(cudaMemcpyAsync(l0.F + (size_t)l0.c_hi * 2 * Fsz, p->d_stageF + (size_t)(r + 1) * 2 * Fsz, sizeof(double) * 2 * Fsz, cudaMemcpyDeviceToDevice, p->stream));
            }
            // the coupling block between the previous rank's closing separator and this rank's first separator comes
            // from this rank's own first chunk: recompute that separator's reduced O here (D/A/S parts are unused)
            if (r > 0) red_lo = l0.c_lo - 1;
        }
        if (red_hi > red_lo) {
            B200_LAUNCH(chain_reduce_kernel, dim3(red_hi - red_lo, 8), dim3(256), 0, p->stream, dm, l0.dev, red_lo, p->sb.skip);
            p->launches++;
        }
    } else if (tev) CK(cudaEventRecord(tev[TE_SWEEP1], p->stream));
    if (tev) { CK(cudaEventRecord(tev[TE_L0DONE], p->stream)); if (nlev < 2) CK(cudaEventRecord(tev[TE_L1DONE], p->stream)); }
    for (int li = 1; li < nlev; li++) {
        B200_NVTX(li == 1 ? "solve: level 1 (separators, rank-aligned)" : "solve: level >= 2 (replicated)");
        auto& l0 = p->levels[0];
        auto& l1 = p->levels[li];
        launch_level_factor(p, l1.dev, l1.c_lo, l1.c_hi);
        if (W > 1 && li == 1) {
            const int gpr = l1.P / W;
            // (the packed blocks are level-1 SOURCES: ready before the level-1 sweep; packed here so that the three exchanges form one group)
            B200_LAUNCH(pack_groups_kernel, dim3(gpr), dim3(256), 0, p->stream, dm, l1.dev, l0.Dred, l0.Ared, l0.Sred, p->d_pack,
                        (long long)p->pack_doubles, l1.c_lo, 0, 0, 0);
            b200_status st = comm_begin(p); if (st != B200_OK) return st;
            st = allgather(p, l1.F, sizeof(double) * 2 * Fsz * gpr); if (st != B200_OK) return st;
            st = allgather(p, l1.S, sizeof(double) * Ssz * gpr); if (st != B200_OK) return st;
            st = allgather(p, p->d_pack, sizeof(double) * p->pack_doubles * gpr); if (st != B200_OK) return st;
            st = comm_end(p); if (st != B200_OK) return st;
            B200_LAUNCH(pack_groups_kernel, dim3(l1.P), dim3(256), 0, p->stream, dm, l1.dev, l0.Dred, l0.Ared, l0.Sred, p->d_pack,
                        (long long)p->pack_doubles, 0, 1, l1.c_lo, l1.c_hi);
            p->launches += 2;
        }
        // (few separators up here: 32 CTAs each, so that every thread handles one batch of four entries)
        B200_LAUNCH(chain_reduce_kernel, dim3(l1.P - 1, (l1.P - 1) * 32 <= 2048 ? 32 : 8), dim3(256), 0, p->stream, dm, l1.dev, 0, p->sb.skip);
        p->launches++;
        if (tev && li == 1) CK(cudaEventRecord(tev[TE_L1DONE], p->stream));
    }
    if (tev) CK(cudaEventRecord(tev[TE_TOP0], p->stream));
    {
    B200_NVTX("solve: top level + static system");
    launch_level_factor(p, p->top, 0, 1);
    B200_LAUNCH(top_finish_kernel, dim3(1), dim3(B200_SOLVE_THREADS), p->smem_backsub, p->stream, dm, p->top, p->sb);
    p->launches++;
    }
    if (tev) CK(cudaEventRecord(tev[TE_TOP1], p->stream));
    B200_NVTX("solve: back-substitution");
    for (int li = nlev - 1; li >= 0; li--) launch_level_backsub(p, p->levels[li]);
    if (p->use_le) {
        B200_LAUNCH(local_backsub_kernel, dim3(p->k_hi - p->k_lo), dim3(128), 0, p->stream, p->le, p->k_lo, p->k_hi, p->sb.delta, p->sb.dstat, p->d_delta,
                    p->sb.status, p->sb.skip);
        p->launches++;
    }
    if (tev) CK(cudaEventRecord(tev[TE_BACK1], p->stream));
    if (W > 1) {  // every rank needs every keyframe's step: the values are replicated
        b200_status st = comm_begin(p); if (st != B200_OK) return st;
        st = allgather(p, p->d_delta, sizeof(double) * (size_t)p->Wk * p->dm.ntp); if (st != B200_OK) return st;
        st = comm_end(p); if (st != B200_OK) return st;
    }
    return B200_OK;
}

static b200_status fetch_scalars(b200_problem* p) {
    CK(cudaMemcpyAsync(p->h_scal, p->sb.scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(p->h_scal + SC_COUNT, p->sb.status, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    return B200_OK;
}
static int host_status(const b200_problem* p) { return *reinterpret_cast<const int*>(p->h_scal + SC_COUNT); }

static b200_status allreduce_scal(b200_problem* p) {
    if (p->world <= 1) return B200_OK;
    if (!p->ar_fn) { set_err("world_size > 1 but no collective registered"); return B200_NCCL_ERROR; }
    b200_status st = comm_begin(p); if (st != B200_OK) return st;
    if (p->ar_fn(p->comm_ctx, p->sb.scal, SC_COUNT, p->stream)) { set_err("allreduce failed"); return B200_NCCL_ERROR; }
    return comm_end(p);
}

// upload the LM state mirror (pinned) to the device; the caller has synchronised the stream since the mirror was last in flight
static b200_status push_lm_state(b200_problem* p) {
    LmDev& h = *p->h_lm;
    h.lambda = p->lambda; h.lambda_try = p->lambda; h.err = p->err; h.new_err = p->err;
    h.iterations = p->iterations; h.total_trials = p->total_trials; h.inner = 0; h.status = B200_OK; h.done = 0; h.accepted = 0;
    CK(cudaMemcpyAsync(p->d_lm, p->h_lm, sizeof(LmDev), cudaMemcpyHostToDevice, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return B200_OK;
}
// set the damping of the next solve from the host (sharded / speculative path, debug_solve)
static b200_status push_lambda_try(b200_problem* p, double lam) {
    p->h_lm->lambda_try = lam;
    CK(cudaMemcpyAsync(&p->d_lm->lambda_try, &p->h_lm->lambda_try, sizeof(double), cudaMemcpyHostToDevice, p->stream));
    return B200_OK;
}

b200_status b200_error(b200_problem* p, double* error) {
    if (!p || !error) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaMemsetAsync(p->sb.scal, 0, sizeof(double) * SC_COUNT, p->stream));
    launch_error(p, p->d_tv, p->d_sv, p->d_e2, SC_CUR_ERR);
    b200_status st = allreduce_scal(p); if (st != B200_OK) return st;
    st = fetch_scalars(p); if (st != B200_OK) return st;
    *error = p->h_scal[SC_CUR_ERR];
    return B200_OK;
}

b200_status b200_lm_begin(b200_problem* p, const b200_lm_params* prm, double* initial_error) {
    if (!p || !prm) return B200_BAD_ARG;
    p->lm = *prm;
    p->lambda = prm->gauss_newton ? 0.0 : prm->lambda_initial;
    p->iterations = 0; p->total_trials = 0; p->have_lin = false;
    double e;
    b200_status st = b200_error(p, &e);
    if (st != B200_OK) return st;
    p->err = e; p->err_stale = false; p->lm_started = true;
    if (initial_error) *initial_error = e;
    return push_lm_state(p);
}

// retract (values + delta -> try buffers): time variables of every keyframe, then the statics
static void launch_retract(b200_problem* p, const double* delta, const double* dstat, const int* skip = nullptr) {
    B200_NVTX("retract");
    B200_LAUNCH(retract_kernel, dim3((p->K + 127) / 128, 1, p->ntv), dim3(128), 0, p->stream, p->d_tvars, p->ntv, p->d_svars, p->nsv, p->K,
                p->dm.Kld, p->d_tv, p->d_sv, delta, p->dm.ntp, dstat, p->d_tv_try, p->d_sv_try, skip);
    p->launches++;
    if (p->nsv) {
        B200_LAUNCH(retract_kernel, dim3((p->nsv + 127) / 128, 2, 1), dim3(128), 0, p->stream, p->d_tvars, 0, p->d_svars, p->nsv, 0, p->dm.Kld,
                    p->d_tv, p->d_sv, delta, p->dm.ntp, dstat, p->d_tv_try, p->d_sv_try, skip);
        p->launches++;
    }
}

// outcome of one lambda trial (gtsam @4.2 LevenbergMarquardtOptimizer::tryLambda): accept the step / stop searching
struct TrialVerdict { bool solved, step_ok, stop; };
__host__ __device__ inline TrialVerdict judge_trial(const b200_lm_params& L, double cur_err, bool solved, double oldLin, double newLin, double newError) {
    TrialVerdict v{solved, false, false};
    if (!solved) return v;
    const double linChange = oldLin - newLin;
    if (L.gauss_newton) v.step_ok = true;
    else if (linChange >= 0) {
        const double costChange = cur_err - newError;
        if (linChange > 2.220446049250313e-16 * oldLin) v.step_ok = (costChange / linChange) > L.min_model_fidelity;
        if (fabs(costChange) < L.relative_error_tol * cur_err) v.stop = true;
    }
    return v;
}

// The damping / accept-reject logic of one lambda trial ON THE DEVICE (north_star subsystem 4): one thread reads the three cost
// scalars of the trial and the pivot status, applies gtsam's tryLambda rules (judge_trial + increaseLambda / decreaseLambda) to
// the LM state in device memory and, when the ladder has ended, raises `done` so that the trials already queued behind this
// one return at once.  The host reads the state once per iterate().
__global__ void lm_judge_kernel(LmDev* st, const double* __restrict__ scal, b200_lm_params L) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || st->done) return;
    const bool solved = scal[SC_FAIL] == 0.0;   // number of ranks whose factorisation met a non-positive pivot (all-reduced)
    const TrialVerdict v = judge_trial(L, st->err, solved, scal[SC_LIN_ERR], scal[SC_MODEL_ERR], scal[SC_NEW_ERR]);
    st->inner++;
    st->status = solved ? B200_OK : B200_INDETERMINATE_SYSTEM;
    if (v.step_ok) {
        st->new_err = scal[SC_NEW_ERR]; st->err = scal[SC_NEW_ERR];
        if (!L.gauss_newton) { const double l = st->lambda / L.lambda_factor; st->lambda = l > L.lambda_lower_bound ? l : L.lambda_lower_bound; }
        st->iterations++; st->total_trials++;
        st->status = B200_OK; st->accepted = 1; st->done = 1;
    } else if (!v.stop) {
        st->lambda *= L.lambda_factor; st->total_trials++;
        if (st->lambda >= L.lambda_upper_bound) { st->status = B200_LAMBDA_MAXED; st->done = 1; }
        else if (L.gauss_newton) { st->status = B200_INDETERMINATE_SYSTEM; st->done = 1; }
    } else st->done = 1;
    st->lambda_try = st->lambda;
}
__global__ void lm_begin_iter_kernel(LmDev* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    st->done = 0; st->accepted = 0; st->inner = 0; st->lambda_try = st->lambda;
}
// pivot status of this rank as a 0/1 entry of the all-reduced scalar block: a failure on ANY rank marks the trial unsolved on all
__global__ void status_to_scal_kernel(const int* __restrict__ status, double* __restrict__ scal) {
    if (threadIdx.x == 0 && blockIdx.x == 0) scal[SC_FAIL] = (*status != 0) ? 1.0 : 0.0;
}

// one lambda trial: solve at *sb.lam, linear-model error of the step, retraction into the try buffers, cost of the try values;
// device_judge: followed by lm_judge_kernel.  Everything is launched on p->stream; no host decision in between.
static b200_status launch_trial(b200_problem* p, cudaEvent_t* tev, bool device_judge) {
    if (tev) CK(cudaEventRecord(tev[TE_SOLVE0], p->stream));
    b200_status st = launch_solve(p, tev);
    if (st != B200_OK) return st;
    if (tev) CK(cudaEventRecord(tev[TE_SOLVE1], p->stream));
    dim3 grid((unsigned)((p->max_rows + 127) / 128), p->n_slots);   // one thread per factor row
    B200_LAUNCH(model_error_rows_kernel, grid, dim3(128), 0, p->stream, p->d_slots, p->d_J, p->d_r, p->d_delta, p->dm.ntp, p->sb.dstat, p->d_e2r,
                p->k_lo, p->k_hi, p->sb.skip);
    B200_LAUNCH(reduce_sum_kernel, dim3(B200_REDUCE_BLOCKS), dim3(B200_REDUCE_THREADS), 0, p->stream, p->d_e2r, p->nr, p->sb.scal + SC_MODEL_ERR, p->sb.skip, p->d_red_partial, p->d_red_ticket);
    p->launches += 2;
    launch_retract(p, p->d_delta, p->sb.dstat, p->sb.skip);
    launch_error(p, p->d_tv_try, p->d_sv_try, p->d_e2, SC_NEW_ERR, p->sb.skip);
    // rank-local linear.error(0) lives outside the all-reduced block (scal[SC_COUNT]); it is copied in at every trial
    CK(cudaMemcpyAsync(p->sb.scal + SC_LIN_ERR, p->sb.scal + SC_COUNT, sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    B200_LAUNCH(status_to_scal_kernel, dim3(1), dim3(32), 0, p->stream, p->sb.status, p->sb.scal);
    p->launches++;
    st = allreduce_scal(p); if (st != B200_OK) return st;   // (sharded: the same scalars, bit for bit, on every rank of the group)
    if (device_judge) {
        B200_LAUNCH(lm_judge_kernel, dim3(1), dim3(32), 0, p->stream, p->d_lm, p->sb.scal, p->lm);
        p->launches++;
    }
    if (tev) CK(cudaEventRecord(tev[TE_TRIAL1], p->stream));
    return B200_OK;
}

static void free_graphs(b200_problem* p) {
#ifndef B200_USE_EMU
    for (int i = 0; i < 2; i++) if (p->trial_graph[i]) { cudaGraphExecDestroy(p->trial_graph[i]); p->trial_graph[i] = nullptr; }
#endif
    p->graph_launches = 0;
}

// the trial of the current buffer parity as an instantiated CUDA graph (captured once per parity and partition)
static b200_status trial_graph(b200_problem* p, int parity) {
#ifndef B200_USE_EMU
    if (p->trial_graph[parity]) return B200_OK;
    const long long l0 = p->launches;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
    b200_status st = launch_trial(p, nullptr, true);
    cudaError_t e = cudaStreamEndCapture(p->stream, &g);
    if (st != B200_OK) { if (g) cudaGraphDestroy(g); return st; }
    if (e != cudaSuccess) { set_err(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); return B200_CUDA_ERROR; }
    e = cudaGraphInstantiate(&p->trial_graph[parity], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { set_err(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); return B200_CUDA_ERROR; }
    p->graph_launches = p->launches - l0;   // kernels inside one trial graph
    p->launches = l0;
#endif
    return B200_OK;
}

static void accumulate_phase_ms(b200_problem* p, cudaEvent_t* tev) {
    float ms;
    cudaEventElapsedTime(&ms, tev[TE_SOLVE0], tev[TE_SOLVE1]); p->phase_ms[1] += ms;
    cudaEventElapsedTime(&ms, tev[TE_SOLVE1], tev[TE_TRIAL1]); p->phase_ms[2] += ms;
    if (p->P > 1) {
        cudaEventElapsedTime(&ms, tev[TE_SWEEP0], tev[TE_SWEEP1]); p->phase_ms[3] += ms; p->phase_ms[4] += 1;
        cudaEventElapsedTime(&ms, tev[TE_TOP1], tev[TE_BACK1]); p->phase_ms[7] += ms;
    }
    cudaEventElapsedTime(&ms, tev[TE_TOP0], tev[TE_TOP1]); p->phase_ms[5] += ms; p->phase_ms[6] += 1;
    cudaEventElapsedTime(&ms, tev[TE_SOLVE0], tev[TE_SWEEP0]); p->phase_ms[8] += ms;
    cudaEventElapsedTime(&ms, tev[TE_SWEEP1], tev[TE_L0DONE]); p->phase_ms[9] += ms;
    cudaEventElapsedTime(&ms, tev[TE_L0DONE], tev[TE_L1DONE]); p->phase_ms[10] += ms;
    cudaEventElapsedTime(&ms, tev[TE_L1DONE], tev[TE_TOP0]); p->phase_ms[11] += ms;
}

// == LevenbergMarquardtOptimizer::iterate() (gtsam @4.2: iterate / tryLambda / increaseLambda / decreaseLambda)
// Single rank: the whole ladder runs on the device -- `spec_trials` trials are queued back to back (each one a CUDA graph), the
// accept / reject / damping decisions are taken by lm_judge_kernel, the host synchronises once per round and reads the state.
// Sharded / speculative (b200_comm_set / b200_spec_set): the ranks exchange the trial scalars through the registered
// collectives and every rank applies the same rules on the host.  With lambda speculation (G groups) every round evaluates the
// next G values of the lambda ladder at once, one per group, and the verdicts are then read in ladder order exactly as the
// sequential loop would have met them: the accepted step, the final lambda and the trial count are those of the sequential algorithm.
b200_status b200_lm_iterate(b200_problem* p, b200_lm_state* state) {
    if (!p || !p->lm_started) return B200_BAD_ARG;
    B200_NVTX("b200_lm_iterate");
    CK(cudaSetDevice(p->device));
    const b200_lm_params& L = p->lm;
    if (p->err_stale) {   // b200_set_values since the last cost evaluation: refresh the cost the trials are judged against
        double e; b200_status st = b200_error(p, &e); if (st != B200_OK) return st;
        p->err = e; p->err_stale = false;
        st = push_lm_state(p); if (st != B200_OK) return st;
    }
    if (p->time_phases) CK(cudaEventRecord(p->ev[0], p->stream));
    CK(cudaMemsetAsync(p->sb.scal, 0, sizeof(double) * SC_COUNT, p->stream));
    launch_linearize(p);
    B200_LAUNCH(reduce_sum_kernel, dim3(B200_REDUCE_BLOCKS), dim3(B200_REDUCE_THREADS), 0, p->stream, p->d_e, p->ne, p->sb.scal + SC_COUNT, (const int*)nullptr, p->d_red_partial, p->d_red_ticket);
    p->launches++;
    if (p->time_phases) CK(cudaEventRecord(p->ev[1], p->stream));
    p->have_lin = true;
    const int G = (p->spec_ag && !L.gauss_newton) ? p->spec_G : 1;
    const int T = p->world;
    int status = B200_OK, inner = 0;
    if (G == 1) {
        // ------------------------------------------------------------ device-resident ladder (single rank, or one time-sharded
        // group: after the all-reduce every rank holds the same scalars, so every rank's judge kernel takes the same decision)
        p->sb.skip = &p->d_lm->done;
        B200_LAUNCH(lm_begin_iter_kernel, dim3(1), dim3(32), 0, p->stream, p->d_lm);
        p->launches++;
        const int nq = L.gauss_newton ? 1 : std::max(1, std::min(p->spec_trials, 4));
#ifndef B200_USE_EMU
        const bool graph = p->use_graph && !p->time_phases && T == 1;
#else
        const bool graph = false;
#endif
        bool first = true;
        while (true) {
            for (int q = 0; q < nq; q++) {
                if (graph) {
#ifndef B200_USE_EMU
                    b200_status st = trial_graph(p, p->parity); if (st != B200_OK) { p->sb.skip = nullptr; return st; }
                    CK(cudaGraphLaunch(p->trial_graph[p->parity], p->stream));
                    p->launches += p->graph_launches;
#endif
                } else {
                    b200_status st = launch_trial(p, p->time_phases ? p->tev[q] : nullptr, true);
                    if (st != B200_OK) { p->sb.skip = nullptr; return st; }
                }
            }
            CK(cudaMemcpyAsync(p->h_lm, p->d_lm, sizeof(LmDev), cudaMemcpyDeviceToHost, p->stream));
            CK(cudaMemcpyAsync(p->h_scal, p->sb.scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, p->stream));   // (b200_debug_get_scalars)
            CK(cudaStreamSynchronize(p->stream));
            CK(cudaGetLastError());
            p->host_syncs++;
            if (p->time_phases) {
                float ms;
                if (first) { cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]); p->phase_ms[0] += ms; }
                for (int q = 0; q < nq && q < p->h_lm->inner - inner; q++) accumulate_phase_ms(p, p->tev[q]);
            }
            first = false;
            inner = p->h_lm->inner;
            if (p->h_lm->done) break;
        }
        p->sb.skip = nullptr;
        const LmDev& h = *p->h_lm;
        status = h.status;
        p->lambda = h.lambda; p->iterations = h.iterations; p->total_trials = h.total_trials;
        if (h.accepted) {
            std::swap(p->d_tv, p->d_tv_try); std::swap(p->d_sv, p->d_sv_try); p->parity ^= 1;
            p->err = h.err; p->have_lin = false;
        }
    } else {
        // ------------------------------------------------------------ sharded / speculative ladder (host verdicts)
        bool done = false;
        while (!done) {  // one round of the lambda ladder: G rungs (sequential algorithm: G == 1)
            double lam_try = p->lambda;
            for (int i = 0; i < p->spec_g && G > 1; i++) lam_try *= L.lambda_factor;
            b200_status st = push_lambda_try(p, lam_try); if (st != B200_OK) return st;
            st = launch_trial(p, p->time_phases ? p->tev[0] : nullptr, false);
            if (st != B200_OK) return st;
            if (G > 1) {
                // record of this rank: [scalars | static step]; gathered over ALL ranks of all groups
                double* rec = p->d_spec_rec + (size_t)(p->spec_g * T + p->rank) * p->spec_rec;
                CK(cudaMemcpyAsync(rec, p->sb.scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToDevice, p->stream));
                if (p->dm.ns) CK(cudaMemcpyAsync(rec + SC_COUNT + 1, p->sb.dstat, sizeof(double) * p->dm.ns, cudaMemcpyDeviceToDevice, p->stream));
                if (p->spec_ag(p->spec_ctx, p->d_spec_rec, sizeof(double) * p->spec_rec, p->stream)) { set_err("speculation allgather failed"); return B200_NCCL_ERROR; }
                CK(cudaMemcpyAsync(p->h_spec_rec, p->d_spec_rec, sizeof(double) * p->spec_rec * G * T, cudaMemcpyDeviceToHost, p->stream));
            }
            st = fetch_scalars(p); if (st != B200_OK) return st;
            p->host_syncs++;
            if (p->time_phases) {
                float ms;
                if (inner == 0) { cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]); p->phase_ms[0] += ms; }
                accumulate_phase_ms(p, p->tev[0]);
            }
            int winner = -1;
            double win_err = 0.0;
            for (int h = 0; h < G && !done; h++) {  // the rungs in the order the sequential loop meets them
                inner++;
                // SC_FAIL is the all-reduced count of ranks (of this group) whose factorisation met a non-positive pivot
                const double* sc = (G > 1) ? p->h_spec_rec + (size_t)h * T * p->spec_rec : p->h_scal;
                const bool solved = sc[SC_FAIL] == 0.0;
                const double newError = sc[SC_NEW_ERR];
                const TrialVerdict v = judge_trial(L, p->err, solved, sc[SC_LIN_ERR], sc[SC_MODEL_ERR], newError);
                status = solved ? B200_OK : B200_INDETERMINATE_SYSTEM;
                if (v.step_ok) {
                    winner = h; win_err = newError;
                    if (!L.gauss_newton) p->lambda = std::max(L.lambda_lower_bound, p->lambda / L.lambda_factor);
                    p->iterations++; p->total_trials++;
                    p->have_lin = false;
                    status = B200_OK;
                    done = true;
                } else if (!v.stop) {
                    p->lambda *= L.lambda_factor; p->total_trials++;
                    if (p->lambda >= L.lambda_upper_bound) { status = B200_LAMBDA_MAXED; done = true; }
                    else if (L.gauss_newton) { status = B200_INDETERMINATE_SYSTEM; done = true; }
                } else done = true;
            }
            if (winner >= 0) {
                if (G > 1) {
                    // every rank adopts the accepted group's step: gather the windows of all groups' steps, retract with the winner's
                    const size_t wdoubles = (size_t)p->Wk * p->dm.ntp;
                    CK(cudaMemcpyAsync(p->d_spec_delta + (size_t)(p->spec_g * T + p->rank) * wdoubles, p->d_delta + (size_t)p->k_lo * p->dm.ntp,
                                       sizeof(double) * wdoubles, cudaMemcpyDeviceToDevice, p->stream));
                    if (p->spec_ag(p->spec_ctx, p->d_spec_delta, sizeof(double) * wdoubles, p->stream)) { set_err("speculation allgather failed"); return B200_NCCL_ERROR; }
                    if (winner != p->spec_g)
                        launch_retract(p, p->d_spec_delta + (size_t)winner * T * wdoubles, p->d_spec_rec + (size_t)winner * T * p->spec_rec + SC_COUNT + 1);
                }
                std::swap(p->d_tv, p->d_tv_try); std::swap(p->d_sv, p->d_sv_try); p->parity ^= 1;
                p->err = win_err;
            }
        }
    }
    if (state) {
        state->error = p->err; state->lambda = p->lambda; state->iterations = p->iterations; state->inner_trials = inner;
        state->total_trials = p->total_trials; state->status = status;
    }
    return (b200_status)status;
}


// bioslam's fastOptimize loops (reference src/lowerBodyPoseEstimator.cpp:651-667 ; src/imuPoseEstimator.cpp:330-359)
b200_status b200_optimize(b200_problem* p, const b200_lm_params* lm, const b200_outer_params* op, double* hist, b200_optimize_result* res) {
    if (!p || !lm || !op) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    double e0 = 0;
    b200_status st = b200_lm_begin(p, lm, &e0);
    if (st != B200_OK) return st;
    CK(cudaEventRecord(p->ev[6], p->stream));
    double current = e0, previous = e0;
    int nh = 0, consec = 0, calls = 0;
    if (hist && nh < op->max_history) hist[nh++] = e0;
    b200_lm_state s{};
    if (op->single_imu_rule) {
        double absDec = 9.0e9, relDec = 9.0e9;
        while (relDec > op->rel_err_decrease_limit && absDec > op->abs_err_decrease_limit && current > op->abs_err_limit &&
               p->iterations < op->max_iterations) {
            const int before = p->iterations;
            st = b200_lm_iterate(p, &s); calls++;
            if (st == B200_CUDA_ERROR || st == B200_BAD_ARG || st == B200_NCCL_ERROR || st == B200_NOT_IMPLEMENTED) return st;
            previous = current; current = s.error;
            absDec = previous - current; relDec = absDec / previous;
            if (hist && nh < op->max_history) hist[nh++] = current;
            if (p->iterations == before) break;
        }
    } else {
        while (consec < op->converge_at_consec_small_iter && p->iterations < op->max_iterations) {
            const int before = p->iterations;
            previous = current;
            st = b200_lm_iterate(p, &s); calls++;
            if (st == B200_CUDA_ERROR || st == B200_BAD_ARG || st == B200_NCCL_ERROR || st == B200_NOT_IMPLEMENTED) return st;
            current = s.error;
            const double absDec = previous - current, relDec = absDec / previous;
            if (absDec < op->abs_err_decrease_limit || relDec < op->rel_err_decrease_limit || current < op->abs_err_limit) consec++;
            else consec = 0;
            if (hist && nh < op->max_history) hist[nh++] = current;
            if (p->iterations == before && st != B200_OK) break;
        }
    }
    CK(cudaEventRecord(p->ev[7], p->stream));
    CK(cudaEventSynchronize(p->ev[7]));
    float ms = 0;
    cudaEventElapsedTime(&ms, p->ev[6], p->ev[7]);
    if (res) {
        res->initial_error = e0; res->final_error = p->err; res->final_lambda = p->lambda; res->iterations = p->iterations;
        res->outer_calls = calls; res->total_trials = p->total_trials; res->n_history = nh; res->status = st; res->device_ms = ms;
    }
    return st;
}

b200_status b200_phase_ms(b200_problem* p, double out[12], int32_t reset) {
    if (!p) return B200_BAD_ARG;
    if (reset >= 0) p->time_phases = true;   // reset < 0: read without switching the per-phase events on
    if (out) for (int i = 0; i < 12; i++) out[i] = p->phase_ms[i];
    if (reset > 0) for (int i = 0; i < 12; i++) p->phase_ms[i] = 0;
    return B200_OK;
}

// runtime switches: "graph" (0/1: launch lambda trials as CUDA graphs), "spec_trials" (trials queued per host synchronisation),
// "phase_timing" (0/1: per-phase CUDA events; implies direct launches)
b200_status b200_set_option(b200_problem* p, const char* name, int32_t value) {
    if (!p || !name) return B200_BAD_ARG;
    const std::string n(name);
    if (n == "graph") p->use_graph = value != 0;
    else if (n == "spec_trials") p->spec_trials = std::max(1, std::min((int)value, 4));
    else if (n == "phase_timing") p->time_phases = value != 0;
    else { set_err("unknown option " + n); return B200_BAD_ARG; }
    return B200_OK;
}
int64_t b200_host_sync_count(const b200_problem* p) { return p ? p->host_syncs : 0; }

b200_status b200_timer_begin(b200_problem* p) {
    if (!p) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaEventRecord(p->ev[13], p->stream));
    return B200_OK;
}
b200_status b200_timer_end(b200_problem* p, double* ms) {
    if (!p || !ms) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    CK(cudaEventRecord(p->ev[14], p->stream));
    CK(cudaEventSynchronize(p->ev[14]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, p->ev[13], p->ev[14]));
    *ms = f;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ preintegration (A0)
b200_status b200_preintegrate(int32_t device, int32_t n_imu, int32_t n_samples, int32_t n_per, int32_t n_intervals, const double* acc,
                              const double* gyro, double dt, const double* bias_hat, const b200_imu_noise* noise, int32_t combined,
                              double* out) {
    if (!acc || !gyro || !bias_hat || !noise || !out || n_imu < 1 || n_per < 1 || n_intervals < 1 || !(dt > 0) ||
        (long long)n_intervals * n_per > n_samples) { set_err("bad preintegration arguments"); return B200_BAD_ARG; }
    CK(cudaSetDevice(device));
    const int rec = combined ? 190 : 115;
    const size_t ns3 = (size_t)n_imu * n_samples * 3;
    double *d_acc = nullptr, *d_gyro = nullptr, *d_bh = nullptr, *d_out = nullptr, *d_outT = nullptr;
    int* d_bad = nullptr;
    b200_status ret = B200_OK;
    auto cleanup = [&]() { cudaFree(d_acc); cudaFree(d_gyro); cudaFree(d_bh); cudaFree(d_out); cudaFree(d_outT); cudaFree(d_bad); };
#define CKP(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_err(std::string(#call) + ": " + cudaGetErrorString(e__)); cleanup(); return B200_CUDA_ERROR; } } while (0)
    CKP(dalloc(&d_acc, ns3)); CKP(dalloc(&d_gyro, ns3)); CKP(dalloc(&d_bh, (size_t)n_imu * 6));
    CKP(dalloc(&d_out, (size_t)n_imu * rec * n_intervals)); CKP(dalloc(&d_outT, (size_t)n_imu * rec * n_intervals)); CKP(dalloc(&d_bad, (size_t)1));
    CKP(cudaMemcpy(d_acc, acc, sizeof(double) * ns3, cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_gyro, gyro, sizeof(double) * ns3, cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(d_bh, bias_hat, sizeof(double) * n_imu * 6, cudaMemcpyHostToDevice));
    CKP(cudaMemset(d_bad, 0, sizeof(int)));
    B200_LAUNCH(preintegrate_kernel, dim3((n_intervals + 63) / 64, n_imu), dim3(64), 0, (cudaStream_t)0, n_imu, n_samples, n_per, n_intervals,
                d_acc, d_gyro, dt, d_bh, *noise, combined, d_out, (long long)n_intervals, d_bad);
    // d_out is [n_imu*rec][n_intervals] (SoA); ABI output is [n_imu][n_intervals][rec]
    for (int m = 0; m < n_imu; m++) {
        const long long n = (long long)n_intervals * rec;
        B200_LAUNCH(transpose_out_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)0, d_out + (size_t)m * rec * n_intervals,
                    n_intervals, rec, (long long)n_intervals, d_outT + (size_t)m * rec * n_intervals);
    }
    CKP(cudaDeviceSynchronize());
    CKP(cudaGetLastError());
    int bad = 0;
    CKP(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    CKP(cudaMemcpy(out, d_outT, sizeof(double) * n_imu * rec * n_intervals, cudaMemcpyDeviceToHost));
    cleanup();
    if (bad) { set_err("preintegrated covariance not positive definite"); ret = B200_INDETERMINATE_SYSTEM; }
    return ret;
}

b200_status b200_preintegrate_combined(int32_t device, int32_t n_imu, int32_t n_samples, int32_t n_per, int32_t n_intervals, const double* acc,
                                       const double* gyro, double dt, const double* bias_hat, const b200_imu_noise* noise, double* out) {
    return b200_preintegrate(device, n_imu, n_samples, n_per, n_intervals, acc, gyro, dt, bias_hat, noise, 1, out);
}

// ------------------------------------------------------------------------------------------ derived joint angles (N1)
b200_status b200_lower_body_joint_angles(b200_problem* p, const b200_joint_angle_desc* d, double* angles) {
    if (!p || !d || !angles) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    JointAngleArgs ja;
    for (int i = 0; i < 5; i++) {
        const int v = d->pose_var[i];
        if (v < 0 || v >= p->ntv || p->tvtype[v] != B200_VAR_POSE3) { set_err("joint angles: pose_var must name Pose3 time variables"); return B200_BAD_ARG; }
        ja.pose_off[i] = p->tv_valoff[v];
    }
    for (int i = 0; i < 14; i++) {
        const int v = d->static_var[i];
        if (v < 0 || v >= p->nsv || (p->svtype[v] != B200_VAR_VEC3 && p->svtype[v] != B200_VAR_UNIT3)) { set_err("joint angles: static_var must name Point3 / Unit3 static variables"); return B200_BAD_ARG; }
        ja.static_off[i] = p->sv_valoff[v];
    }
    const size_t n = (size_t)p->K * 12;
    if (n > p->stage_doubles) { set_err("internal: staging buffer too small"); return B200_NOT_IMPLEMENTED; }
    B200_LAUNCH(joint_angles_kernel, dim3((p->K + 127) / 128), dim3(128), 0, p->stream, ja, p->K, p->dm.Kld, p->d_tv, p->d_sv, p->d_stage);
    p->launches++;
    CK(cudaMemcpyAsync(angles, p->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ multi-rank plumbing
b200_status b200_comm_set(b200_problem* p, int32_t world_size, int32_t rank, b200_allgather_fn ag, b200_allreduce_fn ar, void* ctx) {
    if (!p || world_size < 1 || rank < 0 || rank >= world_size) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    p->world = world_size; p->rank = rank; p->ag_fn = ag; p->ar_fn = ar; p->comm_ctx = ctx;
    free_graphs(p);
    b200_status st = set_partition(p, 0, 0);
    // the speculation buffers are sized by the window of a rank: re-create them if the sharding changes afterwards
    if (st == B200_OK && p->spec_G > 1) st = b200_spec_set(p, p->spec_G, p->spec_g, p->spec_ag, p->spec_ctx);
    return st;
}

// lambda speculation: this process belongs to group `group` of `n_groups`; each group is one (possibly sharded) solver
// instance registered through b200_comm_set BEFORE this call; `allgather` spans all n_groups * world_size ranks, ordered
// group-major (global rank = group * world_size + rank).
b200_status b200_spec_set(b200_problem* p, int32_t n_groups, int32_t group, b200_allgather_fn allgather, void* ctx) {
    if (!p || n_groups < 1 || group < 0 || group >= n_groups || (n_groups > 1 && !allgather)) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    if (p->d_spec_delta) { cudaFree(p->d_spec_delta); p->d_spec_delta = nullptr; }
    if (p->d_spec_rec) { cudaFree(p->d_spec_rec); p->d_spec_rec = nullptr; }
    if (p->h_spec_rec) { cudaFreeHost(p->h_spec_rec); p->h_spec_rec = nullptr; }
    p->spec_G = n_groups; p->spec_g = group; p->spec_ag = n_groups > 1 ? allgather : nullptr; p->spec_ctx = ctx;
    if (n_groups == 1) return B200_OK;
    const size_t ranks = (size_t)n_groups * p->world;
    p->spec_rec = (SC_COUNT + 1 + p->dm.ns + 1) & ~1;
    CK(dalloc(&p->d_spec_delta, ranks * (size_t)p->Wk * p->dm.ntp));
    CK(dalloc(&p->d_spec_rec, ranks * p->spec_rec));
    CK(cudaMallocHost((void**)&p->h_spec_rec, ranks * p->spec_rec * sizeof(double)));
    return B200_OK;
}

// In-library NCCL (alternative to the b200_comm_set / b200_spec_set callbacks): rank 0 calls b200_nccl_unique_id and hands the
// 128 bytes to every rank by any transport; every rank then calls b200_nccl_init with the TOTAL rank count, its global rank and
// the number of lambda groups G (ranks are group-major: global rank = group * (n_ranks / G) + rank in group).
b200_status b200_nccl_unique_id(char out[128]) {
#ifndef B200_USE_EMU
    if (!out) return B200_BAD_ARG;
    NcclApi& N = nccl_api();
    if (!N.ok) { set_err("libnccl.so.2 not found"); return B200_NCCL_ERROR; }
    ncclUniqueId id;
    if (N.GetUniqueId(&id) != ncclSuccess) { set_err("ncclGetUniqueId failed"); return B200_NCCL_ERROR; }
    std::memcpy(out, id.internal, 128);
    return B200_OK;
#else
    (void)out; set_err("the emulator build has no NCCL"); return B200_NOT_IMPLEMENTED;
#endif
}
b200_status b200_nccl_init(b200_problem* p, int32_t n_ranks, int32_t global_rank, const char id_bytes[128], int32_t n_groups) {
#ifndef B200_USE_EMU
    if (!p || !id_bytes || n_ranks < 1 || global_rank < 0 || global_rank >= n_ranks || n_groups < 1 || n_ranks % n_groups) return B200_BAD_ARG;
    NcclApi& N = nccl_api();
    if (!N.ok) { set_err("libnccl.so.2 not found"); return B200_NCCL_ERROR; }
    CK(cudaSetDevice(p->device));
    ncclUniqueId id;
    std::memcpy(id.internal, id_bytes, 128);
    ncclResult_t r = N.CommInitRank(&p->nccl_all.comm, n_ranks, id, global_rank);
    if (r != ncclSuccess) { set_err(std::string("ncclCommInitRank: ") + (N.GetErrorString ? N.GetErrorString(r) : "error")); return B200_NCCL_ERROR; }
    p->nccl_all.nranks = n_ranks; p->nccl_all.rank = global_rank;
    const int T = n_ranks / n_groups, group = global_rank / T, rank = global_rank % T;
    if (n_groups > 1) {
        r = N.CommSplit(p->nccl_all.comm, group, rank, &p->nccl_grp.comm, nullptr);
        if (r != ncclSuccess) { set_err(std::string("ncclCommSplit: ") + (N.GetErrorString ? N.GetErrorString(r) : "error")); return B200_NCCL_ERROR; }
    } else p->nccl_grp.comm = p->nccl_all.comm;
    p->nccl_grp.nranks = T; p->nccl_grp.rank = rank;
    p->own_nccl = true;
    b200_status st = b200_comm_set(p, T, rank, nccl_allgather_cb, nccl_allreduce_cb, &p->nccl_grp);
    if (st != B200_OK) return st;
    return b200_spec_set(p, n_groups, group, n_groups > 1 ? nccl_allgather_cb : nullptr, &p->nccl_all);
#else
    (void)p; (void)n_ranks; (void)global_rank; (void)id_bytes; (void)n_groups; set_err("the emulator build has no NCCL"); return B200_NOT_IMPLEMENTED;
#endif
}
// device time spent in the collectives since the last reset (measured only while phase timing is on) and the number of
// collective groups issued
b200_status b200_comm_ms(b200_problem* p, double* ms, int64_t* calls, int32_t reset) {
    if (!p) return B200_BAD_ARG;
    if (ms) *ms = p->comm_ms;
    if (calls) *calls = p->comm_calls;
    if (reset) { p->comm_ms = 0; p->comm_calls = 0; }
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ parity / debug accessors
b200_status b200_debug_linearize(b200_problem* p) {
    if (!p) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    launch_linearize(p);
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    p->have_lin = true;
    return B200_OK;
}

b200_status b200_debug_get_gradient(b200_problem* p, double* g_time, double* g_static) {
    if (!p || !p->have_lin) return B200_BAD_ARG;
    const Dims& dm = p->dm;
    std::vector<double> A((size_t)p->K * dm.ns1 * dm.ntp), S((size_t)dm.ns1 * dm.ns1);
    CK(cudaMemcpy(A.data(), p->d_A, sizeof(double) * A.size(), cudaMemcpyDeviceToHost));
    for (int k = 0; k < p->K; k++)
        for (int i = 0; i < dm.ntp; i++)
            if (p->int2ext[i] >= 0) g_time[(size_t)k * dm.nt + p->int2ext[i]] = -A[((size_t)k * dm.ns1 + dm.ns) * dm.ntp + i];
    // static gradient: sum over keyframes of the rhs column of Sx_k
    std::vector<double> Sall((size_t)p->K * dm.ns1 * dm.ns1);
    CK(cudaMemcpy(Sall.data(), p->d_S, sizeof(double) * Sall.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < dm.ns; i++) {
        double acc = 0;
        for (int k = 0; k < p->K; k++) acc += Sall[(size_t)k * dm.ns1 * dm.ns1 + (size_t)i * dm.ns1 + dm.ns];
        g_static[i] = -acc;
    }
    return B200_OK;
}

b200_status b200_debug_get_hessian_blocks(b200_problem* p, int32_t k, double* D, double* O, double* A) {
    if (!p || !p->have_lin || k < 0 || k >= p->K) return B200_BAD_ARG;
    const Dims& dm = p->dm;
    const int nt = dm.nt, ntp = dm.ntp, nc = dm.nc, nl = dm.nl, ns = dm.ns, ns1 = dm.ns1;
    const int nco = dm.nco;
    std::vector<double> Db((size_t)ntp * ntp), Ob((size_t)nco * nco), Ab((size_t)ns1 * ntp);
    CK(cudaMemcpy(Db.data(), p->d_D + (size_t)k * ntp * ntp, sizeof(double) * Db.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Ob.data(), p->d_O + (size_t)k * nco * nco, sizeof(double) * Ob.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Ab.data(), p->d_A + (size_t)k * ns1 * ntp, sizeof(double) * Ab.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < nt * nt; i++) { D[i] = 0; O[i] = 0; }
    for (int i = 0; i < ntp; i++)
        for (int j = 0; j < ntp; j++) {
            const int ei = p->int2ext[i], ej = p->int2ext[j];
            if (ei < 0 || ej < 0) continue;
            D[(size_t)ei * nt + ej] = (i >= j) ? Db[(size_t)i * ntp + j] : Db[(size_t)j * ntp + i];   // only the lower triangle is assembled
            // O (k, k+1): rows k, cols k+1 ; device stores (k+1 chain) x (k chain)
            if (k + 1 < p->K && i >= nl && j >= nl) O[(size_t)ej * nt + ei] = Ob[(size_t)(i - nl) * nco + (j - nl)];
        }
    for (int i = 0; i < ntp; i++) {
        const int ei = p->int2ext[i];
        if (ei < 0) continue;
        for (int s = 0; s < ns; s++) A[(size_t)ei * ns + s] = Ab[(size_t)s * ntp + i];
    }
    return B200_OK;
}

b200_status b200_debug_get_static_hessian(b200_problem* p, double* S) {
    if (!p || !p->have_lin) return B200_BAD_ARG;
    const Dims& dm = p->dm;
    std::vector<double> Sall((size_t)p->K * dm.ns1 * dm.ns1);
    CK(cudaMemcpy(Sall.data(), p->d_S, sizeof(double) * Sall.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < dm.ns; i++)
        for (int j = 0; j < dm.ns; j++) {
            double acc = 0;
            for (int k = 0; k < p->K; k++) acc += Sall[(size_t)k * dm.ns1 * dm.ns1 + (size_t)i * dm.ns1 + j];
            S[(size_t)i * dm.ns + j] = acc;
        }
    return B200_OK;
}

b200_status b200_debug_solve(b200_problem* p, double lambda, double* delta_time, double* delta_static) {
    if (!p || !p->have_lin) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    const Dims& dm = p->dm;
    b200_status st = push_lambda_try(p, lambda);
    if (st != B200_OK) return st;
    st = launch_solve(p, nullptr);
    if (st != B200_OK) return st;
    // a non-positive pivot on ANY rank marks the solve indeterminate on every rank
    CK(cudaMemsetAsync(p->sb.scal, 0, sizeof(double) * SC_COUNT, p->stream));
    B200_LAUNCH(status_to_scal_kernel, dim3(1), dim3(32), 0, p->stream, p->sb.status, p->sb.scal);
    p->launches++;
    st = allreduce_scal(p);
    if (st != B200_OK) return st;
    st = fetch_scalars(p);
    if (st != B200_OK) return st;
    if (p->h_scal[SC_FAIL] != 0.0) return B200_INDETERMINATE_SYSTEM;
    std::vector<double> dl((size_t)p->K * dm.ntp);
    CK(cudaMemcpy(dl.data(), p->d_delta, sizeof(double) * dl.size(), cudaMemcpyDeviceToHost));
    for (int k = 0; k < p->K; k++)
        for (int i = 0; i < dm.ntp; i++)
            if (p->int2ext[i] >= 0) delta_time[(size_t)k * dm.nt + p->int2ext[i]] = dl[(size_t)k * dm.ntp + i];
    if (dm.ns) CK(cudaMemcpy(delta_static, p->sb.dstat, sizeof(double) * dm.ns, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// reduced blocks of keyframe k after the keyframe-local dofs have been eliminated at damping lambda (kernels_local.cuh), in the
// internal chain order: Dc[nc*nc] (symmetric, both triangles filled here), Ac[ns1*nc] (static rows + rhs row), Sc[ns1*ns1].
// dims[4] = {nl, nc, ns1, used (0: this problem has no pre-elimination)}; chain_ext[nc]: external tangent index of every chain dof.
b200_status b200_debug_local_elim(b200_problem* p, int32_t k, double lambda, int32_t dims[4], int32_t* chain_ext, double* Dc, double* Ac, double* Sc) {
    if (!p || !p->have_lin || k < 0 || k >= p->K || !dims) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    const Dims& dm = p->dm;
    dims[0] = dm.nl; dims[1] = dm.nc; dims[2] = dm.ns1; dims[3] = p->use_le ? 1 : 0;
    if (chain_ext) for (int i = 0; i < dm.nc; i++) chain_ext[i] = p->int2ext[dm.nl + i];
    if (!p->use_le || !Dc || !Ac || !Sc) return B200_OK;
    if (lambda >= 0.0) {   // (negative: read the blocks the last solve left behind)
        b200_status st = push_lambda_try(p, lambda);
        if (st != B200_OK) return st;
        CK(cudaMemsetAsync(p->sb.status, 0, sizeof(int), p->stream));
        B200_LAUNCH(local_elim_kernel, dim3(1), dim3(B200_LE_THREADS), p->smem_le, p->stream, p->le, (int)k, p->sb.lam, p->sb.status, (const int*)nullptr, (const int*)nullptr);
        p->launches++;
    }
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaGetLastError());
    const int nc = dm.nc, ncp = p->ds.ntp, ns1 = dm.ns1;
    std::vector<double> d((size_t)ncp * ncp), a((size_t)ns1 * ncp);
    CK(cudaMemcpy(d.data(), p->d_Dc + (size_t)k * ncp * ncp, sizeof(double) * d.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(a.data(), p->d_Ac + (size_t)k * ns1 * ncp, sizeof(double) * a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Sc, p->d_Sc + (size_t)k * ns1 * ns1, sizeof(double) * ns1 * ns1, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nc; i++) for (int j = 0; j < nc; j++) Dc[(size_t)i * nc + j] = i >= j ? d[(size_t)i * ncp + j] : d[(size_t)j * ncp + i];
    for (int r = 0; r < ns1; r++) for (int j = 0; j < nc; j++) Ac[(size_t)r * nc + j] = a[(size_t)r * ncp + j];
    return B200_OK;
}

// raw read of a solver buffer (development / tests): name in {"Wst", "Yst", "delta_c", "delta", "Dc", "Ac", "O", "F0"}
b200_status b200_debug_read(b200_problem* p, const char* name, int64_t offset, int64_t n, double* out) {
    if (!p || !name || !out) return B200_BAD_ARG;
    CK(cudaSetDevice(p->device));
    const std::string s(name);
    if (s == "config") {   // out[0..3] = {local pre-elimination on, pipelined level-0 sweep on, solver time-block size, shared bytes of the sweep}
        const double v[4] = {p->use_le ? 1.0 : 0.0, p->ds.pipe ? 1.0 : 0.0, (double)p->ds.ntp, (double)p->smem_sweep};
        for (int64_t i = 0; i < n && i < 4; i++) out[i] = v[i];
        return B200_OK;
    }
    const double* src = nullptr;
    if (s == "Wst") src = p->sb.Wst; else if (s == "Yst") src = p->sb.Yst; else if (s == "delta_c") src = p->sb.delta;
    else if (s == "delta") src = p->d_delta; else if (s == "Dc") src = p->d_Dc; else if (s == "Ac") src = p->d_Ac; else if (s == "O") src = p->d_O;
    else if (s == "F0") src = p->levels.empty() || !p->levels[0].F ? p->d_Ftop : p->levels[0].F;
    if (!src) return B200_BAD_ARG;
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaMemcpy(out, src + offset, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// one factor instance on the device from explicit inputs (tests over all factor types against tests/golden/factors.npz)
b200_status b200_debug_factor_eval(int32_t device, int32_t type, const double* params, const double* consts, const double* xcat,
                                   int32_t whiten, double* r, double* J) {
    FInfo fi;
    if (!finfo(type, fi) || !params || !xcat || !r || !J) { set_err("factor_eval: bad arguments"); return B200_BAD_ARG; }
    if (fi.nconst > 0 && !consts) { set_err("factor_eval: this factor type needs its constant record"); return B200_BAD_ARG; }
    CK(cudaSetDevice(device));
    int nx = 0, cols = 0;
    for (int v = 0; v < fi.nvars; v++) { nx += kValDim[fi.vt[v]]; cols += kTanDim[fi.vt[v]]; }
    const int rows = factor_rows(type), nc = std::max(fi.nconst, 1);
    double *d_prm = nullptr, *d_c = nullptr, *d_x = nullptr, *d_r = nullptr, *d_J = nullptr;
    int* d_vt = nullptr;
    auto cleanup = [&]() { cudaFree(d_prm); cudaFree(d_c); cudaFree(d_x); cudaFree(d_r); cudaFree(d_J); cudaFree(d_vt); };
#define CKF(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_err(std::string(#call) + ": " + cudaGetErrorString(e__)); cleanup(); return B200_CUDA_ERROR; } } while (0)
    CKF(dalloc(&d_prm, (size_t)B200_MAX_FACTOR_PARAMS)); CKF(dalloc(&d_c, (size_t)nc)); CKF(dalloc(&d_x, (size_t)nx));
    CKF(dalloc(&d_r, (size_t)rows)); CKF(dalloc(&d_J, (size_t)rows * cols)); CKF(dalloc(&d_vt, (size_t)fi.nvars));
    CKF(cudaMemcpy(d_prm, params, sizeof(double) * B200_MAX_FACTOR_PARAMS, cudaMemcpyHostToDevice));
    if (fi.nconst > 0) CKF(cudaMemcpy(d_c, consts, sizeof(double) * fi.nconst, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_x, xcat, sizeof(double) * nx, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_vt, fi.vt, sizeof(int) * fi.nvars, cudaMemcpyHostToDevice));
    B200_LAUNCH(factor_eval_kernel, dim3(1), dim3(32), 0, (cudaStream_t)0, type, fi.nvars, d_vt, cols, d_prm, d_c, d_x, whiten, d_r, d_J);
    CKF(cudaDeviceSynchronize());
    CKF(cudaGetLastError());
    CKF(cudaMemcpy(r, d_r, sizeof(double) * rows, cudaMemcpyDeviceToHost));
    CKF(cudaMemcpy(J, d_J, sizeof(double) * rows * cols, cudaMemcpyDeviceToHost));
    cleanup();
    return B200_OK;
}

// scalars of the last lambda trial: [linear.error(0), linear.error(delta), error(retracted values), -]
b200_status b200_debug_get_scalars(b200_problem* p, double out[4]) {
    if (!p || !out) return B200_BAD_ARG;
    for (int i = 0; i < 4; i++) out[i] = p->h_scal[i];
    return B200_OK;
}

#ifdef B200_PHASE_CLOCKS
b200_status b200_debug_phase_clocks(unsigned long long* out, int32_t reset) {
    CK(cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(unsigned long long) * 24));
    if (reset) { unsigned long long z[24] = {0}; CK(cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z))); }
    return B200_OK;
}
#endif

}  // extern "C"
