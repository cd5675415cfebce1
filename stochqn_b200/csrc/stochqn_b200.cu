// stochqn_b200.cu - the C ABI of include/stochqn.h on top of the sm_100a kernels.
//
// Host side = the reference's three re-entrant state machines (src/stochqn.c:978-1315:
// run_oLBFGS / run_SQN / run_adaQN), rewritten as pure control flow: every piece of vector
// arithmetic is enqueued as a kernel on the workspace's stream, and exactly one stream
// synchronisation per call brings back the handful of flag words the control flow needs
// (accept / reject of the direction, the two curvature dots).  Task codes, return values,
// counters and the quirks listed in SURVEY.md section 7.1 are reproduced; the arithmetic is
// not a translation (compact form instead of the two-loop, see kernels.cuh).
//
// Compiled twice: -DUSE_DOUBLE -> libstochqn_b200_f64.so, -DUSE_FLOAT -> libstochqn_b200_f32.so.
// There is no CPU fallback: without a CUDA device every entry point fails loudly.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "stochqn.h"
#include "stochqn_b200.h"
#include "kernels.cuh"
#include "internal_rs.h"

using namespace sqn;

extern std::atomic<unsigned long long> stochqn_b200_cb_launches;    // callbacks.cu

namespace {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "stochqn_b200: %s\n", g_err);
    return code;
}
#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) return fail(-2, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------
// communicator (NCCL, resolved at run time so that single-GPU use has no NCCL dependency)
// ------------------------------------------------------------------------------------------
struct Id128 { char b[128]; };          // same size / passing convention as ncclUniqueId
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;                                   // ncclGetUniqueId(ncclUniqueId*)
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;                // ncclCommInitRank(&comm, nranks, id, rank)
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return fail(-3, "cannot load NCCL (libnccl.so.2): %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank)) dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy)) dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce)) dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather)) dlsym(h, "ncclAllGather");
    g_nccl.ReduceScatter = (decltype(g_nccl.ReduceScatter)) dlsym(h, "ncclReduceScatter");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString)) dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.ReduceScatter)
        return fail(-3, "NCCL library lacks an expected symbol");
    g_nccl.handle = h;
    return 0;
}

struct Comm {
    void* nccl = nullptr;
    int rank = 0, world = 1;
    // peer-memory mailboxes for the fused small all-reduce (p2p.cuh); p2p == false -> library all-reduce
    bool p2p = false;
    unsigned char* box = nullptr;                 // this rank's mailbox (cudaMalloc, IPC-exported)
    unsigned char* peer[kMaxWorld] = {};          // every rank's mailbox as mapped here (peer[rank] == box)
    unsigned long long seq = 0;                   // exchanges issued so far (identical on every rank)
    cudaStream_t last_stream = 0;                 // stream the previous exchange was enqueued on
    cudaEvent_t order_ev = nullptr;               // chains exchanges issued on different streams (see next_exchange)
    int* error_flag = nullptr;                    // device word set by a timed-out stand-alone exchange
    // Buffers that every rank allocates and every rank maps (cudaIpc between processes, plain pointers inside an
    // in-process group): [2 parities][...], double-buffered by the parity of the call count
    struct PeerBuf {
        unsigned char* local = nullptr;
        unsigned char* peer[kMaxWorld] = {};      // peer[rank] == local
        long long blk = 0;                        // block length the buffers were sized for (0: not allocated)
        unsigned long long calls = 0;
    };
    PeerBuf rs;                                   // receive slots of the fused reduce-scatter (multinomial.cu): [2][world senders][blk]
    PeerBuf ag;                                   // gathered vector of the push all-gather: [2][world * blk]
    PeerBuf rp;                                   // send vectors of the pull reduce-scatter: [2][world * blk]
    double* barrier_scratch = nullptr;
    struct LocalGroup* group = nullptr;           // stochqn_b200_comm_init_inprocess: the ranks live in this process, no NCCL
};

// The ranks of an in-process communicator group (several "ranks" driven by one process on one device, each on its own
// stream: how the peer-memory collectives are exercised on a single-GPU box).  Peer buffers are shared by pointer.
struct LocalGroup {
    int world = 0, alive = 0;
    Comm* member[kMaxWorld] = {};
};

// PeerArgs of the NEXT exchange on this communicator (world = 0 when there is nothing to exchange or no p2p).
// The double-buffered mailboxes assume that the exchanges of one communicator execute in issue order on every rank
// (p2p.cuh).  Same stream: stream order gives that.  A caller that moves to another stream (the optimizer runs on the
// workspace stream, the stand-alone all-reduce / barrier / all-gather take any stream) is chained behind everything
// enqueued so far on the previous stream with an event, so exchange k+1 can never overtake exchange k.
PeerArgs next_exchange(Comm* cm, size_t count, cudaStream_t stream)
{
    PeerArgs pa;
    if (!cm || cm->world <= 1 || !cm->p2p || count > (size_t) kBoxCap) return pa;
    if (cm->seq > 0 && stream != cm->last_stream) {
        if (!cm->order_ev) cudaEventCreateWithFlags(&cm->order_ev, cudaEventDisableTiming);
        if (cm->order_ev && cudaEventRecord(cm->order_ev, cm->last_stream) == cudaSuccess) cudaStreamWaitEvent(stream, cm->order_ev, 0);
    }
    cm->last_stream = stream;
    pa.rank = cm->rank;
    pa.world = cm->world;
    pa.seq = ++cm->seq;
    for (int r = 0; r < cm->world; ++r) pa.box[r] = cm->peer[r];
    return pa;
}

// ------------------------------------------------------------------------------------------
// private per-workspace state
// ------------------------------------------------------------------------------------------
struct HostBlock {              // pinned, mapped: written by kernels, read by the host after the sync
    volatile int status;
    int pad;
    volatile double info[4];    // U bound, gamma, g'g
    volatile double pair[2];    // s'y, s's
    volatile double dir[2];     // sum d^2, number of non-finite entries
    volatile unsigned long long seq[3];   // written LAST by the kernel that fills status+info / pair / dir
    volatile double fval;       // objective value of the guided-mode driver (stochqn_b200_fit_batch)
};
enum { FLAG_STATUS = 0, FLAG_PAIR = 1, FLAG_DIR = 2 };

enum Kind { K_OLBFGS = 1, K_SQN = 2, K_ADAQN = 3 };

struct Ctx {
    Kind kind;
    int device = 0;
    cudaStream_t stream = 0;
    long long n = 0;
    size_t ld = 0;
    int msize = 0;
    int sm_count = 148;
    int max_grid = 0;
    // Gram state (fp64, physical-slot indexed)
    double *SY = nullptr, *YY = nullptr, *SS = nullptr;
    double *partials = nullptr, *sums = nullptr, *coef = nullptr;
    int* status_dev = nullptr;
    size_t rec_doubles = 0;             // widest partial record
    HostBlock* hb = nullptr;            // host view
    HostBlock* hb_dev = nullptr;        // device view of the same block
    int pending = -1;                   // slot of an accepted pair whose Gram column is not folded in yet
    int grad_writeback = -1;            // -1: automatic (device-pointer calls write the direction back, host-pointer calls do not)
    int trust_x_mirror = 1;             // host-pointer calls: x is uploaded once, then the device mirror is the truth
    bool host_call = false;             // the run_*() call in progress was given host pointers
    const void* x_host_last = nullptr;  // the caller's x array the mirror mirrors
    // host-pointer calls of oLBFGS: the caller's arrays cross PCIe in pieces on two copy streams while K1 / K3 / K4 work
    // on the pieces that have arrived (take_step_host / pair_host)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_out;
    cudaEvent_t ev_x = nullptr;
    long long chunk_elems = 0;          // piece length in elements (0: not set up yet)
    size_t partials_cap = 0;            // doubles allocated in `partials`
    // device-side request loop (kernels_loop.cuh, guided_impl.inc)
    LoopState* loop_dev = nullptr;      // the ring counters while a fit_batch / fit_batches call is in progress
    LoopState* loop_host = nullptr;     // pinned staging copy
    bool loop_active = false;           // the device record is the truth, the public struct is stale
    unsigned long long* loop_bar = nullptr;   // grid-barrier counter of the loop kernels
    double* rec2 = nullptr;             // 2-value partial records of the loop kernels
    long long loop_max_n = 0;           // largest n served by the loop kernels (0: never)
    double loop_steps = 0;              // steps taken by them
    // fused runs of mini-batches (kernels_fit.cuh)
    int fused_fit = 1;                  // 0: never (option FUSED_FIT)
    unsigned long long* fit_bar = nullptr;    // its grid-barrier counter
    unsigned long long* ada_bar = nullptr;    // grid-barrier words of kl_ada (its grid depends on n)
    bool ada_attr_set = false, mnfit_attr_set = false;
    // run_adaQN: the step of the pending call (take_step, Fisher ring write, niter) was already taken by a device-loop kernel;
    // only what follows it at a pair boundary (stochqn.c:1196-1239) is left to do.  Set by stochqn_b200_fit_batches.
    bool ada_step_done = false;
    int ada_step_changed_x = 0;
    void* fit_work = nullptr;           // the model's scratch buffer (partial column records)
    double fit_steps = 0;               // mini-batches served by it
    unsigned long long* fit_trace = nullptr;   // development aid (stochqn_b200_debug_fit_trace)
    Comm* comm = nullptr;
    long long n_global = 0;
    // host-pointer compatibility mode
    real_t *dx = nullptr, *dg = nullptr, *dhv = nullptr;     // device staging
    real_t *hreq = nullptr, *hreq_vec = nullptr;             // pinned host mirrors for *req / *req_vec
    bool x_mirror_valid = false;
    void* pub = nullptr;
    // optional per-kernel timing (STOCHQN_B200_OPT_PROFILE): CUDA events around K1 / K3 / K4 on the stream
    int profile = 0;
    // kernel classes: 0 = K1 / KA1 (dots), 1 = K3 / KA3 (combine + update), 2 = K4 (pair), 3 = KA2 (adaQN second dots)
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_armed[4] = {false, false, false, false};
    double prof_ms[4] = {0, 0, 0, 0};
    double prof_n[4] = {0, 0, 0, 0};
    double last_bound = 0;
    double exact_norm_steps = 0;        // steps that took the two-pass (exact ||d||) route
    unsigned long long seq_want[3] = {0, 0, 0};   // sequence numbers the host is waiting for (HostBlock::seq)
    int sync_return = 0;                // 1: drain the stream before every return, also for device-pointer calls
    // latency-bound sizes: the whole take_step in one cooperative launch (kernels_small.cuh)
    unsigned long long* bar = nullptr;  // grid-barrier counter (device), monotonically increasing
    unsigned long long bar_total = 0;   // its value once every launch issued so far has passed its barrier
    unsigned int* ticket = nullptr;     // "last block done" ticket of K4
    long long small_n = 0;              // use ks_step when n <= small_n (0: never)
    long long one_cta_n = kOneCtaN;     // ... as one 1024-thread CTA (ordinary launch) up to this n, cooperative grid above
    double small_steps = 0;             // steps that took the one-launch route
};

void prof_collect_one(Ctx* c, int k)      // waits for the closing event of kernel class k (long past when re-armed)
{
    if (!c->ev_armed[k]) return;
    float ms = 0;
    if (cudaEventSynchronize(c->ev[2 * k + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, c->ev[2 * k], c->ev[2 * k + 1]) == cudaSuccess) { c->prof_ms[k] += ms; c->prof_n[k] += 1; }
    c->ev_armed[k] = false;
}
void prof_begin(Ctx* c, int k) { if (c->profile) { prof_collect_one(c, k); cudaEventRecord(c->ev[2 * k], c->stream); } }
void prof_end(Ctx* c, int k) { if (c->profile) { cudaEventRecord(c->ev[2 * k + 1], c->stream); c->ev_armed[k] = true; } }
void prof_collect(Ctx* c)
{
    for (int k = 0; k < 4; ++k) prof_collect_one(c, k);
}

std::unordered_map<const void*, Ctx*> g_registry;
std::mutex g_reg_mu;

Ctx* find_ctx(const void* ws)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registry.find(ws);
    return it == g_registry.end() ? nullptr : it->second;
}

size_t padded_ld(long long n)
{
    const size_t q = 128 / sizeof(real_t);
    return ((size_t) n + q - 1) / q * q;
}

template <typename U>
cudaError_t dev_alloc_zero(U** p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**) p, count * sizeof(U));
    if (e != cudaSuccess) { *p = nullptr; return e; }
    return cudaMemset(*p, 0, count * sizeof(U));
}

int grid_for(const Ctx* c, long long work_items)
{
    long long g = (work_items + kThreads - 1) / kThreads;
    if (g < 1) g = 1;
    if (g > c->max_grid) g = c->max_grid;
    return (int) g;
}

bool is_device_ptr(const void* p)
{
    if (!p) return true;
    // the same few arrays come back call after call: remember the answer (device and host address ranges are
    // disjoint under unified addressing, so a remembered answer cannot go stale)
    thread_local const void* seen[4] = {nullptr, nullptr, nullptr, nullptr};
    thread_local bool kind[4] = {false, false, false, false};
    thread_local int next = 0;
    for (int i = 0; i < 4; ++i) if (seen[i] == p) return kind[i];
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    bool dev = false;
    if (e != cudaSuccess) cudaGetLastError();
    else dev = a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    seen[next] = p; kind[next] = dev; next = (next + 1) & 3;
    return dev;
}

constexpr int VECW = 16 / sizeof(real_t);
constexpr long long kSmallNDefault = 2048;      // measured on B200 (tools/probe_small.py): equal or up to 8 % faster below, slower above
bool aligned16(const void* p) { return (((uintptr_t) p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------
// kernel dispatch (rows per group x vector width)
// ------------------------------------------------------------------------------------------
#define COUNT_LAUNCH() g_launches.fetch_add(1, std::memory_order_relaxed)

int k1_grid(Ctx* c, const void* func, long long chunks, int threads = kThreads, int lanes = kLanes)
{
    // persistent-style launch: as many CTAs as are resident at once (occupancy x SMs), grid-stride inside
    static std::mutex mu;
    static std::unordered_map<const void*, int> occ;
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = occ.find(func);
        if (it == occ.end()) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
            occ[func] = per_sm;
        } else per_sm = it->second;
    }
    long long g = (chunks + lanes - 1) / lanes;
    const long long cap = (long long) c->sm_count * per_sm;
    if (g > cap) g = cap;
    if (g > c->max_grid) g = c->max_grid;
    if (g < 1) g = 1;
    return (int) g;
}

// A contiguous piece [off, off + len) of every n-vector, with the partial records it writes (host-pointer calls stream
// the caller's arrays in pieces so that the kernels overlap the PCIe copies; device-pointer calls use the whole range)
struct Range {
    long long off, len;
    double* partials;
};
Range whole(const Ctx* c) { return Range{0, c->n, c->partials}; }

template <int RPG, bool PENDING, int VEC>
int launch_k1_t(Ctx* c, const real_t* g, const real_t* S, const real_t* Y, int used, int j0,
                const real_t* sc, const real_t* yc, real_t* grad_prev, const Range& R)
{
    auto kern = k1_dots<real_t, RPG, PENDING, VEC>;
    const int grid = k1_grid(c, (const void*) kern, R.len / VEC);
    kern<<<grid, kThreads, 0, c->stream>>>(g + R.off, S + R.off, Y + R.off, c->ld, c->msize, used, j0, sc ? sc + R.off : nullptr,
                                          yc ? yc + R.off : nullptr, R.len, grad_prev ? grad_prev + R.off : nullptr, R.partials);
    COUNT_LAUNCH();
    return grid;
}

template <bool PENDING, int VEC>
int launch_k1_m(Ctx* c, int rpg, const real_t* g, const real_t* S, const real_t* Y, int used, int j0,
                const real_t* sc, const real_t* yc, real_t* gp, const Range& R)
{
    switch (rpg) {
        case 1:  return launch_k1_t<1, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        case 2:  return launch_k1_t<2, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        case 3:  return launch_k1_t<3, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        case 4:  return launch_k1_t<4, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        case 5:  return launch_k1_t<5, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        case 6:  return launch_k1_t<6, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        case 7:  return launch_k1_t<7, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
        default: return launch_k1_t<8, PENDING, VEC>(c, g, S, Y, used, j0, sc, yc, gp, R);
    }
}

// returns the number of CTAs whose partial records must be summed
int launch_k1(Ctx* c, const real_t* g, const real_t* S, const real_t* Y, int used, int pend, real_t* grad_prev, const Range& R)
{
    const bool vec = aligned16(g + R.off) && aligned16(grad_prev ? grad_prev + R.off : nullptr);
    const real_t* sc = pend >= 0 ? S + (size_t) pend * c->ld : nullptr;
    const real_t* yc = pend >= 0 ? Y + (size_t) pend * c->ld : nullptr;
    // RPG rows per group x 4 groups = 2*slots virtual rows -> up to 16 slots per launch; all launches of one
    // step use the same RPG (hence the same grid) so that they fill the same partial records
    const int per_launch = used > 16 ? 16 : used;
    int rpg = (2 * per_launch + kGroups - 1) / kGroups;
    if (rpg < 1) rpg = 1;
    int j0 = 0, grid = 1;
    do {
        if (pend >= 0) {
            grid = vec ? launch_k1_m<true, VECW>(c, rpg, g, S, Y, used, j0, sc, yc, grad_prev, R)
                       : launch_k1_m<true, 1>(c, rpg, g, S, Y, used, j0, sc, yc, grad_prev, R);
        } else {
            grid = vec ? launch_k1_m<false, VECW>(c, rpg, g, S, Y, used, j0, sc, yc, grad_prev, R)
                       : launch_k1_m<false, 1>(c, rpg, g, S, Y, used, j0, sc, yc, grad_prev, R);
        }
        j0 += 2 * rpg;
    } while (j0 < used);
    return grid;
}
int launch_k1(Ctx* c, const real_t* g, const real_t* S, const real_t* Y, int used, int pend, real_t* grad_prev)
{
    return launch_k1(c, g, S, Y, used, pend, grad_prev, whole(c));
}

template <int RPG, int MODE, int VEC>
int launch_k3_t(Ctx* c, const real_t* g, real_t* gout, real_t* S, const real_t* Y, int used, int new_slot,
                real_t* x, real_t* x_sum, real_t step, int force, const Range& R)
{
    auto kern = k3_combine<real_t, RPG, MODE, VEC>;
    const int grid = k1_grid(c, (const void*) kern, R.len / VEC, k3Threads, k3Lanes);
    kern<<<grid, k3Threads, 0, c->stream>>>(g + R.off, gout ? gout + R.off : nullptr, S + R.off, Y + R.off, S + R.off, c->ld, c->msize, used,
                                           new_slot, R.len, x + R.off, x_sum ? x_sum + R.off : nullptr, step,
                                           c->coef, c->status_dev, force, R.partials);
    COUNT_LAUNCH();
    return grid;
}

template <int MODE, int VEC>
int launch_k3_m(Ctx* c, int rpg, const real_t* g, real_t* gout, real_t* S, const real_t* Y, int used, int new_slot,
                real_t* x, real_t* x_sum, real_t step, int force, const Range& R)
{
#define K3_RPG(Q) case Q: return launch_k3_t<Q, MODE, VEC>(c, g, gout, S, Y, used, new_slot, x, x_sum, step, force, R)
    switch (rpg) {
        K3_RPG(1); K3_RPG(2); K3_RPG(3); K3_RPG(4); K3_RPG(5); K3_RPG(6); K3_RPG(7); K3_RPG(8);
        K3_RPG(10); K3_RPG(12); K3_RPG(16); K3_RPG(24);
        default: return launch_k3_t<32, MODE, VEC>(c, g, gout, S, Y, used, new_slot, x, x_sum, step, force, R);
    }
#undef K3_RPG
}


int rpg_for_k3(int used)        // rows per group of K3 = pairs in memory, rounded up to an instantiated bucket
{
    if (used < 1) return 1;
    if (used <= 8) return used;
    return used <= 10 ? 10 : used <= 12 ? 12 : used <= 16 ? 16 : used <= 24 ? 24 : 32;
}

int launch_k3(Ctx* c, int mode, const real_t* g, real_t* gout, real_t* S, const real_t* Y, int used, int new_slot,
              real_t* x, real_t* x_sum, real_t step, int force, const Range& R)
{
    const bool vec = aligned16(g + R.off) && aligned16(gout ? gout + R.off : nullptr) && aligned16(x + R.off) &&
                     aligned16(x_sum ? x_sum + R.off : nullptr);
    const int rpg = rpg_for_k3(used);
#define K3_CASE(M)                                                                                                  \
    return vec ? launch_k3_m<M, VECW>(c, rpg, g, gout, S, Y, used, new_slot, x, x_sum, step, force, R)             \
               : launch_k3_m<M, 1>(c, rpg, g, gout, S, Y, used, new_slot, x, x_sum, step, force, R)
    if (mode == MODE_OLBFGS) { K3_CASE(MODE_OLBFGS); }
    else if (mode == MODE_AVG) { K3_CASE(MODE_AVG); }
    else { K3_CASE(MODE_DIRONLY); }
#undef K3_CASE
}
int launch_k3(Ctx* c, int mode, const real_t* g, real_t* gout, real_t* S, const real_t* Y, int used, int new_slot,
              real_t* x, real_t* x_sum, real_t step, int force)
{
    return launch_k3(c, mode, g, gout, S, Y, used, new_slot, x, x_sum, step, force, whole(c));
}

void launch_k3_apply(Ctx* c, int mode, real_t* grad, real_t* S, int new_slot, real_t* x, real_t* x_sum, real_t step)
{
    const bool vec = aligned16(grad) && aligned16(x) && aligned16(x_sum);
    const int grid = grid_for(c, c->n / (vec ? VECW : 1));
    if (mode == MODE_OLBFGS) {
        if (vec) k3_apply<real_t, MODE_OLBFGS, VECW><<<grid, kThreads, 0, c->stream>>>(grad, S, c->ld, new_slot, c->n, x, x_sum, step);
        else     k3_apply<real_t, MODE_OLBFGS, 1><<<grid, kThreads, 0, c->stream>>>(grad, S, c->ld, new_slot, c->n, x, x_sum, step);
    } else {
        if (vec) k3_apply<real_t, MODE_AVG, VECW><<<grid, kThreads, 0, c->stream>>>(grad, S, c->ld, new_slot, c->n, x, x_sum, step);
        else     k3_apply<real_t, MODE_AVG, 1><<<grid, kThreads, 0, c->stream>>>(grad, S, c->ld, new_slot, c->n, x, x_sum, step);
    }
    COUNT_LAUNCH();
}

// `publish`: the last CTA to finish sums the 2-value records and publishes them to the host pair block itself
// (no k_finalize launch); only when the optimizer is not sharded - the exchange between ranks lives in k_finalize.
template <int KIND>
int launch_k4_k(Ctx* c, const real_t* a, const real_t* b, const real_t* s, real_t* y, real_t y_reg, bool publish, const Range& R,
                const PeerArgs& pa = PeerArgs())
{
    const bool vec = aligned16(a + R.off) && aligned16(b + R.off);
    const int grid = grid_for(c, R.len / (vec ? VECW : 1));
    unsigned int* ticket = publish ? c->ticket : nullptr;
    const unsigned long long seq = publish ? ++c->seq_want[FLAG_PAIR] : 0;
    if (vec) k4_pair<real_t, KIND, VECW><<<grid, kThreads, 0, c->stream>>>(a + R.off, b + R.off, s + R.off, y + R.off, y_reg, R.len, R.partials, ticket,
                                                                          c->sums, c->hb_dev->pair, &c->hb_dev->seq[FLAG_PAIR], seq, pa);
    else     k4_pair<real_t, KIND, 1><<<grid, kThreads, 0, c->stream>>>(a + R.off, b + R.off, s + R.off, y + R.off, y_reg, R.len, R.partials, ticket,
                                                                       c->sums, c->hb_dev->pair, &c->hb_dev->seq[FLAG_PAIR], seq, pa);
    COUNT_LAUNCH();
    return grid;
}
template <int KIND>
int launch_k4_k(Ctx* c, const real_t* a, const real_t* b, const real_t* s, real_t* y, real_t y_reg, bool publish = false,
                const PeerArgs& pa = PeerArgs())
{
    return launch_k4_k<KIND>(c, a, b, s, y, y_reg, publish, whole(c), pa);
}

template <int OP>
void launch_avg(Ctx* c, real_t* x_sum, real_t* other, real_t* s_slot, real_t inv)
{
    const bool vec = aligned16(other);
    const int grid = grid_for(c, c->n / (vec ? VECW : 1));
    if (vec) k_avg<real_t, OP, VECW><<<grid, kThreads, 0, c->stream>>>(x_sum, other, s_slot, inv, c->n);
    else     k_avg<real_t, OP, 1><<<grid, kThreads, 0, c->stream>>>(x_sum, other, s_slot, inv, c->n);
    COUNT_LAUNCH();
}

// ---- all-reduce of the small sum record (multi-GPU) ----------------------------------------
int allreduce_sums(Ctx* c, double* buf, size_t count)
{
    if (!c->comm || c->comm->world <= 1) return 0;
    if (!c->comm->nccl) return fail(-5, "an in-process communicator has no library all-reduce (%zu values do not fit a mailbox)", count);
    int r = g_nccl.AllReduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm->nccl, c->stream);
    if (r != 0) return fail(-4, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return 0;
}

double step_limit(const Ctx* c) { return 1e3 * (double) (c->n_global > 0 ? c->n_global : c->n); }

// Reduce the K1 partials, sum them over the ranks, solve.  ONE launch: on one GPU, and when sharded with
// peer-memory mailboxes (the exchange happens inside the kernel, p2p.cuh); three steps with a library all-reduce.
int launch_solve(Ctx* c, bool ada, int nblocks, int used, int oldest, int pend, int check_nan, double h0)
{
    SolveArgs A;
    A.msize = c->msize; A.used = used; A.oldest = oldest; A.pend = pend;
    A.check_nan = check_nan; A.h0 = h0; A.limit = step_limit(c);
    A.seq = ++c->seq_want[FLAG_STATUS];
    auto go = [&](const SolveArgs& a, const PeerArgs& pa) {
        k2_solve<<<1, kThreads, 0, c->stream>>>(a, pa, c->partials, c->sums, c->SY, c->YY, c->SS, c->coef,
                                                c->status_dev, &c->hb_dev->status, c->hb_dev->info, &c->hb_dev->seq[FLAG_STATUS]);
        COUNT_LAUNCH();
    };
    (void) ada;
    const size_t P = (size_t) (4 * c->msize + 2);
    const bool sharded = c->comm && c->comm->world > 1;
    PeerArgs pa = sharded ? next_exchange(c->comm, P, c->stream) : PeerArgs();
    if (sharded && pa.world == 0) {
        A.nblocks = nblocks; A.do_solve = 0; go(A, pa);
        if (int r = allreduce_sums(c, c->sums, P)) return r;
        A.nblocks = 0; A.do_solve = 1; go(A, pa);
    } else {
        A.nblocks = nblocks; A.do_solve = 1; go(A, pa);
    }
    return 0;
}

// Sum `count`-wide partial records, over the ranks too, and publish them into mapped host memory (flag `which`).
int launch_finalize(Ctx* c, int nblocks, int count, volatile double* host_dst, int which)
{
    const unsigned long long seq = host_dst ? ++c->seq_want[which] : 0;
    volatile unsigned long long* seq_dst = host_dst ? &c->hb_dev->seq[which] : nullptr;
    const bool sharded = c->comm && c->comm->world > 1;
    PeerArgs pa = sharded ? next_exchange(c->comm, (size_t) count, c->stream) : PeerArgs();
    if (sharded && pa.world == 0) {
        k_finalize<<<1, kThreads, 0, c->stream>>>(c->partials, nblocks, count, c->sums, pa, nullptr, nullptr, 0);
        COUNT_LAUNCH();
        if (int r = allreduce_sums(c, c->sums, (size_t) count)) return r;
        if (host_dst) { k_publish<<<1, 32, 0, c->stream>>>(c->sums, count, host_dst, seq_dst, seq); COUNT_LAUNCH(); }
    } else if (!sharded && !host_dst && count > 2 * kWarps) {
        k_finalize_wide<<<(unsigned) ((count + kWarps - 1) / kWarps), kThreads, 0, c->stream>>>(c->partials, nblocks, count, c->sums);
        COUNT_LAUNCH();
    } else {
        k_finalize<<<1, kThreads, 0, c->stream>>>(c->partials, nblocks, count, c->sums, pa, host_dst, seq_dst, seq);
        COUNT_LAUNCH();
    }
    return 0;
}
int launch_pair_finalize(Ctx* c, int nblocks, volatile double* host_dst)
{
    return launch_finalize(c, nblocks, 2, host_dst, host_dst == c->hb_dev->pair ? FLAG_PAIR : FLAG_DIR);
}

// K4 + publication of s'y, s's to the host pair block in ONE launch: the CTA that finishes last sums the records and,
// when the optimizer is sharded over peer memory, exchanges the two values with the other ranks itself.  Only the
// library-all-reduce fallback still needs k_finalize between K4 and the publication.
template <int KIND>
int launch_pair(Ctx* c, const real_t* a, const real_t* b, const real_t* s, real_t* y, real_t y_reg, bool profile = false)
{
    const bool sharded = c->comm && c->comm->world > 1;
    PeerArgs pa = sharded ? next_exchange(c->comm, 2, c->stream) : PeerArgs();
    const bool fused = !sharded || pa.world > 1;
    if (profile) prof_begin(c, 2);
    const int nb = launch_k4_k<KIND>(c, a, b, s, y, y_reg, fused, pa);
    if (profile) prof_end(c, 2);
    if (!fused) return launch_pair_finalize(c, nb, c->hb_dev->pair);
    return 0;
}

// take_step of oLBFGS / SQN in one cooperative launch (kernels_small.cuh); same outputs as K1 -> K2 -> K3
template <int MODE>
int launch_small_step(Ctx* c, const real_t* g, real_t* gout, real_t* S, const real_t* Y, int used, int oldest, int pend,
                      int new_slot, real_t* x, real_t* x_sum, real_t* grad_prev, real_t step, int check_nan, double h0)
{
    SmallArgs K;
    K.msize = c->msize; K.used = used; K.pend = pend; K.new_slot = new_slot; K.n = c->n; K.ld = c->ld;
    // n <= kOneCtaN: ONE CTA of 1024 threads, ordinary launch, no grid barrier; above: cooperative grid of 256-thread CTAs
    const bool one_cta = c->n <= c->one_cta_n;
    long long grid = one_cta ? 1 : (c->n + kThreads - 1) / kThreads;
    if (grid > c->sm_count) grid = c->sm_count;
    if (grid < 1) grid = 1;
    if (grid > 1) c->bar_total += (unsigned long long) grid;
    K.bar_target = c->bar_total;
    SolveArgs A;
    A.msize = c->msize; A.used = used; A.oldest = oldest; A.pend = pend; A.nblocks = (int) grid; A.do_solve = 1;
    A.check_nan = check_nan; A.h0 = h0; A.limit = step_limit(c);
    A.seq = ++c->seq_want[FLAG_STATUS];
    double *partials = c->partials, *SY = c->SY, *YY = c->YY, *SS = c->SS, *coef = c->coef;
    int* status_dev = c->status_dev;
    volatile int* status_host = &c->hb_dev->status;
    volatile double* info_host = c->hb_dev->info;
    volatile unsigned long long* seq_host = &c->hb_dev->seq[FLAG_STATUS];
    unsigned long long* bar = c->bar;
    void* args[] = {&K, &A, &g, &gout, &S, &Y, &x, &x_sum, &grad_prev, &step, &partials, &SY, &YY, &SS, &coef,
                    &status_dev, &status_host, &info_host, &seq_host, &bar};
    if (grid > 1) CUDA_TRY(cudaLaunchCooperativeKernel((const void*) ks_step<real_t, MODE>, dim3((unsigned) grid), dim3(kThreads), args, 0, c->stream));
    else CUDA_TRY(cudaLaunchKernel((const void*) ks_step<real_t, MODE>, dim3(1), dim3(kOneCtaThreads), args, 0, c->stream));
    COUNT_LAUNCH();
    c->small_steps += 1;
    return 0;
}

int sync_stream(Ctx* c)
{
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(-2, "device work failed: %s", cudaGetErrorString(e));
    return 0;
}

// Wait until the kernel that fills flag block `which` has published it.  The host spins on the mapped sequence
// word instead of draining the stream: the answer is there the moment the small kernel retires, while the
// streaming kernel queued behind it (K3) is still running - the call returns and the caller queues its next
// kernel without a bubble on the GPU.
int wait_flag(Ctx* c, int which)
{
    const unsigned long long want = c->seq_want[which];
    volatile unsigned long long* p = &c->hb->seq[which];
    // the payload words (status, info, pair, dir) are read after this returns: the load that observes the sequence number
    // is an acquire, so that a weakly ordered host CPU cannot hoist those reads above it
    auto seen = [&]() { return __atomic_load_n(const_cast<unsigned long long*>(p), __ATOMIC_ACQUIRE) == want; };
    for (unsigned long spins = 1;; ++spins) {
        if (seen()) return 0;
        if ((spins & 0xfffu) == 0) {                     // every 4096 polls: has the stream died or drained?
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e == cudaSuccess) {
                if (seen()) return 0;
                return fail(-2, "device work finished without publishing its result (flag %d)", which);
            }
            if (e != cudaErrorNotReady) return fail(-2, "device work failed: %s", cudaGetErrorString(e));
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
}

// End of a call: host-pointer callers (and sync_return) get the reference's contract - everything final on
// return; device-pointer callers get stream-ordered semantics (results are final in the workspace stream's order).
int finish_call(Ctx* c, bool host_mode)
{
    if (host_mode || c->sync_return) return sync_stream(c);
    return 0;
}

// ------------------------------------------------------------------------------------------
// Ctx construction / destruction
// ------------------------------------------------------------------------------------------
void free_ctx(Ctx* c)
{
    if (!c) return;
    cudaFree(c->SY); cudaFree(c->YY); cudaFree(c->SS);
    cudaFree(c->partials); cudaFree(c->sums); cudaFree(c->coef); cudaFree(c->status_dev);
    cudaFree(c->bar); cudaFree(c->ticket);
    if (c->hb) cudaFreeHost((void*) c->hb);
    cudaFree(c->dx); cudaFree(c->dg); cudaFree(c->dhv);
    if (c->hreq) cudaFreeHost(c->hreq);
    if (c->hreq_vec) cudaFreeHost(c->hreq_vec);
    for (int k = 0; k < 8; ++k) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_out) cudaEventDestroy(e);
    if (c->ev_x) cudaEventDestroy(c->ev_x);
    cudaFree(c->loop_dev); cudaFree(c->loop_bar); cudaFree(c->rec2); cudaFree(c->fit_bar); cudaFree(c->ada_bar);
    if (c->loop_host) cudaFreeHost(c->loop_host);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    delete c;
}

Ctx* make_ctx(Kind kind, long long n, int msize, int fisher_size)
{
    Ctx* c = new Ctx();
    c->kind = kind;
    c->n = n;
    c->n_global = n;
    c->ld = padded_ld(n);
    c->msize = msize;
    bool ok = cudaGetDevice(&c->device) == cudaSuccess;
    cudaDeviceProp prop;
    ok = ok && cudaGetDeviceProperties(&prop, c->device) == cudaSuccess;
    if (ok) {
        c->sm_count = prop.multiProcessorCount;
        c->max_grid = c->sm_count * 8;       // upper bound on resident CTAs of 256 threads (2048 threads / SM)
    }
    const size_t m = (size_t) msize;
    size_t rec = 4 * m + 2;
    if (kind == K_ADAQN) {
        size_t r2 = (size_t) ka1_record((int) m) + (size_t) ka2_record((int) m);     // both sum records are live at once
        if (r2 > rec) rec = r2;
        if ((size_t) fisher_size > rec) rec = (size_t) fisher_size;
    }
    c->rec_doubles = rec;
    ok = ok && dev_alloc_zero(&c->SY, m * m) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->YY, m * m) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->SS, m) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->partials, (size_t) c->max_grid * rec) == cudaSuccess;
    c->partials_cap = (size_t) c->max_grid * rec;
    ok = ok && dev_alloc_zero(&c->sums, rec) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->coef, 2 * m + 4) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->status_dev, 1) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->bar, 1) == cudaSuccess;
    ok = ok && dev_alloc_zero(&c->ticket, 1) == cudaSuccess;
    if (ok) {
        // one-launch step for latency-bound sizes (measured crossover on B200: tools/probe_small.py); needs cooperative
        // launch; STOCHQN_B200_SMALL_N overrides the threshold (0 disables)
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        const char* e = getenv("STOCHQN_B200_SMALL_N");
        c->small_n = e ? atoll(e) : kSmallNDefault;
        const char* e1 = getenv("STOCHQN_B200_ONE_CTA_N");
        if (e1) c->one_cta_n = atoll(e1);
        if (!coop && c->small_n > c->one_cta_n) c->small_n = c->one_cta_n;       // the multi-CTA form needs a cooperative launch
        const char* e2 = getenv("STOCHQN_B200_LOOP_MAX_N");
        c->loop_max_n = e2 ? atoll(e2) : (kind == K_ADAQN ? (1ll << 19) : (1ll << 16));     // adaQN: its one-launch step (kl_ada) pays up to L2-sized states
        if (!coop && c->loop_max_n > kOneCtaN) c->loop_max_n = kOneCtaN;
        const char* e3 = getenv("STOCHQN_B200_FUSED_FIT");
        c->fused_fit = (coop && !(e3 && atoi(e3) == 0)) ? 1 : 0;
    }
    ok = ok && cudaHostAlloc((void**) &c->hb, sizeof(HostBlock), cudaHostAllocMapped) == cudaSuccess;
    if (ok) {
        memset((void*) c->hb, 0, sizeof(HostBlock));
        ok = cudaHostGetDevicePointer((void**) &c->hb_dev, (void*) c->hb, 0) == cudaSuccess;
    }
    if (!ok) {
        cudaError_t e = cudaGetLastError();
        fail(-2, "cannot set up the device workspace (%s) - a CUDA device is required, there is no CPU path",
             cudaGetErrorString(e));
        free_ctx(c);
        return nullptr;
    }
    return c;
}

void register_ctx(void* pub, Ctx* c)
{
    c->pub = pub;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_registry[pub] = c;
}

Ctx* unregister_ctx(const void* pub)
{
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registry.find(pub);
    if (it == g_registry.end()) return nullptr;
    Ctx* c = it->second;
    g_registry.erase(it);
    return c;
}

// ------------------------------------------------------------------------------------------
// host-pointer compatibility: stage caller arrays through device mirrors
// ------------------------------------------------------------------------------------------
struct Staged {
    real_t* dev = nullptr;     // pointer the kernels use
    real_t* host = nullptr;    // caller's host pointer when staged, else NULL
};

int stage_in(Ctx* c, real_t* p, real_t** mirror, Staged* out, bool upload)
{
    out->dev = p;
    out->host = nullptr;
    if (!p || is_device_ptr(p)) return 0;
    if (!*mirror) CUDA_TRY(cudaMalloc((void**) mirror, (size_t) c->n * sizeof(real_t)));
    if (upload) CUDA_TRY(cudaMemcpyAsync(*mirror, p, (size_t) c->n * sizeof(real_t), cudaMemcpyHostToDevice, c->stream));
    out->dev = *mirror;
    out->host = p;
    return 0;
}

int stage_out(Ctx* c, const Staged& s)
{
    if (!s.host) return 0;
    CUDA_TRY(cudaMemcpyAsync(s.host, s.dev, (size_t) c->n * sizeof(real_t), cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

// `*req` for a workspace-owned device buffer: device callers get the device pointer, host
// callers a pinned host mirror filled here.
int publish_req(Ctx* c, bool host_mode, real_t* dev_buf, real_t** mirror, real_t** out)
{
    if (!host_mode) { *out = dev_buf; return 0; }
    if (!*mirror) CUDA_TRY(cudaHostAlloc((void**) mirror, (size_t) c->n * sizeof(real_t), cudaHostAllocDefault));
    CUDA_TRY(cudaMemcpyAsync(*mirror, dev_buf, (size_t) c->n * sizeof(real_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = *mirror;
    return 0;
}

int enter(Ctx* c)
{
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device) CUDA_TRY(cudaSetDevice(c->device));
    return 0;
}

// ------------------------------------------------------------------------------------------
// shared pieces of the three state machines
// ------------------------------------------------------------------------------------------
void flush_bfgs(bfgs_mem* m, Ctx* c)           // stochqn.c:554-558
{
    m->mem_used = 0;
    m->mem_st_ix = 0;
    c->pending = -1;
}

inline int oldest_slot(const bfgs_mem* m)      // stochqn.c:820 (quirk Q8)
{
    return (m->mem_st_ix == m->mem_used) ? 0 : (int) m->mem_st_ix;
}

// does `grad` receive the search direction?  (the reference documents the array only as "modified in place")
inline bool writeback(const Ctx* c) { return c->grad_writeback < 0 ? !c->host_call : c->grad_writeback != 0; }
inline bool mirror_is_current(const Ctx* c, const void* x) { return c->trust_x_mirror && c->x_mirror_valid && c->x_host_last == x; }

// take_step (stochqn.c:802-840) for oLBFGS / SQN.  mode = MODE_OLBFGS or MODE_AVG.
// Returns <0 on a CUDA failure, else 0 and *info is set to search_direction_was_nan on rejection.
int take_step_qn(Ctx* c, bfgs_mem* m, int mode, real_t step, real_t* x, real_t* g, real_t* grad_prev,
                 real_t* x_sum, double h0, int check_nan, info_enum* info)
{
    const int used = (int) m->mem_used;
    const int st = (int) m->mem_st_ix;
    real_t* gout = writeback(c) ? g : nullptr;
    const bool sharded = c->comm && c->comm->world > 1;
    if (!sharded && c->small_n > 0 && c->n <= c->small_n) {
        // latency-bound size: dots, solve and update in ONE cooperative launch
        int r = mode == MODE_OLBFGS
            ? launch_small_step<MODE_OLBFGS>(c, g, gout, m->s_mem, m->y_mem, used, oldest_slot(m), c->pending, st, x, x_sum, grad_prev, step, check_nan, h0)
            : launch_small_step<MODE_AVG>(c, g, gout, m->s_mem, m->y_mem, used, oldest_slot(m), c->pending, st, x, x_sum, grad_prev, step, check_nan, h0);
        if (r) return r;
        c->pending = -1;
    } else {
        prof_begin(c, 0);
        int nb = launch_k1(c, g, m->s_mem, m->y_mem, used, c->pending, grad_prev);
        prof_end(c, 0);
        if (int r = launch_solve(c, false, nb, used, oldest_slot(m), c->pending, check_nan, h0)) return r;
        c->pending = -1;                    // the Gram column is folded in whatever happens next
        prof_begin(c, 1);
        launch_k3(c, mode, g, gout, m->s_mem, m->y_mem, used, st, x, x_sum, step, 0);
        prof_end(c, 1);
    }
    if (int r = wait_flag(c, FLAG_STATUS)) return r;          // K2 has retired; K3 may still be running
    int status = c->hb->status;
    c->last_bound = c->hb->info[0];
    if (status == ST_COMM_TIMEOUT) return fail(-4, "a peer rank did not join the all-reduce of this step");
    if (status == ST_NEED_EXACT_NORM) {
        c->exact_norm_steps += 1;
        // rare: the cheap bound could not certify ||d|| <= 1e3*n.  Materialise d in `grad`, measure it
        // exactly, and only then touch x - the reference's order (stochqn.c:825-838).
        int nb2 = launch_k3(c, MODE_DIRONLY, g, g, m->s_mem, m->y_mem, used, st, x, x_sum, step, 1);
        if (int r = launch_pair_finalize(c, nb2, c->hb_dev->dir)) return r;
        if (int r = wait_flag(c, FLAG_DIR)) return r;
        const double dd = c->hb->dir[0], bad = c->hb->dir[1];
        if (bad > 0 || !(sqrt(dd) <= step_limit(c))) status = ST_REJECT_NONFINITE;
        else {
            launch_k3_apply(c, mode, g, m->s_mem, st, x, x_sum, step);
            status = ST_ACCEPT;
        }
    }
    if (status != ST_ACCEPT) {
        flush_bfgs(m, c);
        *info = search_direction_was_nan;
    }
    return 0;
}

// check_min_curvature (stochqn.c:883-900) given the two dots; handles quirk Q1 on rejection.
int curvature_decision(Ctx* c, bfgs_mem* m, double sy, double ss, info_enum* info)
{
    const int slot = (int) m->mem_st_ix;
    if (m->min_curvature > 0) {
        const real_t curv = (real_t) sy / (real_t) ss;            // the reference divides in real_t (stochqn.c:892); the sums are fp64
        if (curv <= m->min_curvature) {
            // rollback_corr_pair copies the never-written (zero) backup over the slot (stochqn.c:597-604)
            CUDA_TRY(cudaMemsetAsync(m->s_mem + (size_t) slot * c->ld, 0, (size_t) c->n * sizeof(real_t), c->stream));
            CUDA_TRY(cudaMemsetAsync(m->y_mem + (size_t) slot * c->ld, 0, (size_t) c->n * sizeof(real_t), c->stream));
            k_gram_zero_slot<<<1, 64, 0, c->stream>>>(c->SY, c->YY, c->SS, c->msize, slot);
            COUNT_LAUNCH();
            *info = curvature_too_small;
            return 0;
        }
    }
    m->mem_st_ix = (m->mem_st_ix + 1) % m->mem_size;                                   // stochqn.c:569-573
    m->mem_used = (m->mem_used + 1 >= m->mem_size) ? m->mem_size : m->mem_used + 1;
    c->pending = slot;
    return 0;
}

// update_y_grad_diff (stochqn.c:915-926): y = grad - grad_prev (+ y_reg*s), curvature test.
int update_y_grad_diff_dev(Ctx* c, bfgs_mem* m, const real_t* grad, const real_t* grad_prev, info_enum* info)
{
    const size_t slot = m->mem_st_ix;
    real_t* s = m->s_mem + slot * c->ld;
    real_t* y = m->y_mem + slot * c->ld;
    if (int r = launch_pair<PAIR_GRAD_DIFF>(c, grad, grad_prev, s, y, m->y_reg, true)) return r;
    if (int r = wait_flag(c, FLAG_PAIR)) return r;
    return curvature_decision(c, m, c->hb->pair[0], c->hb->pair[1], info);
}

// ------------------------------------------------------------------------------------------
// oLBFGS with HOST pointers (the drop-in compatibility mode), large n: the call is PCIe-bound, so the caller's arrays
// cross the bus in pieces on two copy streams and the kernels work on the pieces that have arrived:
//   step call:  H2D grad piece i  ->  K1 on piece i (own partial records)   | all pieces |  K2
//               K3 on piece i     ->  D2H x piece i (and grad when write-back is asked for)
//   pair call:  H2D grad piece i  ->  K4 on piece i                          | all pieces |  finalize + publish
// x itself is uploaded only when the device mirror is not known to be current (first call, another array, option 0).
// The partial records of the pieces are summed by the same fixed-order reduction, so results do not depend on timing.
// ------------------------------------------------------------------------------------------
constexpr long long kPipelineMinBytes = 32ll << 20;      // below this a call is latency-bound: one copy, one launch

bool use_pipeline(const Ctx* c)
{
    const bool sharded_nccl = c->comm && c->comm->world > 1 && !c->comm->p2p;
    return (long long) c->n * (long long) sizeof(real_t) >= kPipelineMinBytes && !sharded_nccl;
}

int pipeline_setup(Ctx* c, int* nchunks_out)
{
    if (!c->chunk_elems) {
        long long mb = 64;
        if (const char* e = getenv("STOCHQN_B200_STAGE_CHUNK_MB")) { long long v = atoll(e); if (v >= 1) mb = v; }
        long long elems = (mb << 20) / (long long) sizeof(real_t);
        while ((c->n + elems - 1) / elems > 64) elems *= 2;            // at most 64 pieces
        c->chunk_elems = elems;
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
    }
    const int nchunks = (int) ((c->n + c->chunk_elems - 1) / c->chunk_elems);
    while ((int) c->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_in.push_back(a);
        c->ev_out.push_back(b);
    }
    const size_t need = (size_t) nchunks * (size_t) c->max_grid * c->rec_doubles;
    if (c->partials_cap < need) {                // every piece writes its own block of partial records
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->partials);
        c->partials = nullptr;
        c->partials_cap = 0;
        CUDA_TRY(dev_alloc_zero(&c->partials, need));
        c->partials_cap = need;
    }
    if (!c->dx) CUDA_TRY(cudaMalloc((void**) &c->dx, (size_t) c->n * sizeof(real_t)));
    if (!c->dg) CUDA_TRY(cudaMalloc((void**) &c->dg, (size_t) c->n * sizeof(real_t)));
    *nchunks_out = nchunks;
    return 0;
}

// step call of run_oLBFGS with host pointers (stochqn.c:992-1021): returns <0 on failure, else 0 with *info set
int take_step_host(Ctx* c, workspace_oLBFGS* ws, real_t step, real_t* xh, real_t* gh, info_enum* info)
{
    bfgs_mem* m = ws->bfgs_memory;
    int nchunks = 0;
    if (int r = pipeline_setup(c, &nchunks)) return r;
    const int used = (int) m->mem_used, st = (int) m->mem_st_ix;
    const size_t P = (size_t) (4 * c->msize + 2);
    const bool up_x = !mirror_is_current(c, xh);
    const bool wb = writeback(c);
    long long recs = 0;
    prof_begin(c, 0);
    for (int i = 0; i < nchunks; ++i) {
        const long long off = (long long) i * c->chunk_elems;
        const long long len = (c->n - off < c->chunk_elems) ? c->n - off : c->chunk_elems;
        CUDA_TRY(cudaMemcpyAsync(c->dg + off, gh + off, (size_t) len * sizeof(real_t), cudaMemcpyHostToDevice, c->copy_in));
        CUDA_TRY(cudaEventRecord(c->ev_in[i], c->copy_in));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[i], 0));
        Range R{off, len, c->partials + (size_t) recs * P};
        recs += launch_k1(c, c->dg, m->s_mem, m->y_mem, used, c->pending, ws->grad_prev, R);
    }
    prof_end(c, 0);
    if (up_x) {
        CUDA_TRY(cudaMemcpyAsync(c->dx, xh, (size_t) c->n * sizeof(real_t), cudaMemcpyHostToDevice, c->copy_in));
        CUDA_TRY(cudaEventRecord(c->ev_x, c->copy_in));
    }
    if (int r = launch_solve(c, false, (int) recs, used, oldest_slot(m), c->pending, ws->check_nan, (double) ws->hess_init)) return r;
    c->pending = -1;
    if (up_x) CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_x, 0));
    // K3 reads the accept flag on the device: the pieces are queued at once, a rejected direction makes them no-ops
    prof_begin(c, 1);
    for (int i = 0; i < nchunks; ++i) {
        const long long off = (long long) i * c->chunk_elems;
        const long long len = (c->n - off < c->chunk_elems) ? c->n - off : c->chunk_elems;
        Range R{off, len, c->partials};
        launch_k3(c, MODE_OLBFGS, c->dg, wb ? c->dg : nullptr, m->s_mem, m->y_mem, used, st, c->dx, nullptr, step, 0, R);
        CUDA_TRY(cudaEventRecord(c->ev_out[i], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_out, c->ev_out[i], 0));
        CUDA_TRY(cudaMemcpyAsync(xh + off, c->dx + off, (size_t) len * sizeof(real_t), cudaMemcpyDeviceToHost, c->copy_out));
        if (wb) CUDA_TRY(cudaMemcpyAsync(gh + off, c->dg + off, (size_t) len * sizeof(real_t), cudaMemcpyDeviceToHost, c->copy_out));
    }
    prof_end(c, 1);
    if (int r = wait_flag(c, FLAG_STATUS)) return r;
    int status = c->hb->status;
    c->last_bound = c->hb->info[0];
    if (status == ST_COMM_TIMEOUT) return fail(-4, "a peer rank did not join the all-reduce of this step");
    if (status == ST_NEED_EXACT_NORM) {
        // rare: measure ||d|| exactly on the whole (now resident) vectors before touching x, as take_step_qn does
        c->exact_norm_steps += 1;
        CUDA_TRY(cudaStreamSynchronize(c->copy_out));          // the no-op pieces above copied the unchanged x: let them finish
        int nb2 = launch_k3(c, MODE_DIRONLY, c->dg, c->dg, m->s_mem, m->y_mem, used, st, c->dx, nullptr, step, 1);
        if (int r = launch_pair_finalize(c, nb2, c->hb_dev->dir)) return r;
        if (int r = wait_flag(c, FLAG_DIR)) return r;
        const double dd = c->hb->dir[0], bad = c->hb->dir[1];
        if (bad > 0 || !(sqrt(dd) <= step_limit(c))) status = ST_REJECT_NONFINITE;
        else {
            launch_k3_apply(c, MODE_OLBFGS, c->dg, m->s_mem, st, c->dx, nullptr, step);
            CUDA_TRY(cudaMemcpyAsync(xh, c->dx, (size_t) c->n * sizeof(real_t), cudaMemcpyDeviceToHost, c->stream));
            if (wb) CUDA_TRY(cudaMemcpyAsync(gh, c->dg, (size_t) c->n * sizeof(real_t), cudaMemcpyDeviceToHost, c->stream));
            status = ST_ACCEPT;
        }
    }
    if (status != ST_ACCEPT) {
        flush_bfgs(m, c);
        *info = search_direction_was_nan;
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->copy_out));
    c->x_mirror_valid = true;               // x on the host and its mirror agree again (updated or untouched on both sides)
    c->x_host_last = xh;
    return 0;
}

// pair call of run_oLBFGS with host pointers (stochqn.c:1024-1031)
int pair_host(Ctx* c, workspace_oLBFGS* ws, real_t* gh, info_enum* info)
{
    bfgs_mem* m = ws->bfgs_memory;
    int nchunks = 0;
    if (int r = pipeline_setup(c, &nchunks)) return r;
    const size_t slot = m->mem_st_ix;
    real_t* s = m->s_mem + slot * c->ld;
    real_t* y = m->y_mem + slot * c->ld;
    long long recs = 0;
    prof_begin(c, 2);
    for (int i = 0; i < nchunks; ++i) {
        const long long off = (long long) i * c->chunk_elems;
        const long long len = (c->n - off < c->chunk_elems) ? c->n - off : c->chunk_elems;
        CUDA_TRY(cudaMemcpyAsync(c->dg + off, gh + off, (size_t) len * sizeof(real_t), cudaMemcpyHostToDevice, c->copy_in));
        CUDA_TRY(cudaEventRecord(c->ev_in[i], c->copy_in));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[i], 0));
        Range R{off, len, c->partials + (size_t) recs * 2};
        recs += launch_k4_k<PAIR_GRAD_DIFF>(c, c->dg, ws->grad_prev, s, y, m->y_reg, false, R);
    }
    prof_end(c, 2);
    if (int r = launch_pair_finalize(c, (int) recs, c->hb_dev->pair)) return r;
    if (int r = wait_flag(c, FLAG_PAIR)) return r;
    if (int r = curvature_decision(c, m, c->hb->pair[0], c->hb->pair[1], info)) return r;
    return sync_stream(c);
}

int copy_vec_dev(Ctx* c, real_t* dst, const real_t* src)
{
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t) c->n * sizeof(real_t), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int invalid_ws(const char* who, task_enum* task)
{
    *task = invalid_input;
    fprintf(stderr, "%s got an invalid workspace as input.\n", who);
    return -1000;
}

// ---- bfgs_mem / fisher_mem allocation on the device -----------------------------------------
bfgs_mem* alloc_bfgs(size_t mem_size, long long n, real_t min_curvature, real_t y_reg, size_t upd_freq)
{
    bfgs_mem* out = (bfgs_mem*) calloc(1, sizeof(bfgs_mem));
    if (!out) return nullptr;
    const size_t ld = padded_ld(n);
    bool ok = dev_alloc_zero(&out->s_mem, mem_size * ld) == cudaSuccess;
    ok = ok && dev_alloc_zero(&out->y_mem, mem_size * ld) == cudaSuccess;
    // buffer_rho / buffer_alpha (two-loop scratch) and s_bak / y_bak (never-written backups, quirk Q1) have
    // no role in the compact form; tiny placeholders keep the pointers non-NULL for code that checks them
    ok = ok && dev_alloc_zero(&out->buffer_rho, mem_size) == cudaSuccess;
    ok = ok && dev_alloc_zero(&out->buffer_alpha, mem_size) == cudaSuccess;
    out->s_bak = nullptr;
    out->y_bak = nullptr;
    out->mem_size = mem_size;
    out->mem_used = 0;
    out->mem_st_ix = 0;
    out->upd_freq = upd_freq;
    out->y_reg = y_reg;
    out->min_curvature = min_curvature;
    if (!ok) {
        cudaGetLastError();
        fprintf(stderr, "Error: Could not allocate memory for BFGS storage.\n");
        cudaFree(out->s_mem); cudaFree(out->y_mem); cudaFree(out->buffer_rho); cudaFree(out->buffer_alpha);
        free(out);
        return nullptr;
    }
    return out;
}

void free_bfgs(bfgs_mem* m)
{
    if (!m) return;
    cudaFree(m->s_mem); cudaFree(m->y_mem); cudaFree(m->buffer_rho); cudaFree(m->buffer_alpha);
    free(m);
}

fisher_mem* alloc_fisher(size_t mem_size, long long n)
{
    fisher_mem* out = (fisher_mem*) calloc(1, sizeof(fisher_mem));
    if (!out) return nullptr;
    bool ok = dev_alloc_zero(&out->F, mem_size * padded_ld(n)) == cudaSuccess;
    ok = ok && dev_alloc_zero(&out->buffer_y, mem_size) == cudaSuccess;
    out->mem_size = mem_size;
    if (!ok) {
        cudaGetLastError();
        fprintf(stderr, "Error: Could not allocate memory for Fisher storage.\n");
        cudaFree(out->F); cudaFree(out->buffer_y);
        free(out);
        return nullptr;
    }
    return out;
}

void free_fisher(fisher_mem* f)
{
    if (!f) return;
    cudaFree(f->F); cudaFree(f->buffer_y);
    free(f);
}

bool bad_sizes(const char* who, int n, size_t mem_size)
{
    if (n <= 0 || mem_size == 0 || mem_size > (size_t) kMaxMem) {
        fail(-1, "%s: n must be positive and 1 <= mem_size <= %d (got n=%d, mem_size=%zu)", who, kMaxMem, n, mem_size);
        return true;
    }
    return false;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

bfgs_mem* initialize_bfgs_mem(const size_t mem_size, const int n, const real_t min_curvature, const real_t y_reg,
                              const size_t upd_freq)
{
    return alloc_bfgs(mem_size, n, min_curvature, y_reg, upd_freq);
}
void dealloc_bfgs_mem(bfgs_mem* bfgs_memory) { free_bfgs(bfgs_memory); }
fisher_mem* initialize_fisher_mem(const size_t mem_size, const int n) { return alloc_fisher(mem_size, n); }
void dealloc_fisher_mem(fisher_mem* fisher_memory) { free_fisher(fisher_memory); }

// ---- oLBFGS ---------------------------------------------------------------------------------
workspace_oLBFGS* initialize_oLBFGS(const int n, const size_t mem_size, const real_t hess_init, const real_t y_reg,
                                    const real_t min_curvature, const int check_nan, const int nthreads)
{
    if (bad_sizes("initialize_oLBFGS", n, mem_size)) return nullptr;
    workspace_oLBFGS* out = (workspace_oLBFGS*) calloc(1, sizeof(workspace_oLBFGS));
    Ctx* c = make_ctx(K_OLBFGS, n, (int) mem_size, 0);
    if (out) out->bfgs_memory = c ? alloc_bfgs(mem_size, n, min_curvature, y_reg, 1) : nullptr;
    bool ok = out && c && out->bfgs_memory && dev_alloc_zero(&out->grad_prev, (size_t) n) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        fprintf(stderr, "Error: Could not allocate memory for oLBFGS.\n");
        if (out) { free_bfgs(out->bfgs_memory); cudaFree(out->grad_prev); free(out); }
        free_ctx(c);
        return nullptr;
    }
    out->hess_init = hess_init;
    out->niter = 0;
    out->section = 0;
    out->check_nan = check_nan;
    out->nthreads = nthreads;
    out->n = n;
    register_ctx(out, c);
    return out;
}

void dealloc_oLBFGS(workspace_oLBFGS* ws)
{
    if (!ws) return;
    Ctx* c = unregister_ctx(ws);
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    free_bfgs(ws->bfgs_memory);
    cudaFree(ws->grad_prev);
    free_ctx(c);
    free(ws);
}

int run_oLBFGS(real_t step_size, real_t x[], real_t grad[], real_t** req, task_enum* task, workspace_oLBFGS* ws,
               info_enum* iter_info)
{
    *iter_info = no_problems_encountered;
    Ctx* c = ws ? find_ctx(ws) : nullptr;
    if (!c || c->kind != K_OLBFGS) return invalid_ws("oLBFGS", task);
    if (enter(c)) return invalid_ws("oLBFGS", task);
    bfgs_mem* m = ws->bfgs_memory;
    const bool host_mode = (x && !is_device_ptr(x)) || (grad && !is_device_ptr(grad));
    c->host_call = host_mode;

    if (ws->section == 0) {                                     // stochqn.c:983-989
        *task = calc_grad;
        *req = x;
        ws->section = 1;
        return 0;
    }

    const bool piped = x && grad && !is_device_ptr(x) && !is_device_ptr(grad) && use_pipeline(c);
    if (ws->section == 1 && piped) {                            // the same section, PCIe copies overlapped with the kernels
        if (take_step_host(c, ws, step_size, x, grad, iter_info) < 0) return invalid_ws("oLBFGS", task);
        ws->niter++;
        *req = x;
        if (*iter_info == no_problems_encountered) { *task = calc_grad_same_batch; ws->section = 2; return 1; }
        *task = calc_grad;
        return 0;
    }
    if (ws->section == 2 && piped) {
        if (pair_host(c, ws, grad, iter_info) < 0) return invalid_ws("oLBFGS", task);
        *task = calc_grad;
        *req = x;
        ws->section = 1;
        return 0;
    }

    if (ws->section == 1) {                                     // stochqn.c:992-1021
        Staged sx, sg;
        if (stage_in(c, x, &c->dx, &sx, !mirror_is_current(c, x))) return invalid_ws("oLBFGS", task);
        if (stage_in(c, grad, &c->dg, &sg, true)) return invalid_ws("oLBFGS", task);
        // grad_prev <- grad rides inside K1; the step writes s = -step*d into the next slot (1006-1007)
        if (take_step_qn(c, m, MODE_OLBFGS, step_size, sx.dev, sg.dev, ws->grad_prev, nullptr,
                         (double) ws->hess_init, ws->check_nan, iter_info) < 0)
            return invalid_ws("oLBFGS", task);
        ws->niter++;                                            // quirk Q7: also when the step was rejected
        *task = (*iter_info == no_problems_encountered) ? calc_grad_same_batch : calc_grad;
        *req = x;
        if (*iter_info == no_problems_encountered) {
            if (sx.host) { if (stage_out(c, sx)) return invalid_ws("oLBFGS", task); c->x_mirror_valid = true; c->x_host_last = x; }
            if (sg.host && writeback(c)) { if (stage_out(c, sg)) return invalid_ws("oLBFGS", task); }
            if (finish_call(c, host_mode)) return invalid_ws("oLBFGS", task);
            ws->section = 2;
            return 1;
        }
        if (finish_call(c, host_mode)) return invalid_ws("oLBFGS", task);
        ws->section = 1;
        return 0;
    }

    if (ws->section == 2) {                                     // stochqn.c:1024-1031
        Staged sg;
        if (stage_in(c, grad, &c->dg, &sg, true)) return invalid_ws("oLBFGS", task);
        if (update_y_grad_diff_dev(c, m, sg.dev, ws->grad_prev, iter_info) < 0) return invalid_ws("oLBFGS", task);
        if (finish_call(c, host_mode)) return invalid_ws("oLBFGS", task);
        *task = calc_grad;
        *req = x;
        ws->section = 1;
        return 0;
    }
    return invalid_ws("oLBFGS", task);                          // stochqn.c:1033-1035
}

// ---- SQN --------------------------------------------------------------------------------------
workspace_SQN* initialize_SQN(const int n, const size_t mem_size, const size_t bfgs_upd_freq, const real_t min_curvature,
                              const int use_grad_diff, const real_t y_reg, const int check_nan, const int nthreads)
{
    if (bad_sizes("initialize_SQN", n, mem_size)) return nullptr;
    workspace_SQN* out = (workspace_SQN*) calloc(1, sizeof(workspace_SQN));
    Ctx* c = make_ctx(K_SQN, n, (int) mem_size, 0);
    if (out) out->bfgs_memory = c ? alloc_bfgs(mem_size, n, min_curvature, y_reg, bfgs_upd_freq) : nullptr;
    bool ok = out && c && out->bfgs_memory;
    if (ok && use_grad_diff) ok = dev_alloc_zero(&out->grad_prev, (size_t) n) == cudaSuccess;
    ok = ok && dev_alloc_zero(&out->x_sum, (size_t) n) == cudaSuccess;
    ok = ok && dev_alloc_zero(&out->x_avg_prev, (size_t) n) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        fprintf(stderr, "Error: Could not allocate memory for SQN.\n");
        if (out) { free_bfgs(out->bfgs_memory); cudaFree(out->grad_prev); cudaFree(out->x_sum); cudaFree(out->x_avg_prev); free(out); }
        free_ctx(c);
        return nullptr;
    }
    out->use_grad_diff = use_grad_diff;
    out->niter = 0;
    out->section = 0;
    out->check_nan = check_nan;
    out->nthreads = nthreads;
    out->n = n;
    register_ctx(out, c);
    return out;
}

void dealloc_SQN(workspace_SQN* ws)
{
    if (!ws) return;
    Ctx* c = unregister_ctx(ws);
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    free_bfgs(ws->bfgs_memory);
    cudaFree(ws->grad_prev); cudaFree(ws->x_sum); cudaFree(ws->x_avg_prev);
    free_ctx(c);
    free(ws);
}

int run_SQN(real_t step_size, real_t x[], real_t grad[], real_t hess_vec[], real_t** req, real_t** req_vec,
            task_enum* task, workspace_SQN* ws, info_enum* iter_info)
{
    *iter_info = no_problems_encountered;
    int return_value = 0;
    Ctx* c = ws ? find_ctx(ws) : nullptr;
    if (!c || c->kind != K_SQN) return invalid_ws("SQN", task);
    if (enter(c)) return invalid_ws("SQN", task);
    bfgs_mem* m = ws->bfgs_memory;
    const bool host_mode = x && !is_device_ptr(x);
    c->host_call = host_mode;
#define SQN_FAIL() return invalid_ws("SQN", task)
#define SQN_RESUME() do { ws->section = 1; *task = calc_grad; *req = x; return return_value; } while (0)

    if (ws->section == 0) SQN_RESUME();                         // stochqn.c:1044-1048

    if (ws->section == 1) {                                     // stochqn.c:1051-1115
        Staged sx, sg;
        if (stage_in(c, x, &c->dx, &sx, !mirror_is_current(c, x))) SQN_FAIL();
        if (stage_in(c, grad, &c->dg, &sg, true)) SQN_FAIL();
        if (take_step_qn(c, m, MODE_AVG, step_size, sx.dev, sg.dev, nullptr, ws->x_sum, 0.0, ws->check_nan, iter_info) < 0)
            SQN_FAIL();
        ws->niter++;
        return_value = (*iter_info == search_direction_was_nan) ? 0 : 1;
        if (return_value == 0) launch_avg<AVG_ADD>(c, ws->x_sum, sx.dev, nullptr, (real_t) 0);   // 1067, quirk Q7
        if (sx.host && return_value) { if (stage_out(c, sx)) SQN_FAIL(); c->x_mirror_valid = true; c->x_host_last = x; }
        if (sg.host && return_value && writeback(c)) { if (stage_out(c, sg)) SQN_FAIL(); }

        const size_t L = m->upd_freq;
        if ((ws->niter % L) != 0) { if (finish_call(c, host_mode)) SQN_FAIL(); SQN_RESUME(); }
        const real_t inv = (real_t) 1 / (real_t) L;             // average_from_sum, stochqn.c:286-291
        if (ws->niter == L) {                                   // 1078-1094
            launch_avg<AVG_ARCHIVE>(c, ws->x_sum, ws->x_avg_prev, nullptr, L > 1 ? inv : (real_t) 1);
            if (ws->use_grad_diff) {
                *task = calc_grad_big_batch;
                if (publish_req(c, host_mode, ws->x_avg_prev, &c->hreq, req)) SQN_FAIL();
                if (finish_call(c, host_mode)) SQN_FAIL();
                ws->section = 2;
                return return_value;
            }
            if (finish_call(c, host_mode)) SQN_FAIL();
            SQN_RESUME();
        }
        // update_s_vector (861-870): x_sum becomes the average, s = x_avg - x_avg_prev into the next slot
        real_t* s_slot = m->s_mem + m->mem_st_ix * c->ld;
        launch_avg<AVG_S_VECTOR>(c, ws->x_sum, ws->x_avg_prev, s_slot, L > 1 ? inv : (real_t) 1);
        if (publish_req(c, host_mode, ws->x_sum, &c->hreq, req)) SQN_FAIL();
        if (ws->use_grad_diff) {                                // 1100-1105
            *task = calc_grad_big_batch;
            ws->section = 3;
        } else {                                                // 1107-1113
            *task = calc_hess_vec;
            ws->section = 4;
            if (publish_req(c, host_mode, s_slot, &c->hreq_vec, req_vec)) SQN_FAIL();
        }
        if (finish_call(c, host_mode)) SQN_FAIL();
        return return_value;
    }

    if (ws->section == 2) {                                     // 1118-1122
        Staged sg;
        if (stage_in(c, grad, &c->dg, &sg, true)) SQN_FAIL();
        if (copy_vec_dev(c, ws->grad_prev, sg.dev)) SQN_FAIL();
        if (finish_call(c, host_mode)) SQN_FAIL();
        SQN_RESUME();
    }

    if (ws->section == 3) {                                     // 1125-1134
        Staged sg;
        if (stage_in(c, grad, &c->dg, &sg, true)) SQN_FAIL();
        if (update_y_grad_diff_dev(c, m, sg.dev, ws->grad_prev, iter_info) < 0) SQN_FAIL();
        if (*iter_info == no_problems_encountered) {
            if (copy_vec_dev(c, ws->grad_prev, sg.dev)) SQN_FAIL();
            if (copy_vec_dev(c, ws->x_avg_prev, ws->x_sum)) SQN_FAIL();
        }
        if (cudaMemsetAsync(ws->x_sum, 0, (size_t) c->n * sizeof(real_t), c->stream) != cudaSuccess) SQN_FAIL();
        if (finish_call(c, host_mode)) SQN_FAIL();
        SQN_RESUME();
    }

    if (ws->section == 4) {                                     // 1137-1142 (quirk Q6: archive before the test)
        Staged shv;
        if (stage_in(c, hess_vec, &c->dhv, &shv, true)) SQN_FAIL();
        launch_avg<AVG_ARCHIVE>(c, ws->x_sum, ws->x_avg_prev, nullptr, (real_t) 1);
        const size_t slot = m->mem_st_ix;
        if (launch_pair<PAIR_COPY>(c, shv.dev, shv.dev, m->s_mem + slot * c->ld, m->y_mem + slot * c->ld, (real_t) 0)) SQN_FAIL();
        if (wait_flag(c, FLAG_PAIR)) SQN_FAIL();
        if (curvature_decision(c, m, c->hb->pair[0], c->hb->pair[1], iter_info)) SQN_FAIL();
        if (finish_call(c, host_mode)) SQN_FAIL();
        SQN_RESUME();
    }
    return invalid_ws("SQN", task);                             // 1144-1146
#undef SQN_FAIL
#undef SQN_RESUME
}

}  // extern "C"

#include "adaqn_impl.inc"
#include "ext_impl.inc"
#include "guided_impl.inc"
