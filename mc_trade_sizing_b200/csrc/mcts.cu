// mcts.cu -- host side of libmcts.so: the C ABI of include/mcts.h over the kernels of kernels.cuh.
// One handle = one device, one stream; every call is chunked over nodes of NumPy's pairwise tree
// over the path index so that per-chunk partial sums combine to np.mean's exact result.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mcts.h"
#include "aux_kernels.cuh"
#include "kernels.cuh"
#include "planner.hpp"
#include "stager.hpp"

using namespace mcts;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char buf__[512];                                                                             \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return fail(e__ == cudaErrorMemoryAllocation ? MCTS_ENOMEM : MCTS_ECUDA, buf__);            \
        }                                                                                                \
    } while (0)

struct EventPair {
    cudaEvent_t a, b;
};

struct mcts_handle {
    int device = 0;
    int sm_count = 0;
    int block_threads = 0;  // experiment override (mcts_options.block_threads)
    int int_fast_path = 0;  // mcts_options.int_fast_path
    // Two lanes (stream + scratch buffers): consecutive chunks alternate between them so that the
    // memory-bound reduction and the wave tail of one chunk overlap the issue-bound path kernel of the
    // next.  lane[0].stream is "the" stream of the handle (mcts_get_stream).
    struct Lane {
        cudaStream_t stream = nullptr;
        uint32_t *d_counts = nullptr;
        int64_t cap_sims = 0;
        double *d_out = nullptr;
        int64_t cap_paths = 0;
        double *d_part = nullptr;
        int64_t cap_part = 0;  // 160 * pstride
        uint32_t *d_scratch = nullptr;  // global code store for very long paths
        int64_t cap_scratch = 0;
        cudaEvent_t done = nullptr;
    } lane[2];
    cudaStream_t stream = nullptr;  // == lane[0].stream
    mcts_chunk_result *d_recs = nullptr;
    int64_t cap_recs = 0;
    LeafSched *d_sched = nullptr;
    int64_t sched_L = -1;
    uint32_t *d_flags = nullptr;  // replay: non-finite equity / profit seen (kernels.cuh PathArgs::flags)
    bool replay_attr[2] = {false, false};  // dynamic shared-memory limit of k_replay_paths<FULL / DEDUP> raised on this device
    // replay: two staging areas for caller sequences (chunk c+1 is copied while chunk c is evaluated)
    double *d_values[2] = {nullptr, nullptr};
    int64_t cap_values[2] = {0, 0};
    cudaEvent_t values_free[2] = {nullptr, nullptr};  // recorded after the kernels that read d_values[i]
    Stager stager;
    // event pool + timing of the last call
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<EventPair> ev_gen, ev_paths, ev_reduce;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    mcts_timing timing;
};

static int get_event(mcts_handle *h, cudaEvent_t *e)
{
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t n;
        CUDA_TRY(cudaEventCreate(&n));
        h->ev_pool.push_back(n);
    }
    *e = h->ev_pool[h->ev_used++];
    return MCTS_OK;
}

typedef mcts_handle::Lane Lane;

static int mark(mcts_handle *h, cudaStream_t stream, std::vector<EventPair> &v, bool begin)
{
    if (begin) {
        EventPair p;
        int rc = get_event(h, &p.a);
        if (rc) return rc;
        rc = get_event(h, &p.b);
        if (rc) return rc;
        v.push_back(p);
        CUDA_TRY(cudaEventRecord(p.a, stream));
    } else {
        CUDA_TRY(cudaEventRecord(v.back().b, stream));
    }
    return MCTS_OK;
}

static void timing_reset(mcts_handle *h)
{
    h->ev_used = 0;
    h->ev_gen.clear();
    h->ev_paths.clear();
    h->ev_reduce.clear();
    memset(&h->timing, 0, sizeof h->timing);
}

static int timing_collect(mcts_handle *h)
{
    auto total = [&](std::vector<EventPair> &v, double *ms) -> int {
        double acc = 0;
        for (auto &p : v) {
            float f = 0;
            CUDA_TRY(cudaEventElapsedTime(&f, p.a, p.b));
            acc += f;
        }
        *ms = acc;
        return MCTS_OK;
    };
    int rc;
    if ((rc = total(h->ev_gen, &h->timing.gen_ms))) return rc;
    if ((rc = total(h->ev_paths, &h->timing.paths_ms))) return rc;
    if ((rc = total(h->ev_reduce, &h->timing.reduce_ms))) return rc;
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, h->ev_begin, h->ev_end));
    h->timing.total_ms = f;
    h->timing.paths_launches = (int64_t)h->ev_paths.size();
    return MCTS_OK;
}

template <class T>
static int ensure(T **ptr, int64_t *cap, int64_t need, size_t elem = sizeof(T))
{
    if (need <= *cap) return MCTS_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)need * elem);
    if (e != cudaSuccess) {
        char buf[256];
        snprintf(buf, sizeof buf, "cudaMalloc of %lld bytes failed: %s", (long long)((size_t)need * elem), cudaGetErrorString(e));
        cudaGetLastError();
        return fail(MCTS_ENOMEM, buf);
    }
    *ptr = (T *)p;
    *cap = need;
    return MCTS_OK;
}

// streams, events and the leaf-schedule buffer of a fresh handle (mcts_create destroys the handle if this fails)
static int init_handle(mcts_handle *h)
{
    for (auto &L : h->lane) {
        CUDA_TRY(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
    }
    h->stream = h->lane[0].stream;
    CUDA_TRY(cudaEventCreate(&h->ev_begin));
    CUDA_TRY(cudaEventCreate(&h->ev_end));
    CUDA_TRY(cudaMalloc(&h->d_sched, sizeof(LeafSched)));
    CUDA_TRY(cudaMalloc(&h->d_flags, sizeof(uint32_t)));
    memset(&h->timing, 0, sizeof h->timing);
    return MCTS_OK;
}

// No C++ exception may cross the C ABI (ctypes / cgo callers cannot unwind it): allocation failures of the host-side
// containers and thread start-up failures become error codes like every other failure.
template <class F>
static int no_throw(F &&body) noexcept
{
    try {
        return body();
    } catch (const std::bad_alloc &) {
        try { return fail(MCTS_ENOMEM, "out of host memory"); } catch (...) { return MCTS_ENOMEM; }
    } catch (const std::exception &e) {
        try { return fail(MCTS_ECUDA, std::string("internal error: ") + e.what()); } catch (...) { return MCTS_ECUDA; }
    } catch (...) {
        return MCTS_ECUDA;
    }
}

extern "C" {

int mcts_abi_version(void) { return MCTS_ABI_VERSION; }

int mcts_generator_version(void) { return MCTS_GENERATOR_VERSION; }

const char *mcts_last_error(void) { return g_err.c_str(); }

static int mcts_create_impl(int device, mcts_handle **out)
{
    if (!out) return fail(MCTS_EINVAL, "mcts_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(MCTS_ENODEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libmcts has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(MCTS_ENODEVICE, "device index out of range");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char buf[512];
        snprintf(buf, sizeof buf, "device %d (%s) is compute capability %d.%d; libmcts is built for sm_100a only", device, prop.name, prop.major, prop.minor);
        return fail(MCTS_ENODEVICE, buf);
    }
    CUDA_TRY(cudaSetDevice(device));
    mcts_handle *h = new (std::nothrow) mcts_handle();
    if (!h) return fail(MCTS_ENOMEM, "out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    const int rc = init_handle(h);
    if (rc) {  // g_err already holds the message; release whatever was created
        mcts_destroy(h);
        return rc;
    }
    *out = h;
    return MCTS_OK;
}

void mcts_destroy(mcts_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto &L : h->lane) {
        if (L.stream) cudaStreamSynchronize(L.stream);
        cudaFree(L.d_counts);
        cudaFree(L.d_out);
        cudaFree(L.d_part);
        cudaFree(L.d_scratch);
        if (L.done) cudaEventDestroy(L.done);
    }
    cudaFree(h->d_recs);
    cudaFree(h->d_sched);
    cudaFree(h->d_flags);
    for (int i = 0; i < 2; i++) {
        cudaFree(h->d_values[i]);
        if (h->values_free[i]) cudaEventDestroy(h->values_free[i]);
    }
    h->stager.destroy();
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->ev_begin) cudaEventDestroy(h->ev_begin);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    for (auto &L : h->lane)
        if (L.stream) cudaStreamDestroy(L.stream);
    delete h;
}

int64_t mcts_sequence_length(const mcts_params *p)
{
    if (!p) return fail(MCTS_EINVAL, "params is NULL");
    GenSpec g;
    std::string err;
    mcts_params q = *p;
    if (q.num_simulations <= 0) q.num_simulations = 1;
    if (q.num_mc_shuffles <= 0) q.num_mc_shuffles = 1;
    int rc = build_spec(q, &g, &err);
    if (rc == MCTS_EEMPTY) return 0;
    if (rc) return fail(rc, err);
    return g.L;
}

int64_t mcts_plan_count(int64_t P, int64_t chunk_paths)
{
    if (P <= 0) return 0;
    std::vector<int64_t> f, c;
    plan_visit(0, P, clamp_chunk(chunk_paths), &f, &c);
    return (int64_t)f.size();
}

static int mcts_plan_impl(int64_t P, int64_t chunk_paths, int64_t *first_path, int64_t *n_paths)
{
    if (P <= 0) return fail(MCTS_EEMPTY, "no paths");
    std::vector<int64_t> f, c;
    plan_visit(0, P, clamp_chunk(chunk_paths), &f, &c);
    for (size_t i = 0; i < f.size(); i++) {
        first_path[i] = f[i];
        n_paths[i] = c[i];
    }
    return MCTS_OK;
}

static int mcts_combine_impl(int64_t P, int64_t chunk_paths, const mcts_chunk_result *chunks, int64_t n_chunks, mcts_agg *agg)
{
    if (P <= 0 || !chunks || !agg) return fail(MCTS_EINVAL, "mcts_combine: bad arguments");
    const int64_t chunk = clamp_chunk(chunk_paths);
    std::vector<int64_t> f, c;
    plan_visit(0, P, chunk, &f, &c);
    if ((int64_t)f.size() != n_chunks) return fail(MCTS_EINVAL, "mcts_combine: chunk count does not match the plan");
    std::vector<const mcts_chunk_result *> order(n_chunks);
    for (int64_t i = 0; i < n_chunks; i++) order[i] = &chunks[i];
    std::sort(order.begin(), order.end(), [](const mcts_chunk_result *a, const mcts_chunk_result *b) { return a->first_path < b->first_path; });
    for (int64_t i = 0; i < n_chunks; i++)
        if (order[i]->first_path != f[i] || order[i]->n_paths != c[i]) return fail(MCTS_EINVAL, "mcts_combine: chunk records do not tile the plan");
    std::vector<double> part(n_chunks);
    for (int s = 0; s < MCTS_NSTRAT; s++) {
        int64_t idx = 0;
        for (int64_t i = 0; i < n_chunks; i++) part[i] = order[i]->sum_profit[s];
        agg[s].sum_profit = 0.0 + combine_visit(P, chunk, part.data(), &idx);
        idx = 0;
        for (int64_t i = 0; i < n_chunks; i++) part[i] = order[i]->sum_dd[s];
        agg[s].sum_dd = 0.0 + combine_visit(P, chunk, part.data(), &idx);
        double mnp = order[0]->min_profit[s], mxp = order[0]->max_profit[s], mnd = order[0]->min_dd[s], mxd = order[0]->max_dd[s];
        for (int64_t i = 1; i < n_chunks; i++) {
            mnp = std::min(mnp, order[i]->min_profit[s]);
            mxp = std::max(mxp, order[i]->max_profit[s]);
            mnd = std::min(mnd, order[i]->min_dd[s]);
            mxd = std::max(mxd, order[i]->max_dd[s]);
        }
        agg[s].min_profit = mnp;
        agg[s].max_profit = mxp;
        agg[s].min_dd = mnd;
        agg[s].max_dd = mxd;
        agg[s].count = P;
    }
    return MCTS_OK;
}

int mcts_summary_rows(const mcts_agg *agg, int32_t num_trades, double *rows)
{
    if (!agg || !rows) return fail(MCTS_EINVAL, "mcts_summary_rows: NULL argument");
    for (int s = 0; s < MCTS_NSTRAT; s++) {
        const double n = (double)agg[s].count;
        const double avg_p = agg[s].sum_profit / n, avg_d = agg[s].sum_dd / n;  // np.mean = add.reduce / n
        double *r = rows + 9 * s;
        r[0] = avg_p;
        r[1] = avg_d;
        r[2] = avg_d != 0 ? avg_p / fabs(avg_d) : INFINITY;  // :520
        r[3] = agg[s].min_profit;
        r[4] = agg[s].max_profit;
        r[5] = agg[s].min_dd;
        r[6] = agg[s].max_dd;
        r[7] = avg_p / (double)num_trades;                                  // :519
        r[8] = agg[s].max_dd != 0 ? avg_p / fabs(agg[s].max_dd) : INFINITY;  // :521
    }
    return MCTS_OK;
}

int mcts_get_timing(const mcts_handle *h, mcts_timing *out)
{
    if (!h || !out) return fail(MCTS_EINVAL, "mcts_get_timing: NULL argument");
    *out = h->timing;
    return MCTS_OK;
}

void *mcts_get_stream(const mcts_handle *h) { return h ? (void *)h->stream : nullptr; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------

static int upload_sched(mcts_handle *h, int64_t L)
{
    if (h->sched_L == L) return MCTS_OK;
    LeafSched s;
    std::string err;
    int rc = build_leaf_sched(L, &s, &err);
    if (rc) return fail(rc, err);
    CUDA_TRY(cudaMemcpyAsync(h->d_sched, &s, sizeof s, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));  // `s` is a stack temporary
    h->timing.h2d_bytes += sizeof s;
    h->sched_L = L;
    return MCTS_OK;
}

static size_t path_smem_bytes(int L, int block) { return (size_t)kTabWords * 4 + (size_t)((L + 7) / 8) * block * 4; }

// defined in paths_<mode>.cu (one translation unit per PathMode so they compile in parallel)
cudaError_t mcts_launch_paths_full(const PathArgs &a, int block, cudaStream_t stream);
cudaError_t mcts_launch_paths_dedup(const PathArgs &a, int block, cudaStream_t stream);
cudaError_t mcts_launch_paths_curves(const PathArgs &a, int block, cudaStream_t stream);
cudaError_t mcts_launch_paths_int_full(const PathArgs &a, int block, cudaStream_t stream);
cudaError_t mcts_launch_paths_int_dedup(const PathArgs &a, int block, cudaStream_t stream);

// Block size: the largest of 128 / 64 / 32 that keeps at least 80 % of the best achievable number
// of threads resident per SM given the shared-memory code columns (4 bits per trade per thread,
// 227 KB per SM, 1 KB reserved per block).  Larger blocks are preferred because their warps stay in
// the same unrolled loop body (instruction-cache locality, see k_sim_paths).  `want`
// (mcts_options.block_threads) overrides for experiments.
static int pick_block(int L, int want)
{
#ifdef MCTS_BLOCK192
    if (want == 192 && path_smem_bytes(L, 192) * 2 <= 227 * 1024) return 192;
#endif
    const int cands[3] = {128, 64, 32};
    int resident[3] = {0, 0, 0}, best = 0;
    for (int i = 0; i < 3; i++) {
        const size_t smem = path_smem_bytes(L, cands[i]);
        if (smem > 227 * 1024) continue;
        if (want == cands[i]) return want;
        const int by_smem = (int)((size_t)228 * 1024 / (smem + 1024));
        resident[i] = std::min(kThreadsPerSM / cands[i], by_smem) * cands[i];
        best = std::max(best, resident[i]);
    }
    for (int i = 0; i < 3; i++)
        if (resident[i] > 0 && resident[i] * 5 >= best * 4) return cands[i];
    return 0;
}

// Paths too long for a useful shared-memory residency (fewer than 128 threads per SM, i.e. more than
// ~3,500 trades) keep their code columns in global memory instead.
static bool needs_global_codes(int L)
{
    int best = 0;
    for (int b : {128, 64, 32}) {
        const size_t smem = path_smem_bytes(L, b);
        if (smem > 227 * 1024) continue;
        best = std::max(best, std::min(kThreadsPerSM / b, (int)((size_t)228 * 1024 / (smem + 1024))) * b);
    }
    return best < 128;
}

static int launch_one(mcts_handle *h, Lane &L, int mode, const PathArgs &a, int block)
{
    cudaError_t e;
    if (block != 0 && h->int_fast_path && a.g.all_int && mode != PM_CURVES) {  // exact int32 evaluator, identical results
        e = mode == PM_DEDUP ? mcts_launch_paths_int_dedup(a, block, L.stream) : mcts_launch_paths_int_full(a, block, L.stream);
    } else {
        switch (mode) {
            case PM_DEDUP: e = mcts_launch_paths_dedup(a, block, L.stream); break;
            case PM_CURVES: e = mcts_launch_paths_curves(a, block, L.stream); break;
            default: e = mcts_launch_paths_full(a, block, L.stream); break;
        }
    }
    if (e != cudaSuccess) return fail(MCTS_ECUDA, std::string("k_sim_paths launch failed: ") + cudaGetErrorString(e));
    h->timing.launches++;
    return MCTS_OK;
}

static int launch_paths(mcts_handle *h, Lane &L, int mode, const PathArgs &a)
{
    if (!needs_global_codes(a.g.L)) {
        const int block = pick_block(a.g.L, h->block_threads);
        if (!block) return fail(MCTS_EUNSUPPORTED, "internal: no block size fits");
        return launch_one(h, L, mode, a, block);
    }
    if (mode == PM_CURVES) return fail(MCTS_EUNSUPPORTED, "curve export needs paths of at most ~3,500 trades");
    // global code store: [words][launch threads] u32, at most 1 GiB per launch -> sub-launches of the chunk
    const int64_t words = (a.g.L + 7) / 8;
    int64_t sub = (((int64_t)1 << 30) / (words * 4)) / 128 * 128;
    if (sub < 128) return fail(MCTS_EUNSUPPORTED, "num_trades too large (more than ~16 million trades per path)");
    sub = std::min<int64_t>(sub, (a.n + 127) / 128 * 128);
    int rc = ensure(&L.d_scratch, &L.cap_scratch, words * sub);
    if (rc) return rc;
    for (int64_t off = 0; off < a.n; off += sub) {
        PathArgs b = a;
        b.path0 = a.path0 + off;
        b.n = std::min<int64_t>(sub, a.n - off);
        b.out_off = a.out_off + off;
        b.scratch = L.d_scratch;
        b.scratch_stride = (int32_t)((b.n + 127) / 128 * 128);
        if ((rc = launch_one(h, L, mode, b, 0))) return rc;
    }
    return MCTS_OK;
}

// the evaluator over fp64 sequences values[a.n][T] already on the device: one launch, all passes (kernels.cuh k_replay_paths)
static int launch_replay_passes(mcts_handle *h, cudaStream_t stream, int mode, const PathArgs &a, const double *values, int64_t T)
{
    bool *attr_set = h->replay_attr;  // function attributes are per device: remembered per handle
    const unsigned grid = (unsigned)((a.n + kReplayBlock - 1) / kReplayBlock);
    const int m = mode == PM_DEDUP ? 1 : 0;
    if (!attr_set[m]) {  // 66 KB of dynamic shared memory per block
        CUDA_TRY(m ? cudaFuncSetAttribute(k_replay_paths<PM_DEDUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReplaySmem)
                   : cudaFuncSetAttribute(k_replay_paths<PM_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReplaySmem));
        attr_set[m] = true;
    }
    if (m) k_replay_paths<PM_DEDUP><<<grid, kReplayBlock, kReplaySmem, stream>>>(a, values, T);
    else k_replay_paths<PM_FULL><<<grid, kReplayBlock, kReplaySmem, stream>>>(a, values, T);
    CUDA_TRY(cudaGetLastError());
    h->timing.launches += 1;
    return MCTS_OK;
}

// Runs with more distinct values than 4-bit codes can name: shuffle fp64 sequences in global memory,
// then the replay evaluator.  Sub-batches of at most 1 GiB of values.
static int launch_wide(mcts_handle *h, Lane &L, int mode, const PathArgs &a)
{
    if (mode == PM_CURVES) return fail(MCTS_EUNSUPPORTED, "curve export needs at most 16 distinct trade values (8 per sign)");
    int64_t sub = (((int64_t)1 << 30) / ((int64_t)a.g.L * 8)) / 128 * 128;
    if (sub < 128) return fail(MCTS_EUNSUPPORTED, "num_trades too large for a run with more than 16 distinct trade values");
    sub = std::min<int64_t>(sub, (a.n + 127) / 128 * 128);
    int rc = ensure(&h->d_values[0], &h->cap_values[0], sub * a.g.L);
    if (rc) return rc;
    for (int64_t off = 0; off < a.n; off += sub) {
        PathArgs b = a;
        b.path0 = a.path0 + off;
        b.n = std::min<int64_t>(sub, a.n - off);
        b.out_off = a.out_off + off;
        k_shuffle_values<<<(unsigned)((b.n + 127) / 128), 128, 0, L.stream>>>(b, h->d_values[0]);
        h->timing.launches++;
        if ((rc = launch_replay_passes(h, L.stream, mode, b, h->d_values[0], a.g.L))) return rc;
    }
    return MCTS_OK;
}

static int launch_gen_base(mcts_handle *h, Lane &L, const GenSpec &g, uint64_t seed, int64_t sim0, int64_t n_sims)
{
    int rc = ensure(&L.d_counts, &L.cap_sims, n_sims, sizeof(uint32_t) * kMaxCat);
    if (rc) return rc;
    if ((rc = mark(h, L.stream, h->ev_gen, true))) return rc;
    k_gen_base<<<(unsigned)((n_sims + 127) / 128), 128, 0, L.stream>>>(g, seed, sim0, n_sims, L.d_counts);
    CUDA_TRY(cudaGetLastError());
    h->timing.launches++;
    return mark(h, L.stream, h->ev_gen, false);
}

// reduce the [40][ld] per-path rows of one chunk into one record
static int launch_reduce(mcts_handle *h, Lane &L, int64_t n, int64_t first_path, mcts_chunk_result *d_rec)
{
    const int d = reduce_depth(n);
    const int64_t nodes = (int64_t)1 << d;
    int rc = ensure(&L.d_part, &L.cap_part, std::max<int64_t>(nodes, 1024) * 160);  // 120 rows + 40 ping-pong rows
    if (rc) return rc;
    const int64_t pstride = L.cap_part / 160;
    if ((rc = mark(h, L.stream, h->ev_reduce, true))) return rc;
    dim3 grid((unsigned)((nodes + 7) / 8), 40);  // one 8-lane group per node
    k_reduce_nodes<<<grid, 64, 0, L.stream>>>(L.d_out, L.cap_paths, n, d, L.d_part, pstride);
    CUDA_TRY(cudaGetLastError());
    k_reduce_top<<<40, 64, 0, L.stream>>>(L.d_part, pstride, d, d_rec, first_path, n);
    CUDA_TRY(cudaGetLastError());
    h->timing.launches += 2;
    return mark(h, L.stream, h->ev_reduce, false);
}

static int ensure_out(Lane &L, int64_t max_n)
{
    if (max_n <= L.cap_paths) return MCTS_OK;
    return ensure(&L.d_out, &L.cap_paths, (max_n + 7) / 8 * 8, sizeof(double) * 40);  // rows stay 64-byte aligned
}

// lane 1 starts after everything already queued on lane 0 (the handle's stream) ...
static int fork_lanes(mcts_handle *h)
{
    CUDA_TRY(cudaEventRecord(h->ev_begin, h->lane[0].stream));
    CUDA_TRY(cudaStreamWaitEvent(h->lane[1].stream, h->ev_begin, 0));
    return MCTS_OK;
}
// ... and lane 0 ends after lane 1, so events recorded on the handle's stream bracket the whole call
static int join_lanes(mcts_handle *h)
{
    CUDA_TRY(cudaEventRecord(h->lane[1].done, h->lane[1].stream));
    CUDA_TRY(cudaStreamWaitEvent(h->lane[0].stream, h->lane[1].done, 0));
    CUDA_TRY(cudaEventRecord(h->ev_end, h->lane[0].stream));
    return MCTS_OK;
}

static int copy_paths_out(mcts_handle *h, Lane &L, int64_t first, int64_t n, int64_t P, double *profit, double *dd)
{
    if (profit) {
        CUDA_TRY(cudaMemcpy2DAsync(profit + first, (size_t)P * 8, L.d_out, (size_t)L.cap_paths * 8, (size_t)n * 8, 20, cudaMemcpyDeviceToHost, L.stream));
        h->timing.d2h_bytes += n * 8 * 20;
    }
    if (dd) {
        CUDA_TRY(cudaMemcpy2DAsync(dd + first, (size_t)P * 8, L.d_out + 20 * L.cap_paths, (size_t)L.cap_paths * 8, (size_t)n * 8, 20, cudaMemcpyDeviceToHost, L.stream));
        h->timing.d2h_bytes += n * 8 * 20;
    }
    return MCTS_OK;
}

static int simulate_range(mcts_handle *h, const mcts_params *p, uint64_t seed, const mcts_options *opt, int64_t cb, int64_t ce,
                          mcts_chunk_result *out, double *profit, double *dd)
{
    if (!h || !p) return fail(MCTS_EINVAL, "NULL handle or params");
    CUDA_TRY(cudaSetDevice(h->device));
    GenSpec g;
    std::string err;
    int rc = build_spec(*p, &g, &err);
    if (rc) return fail(rc, err);
    const int64_t P = (int64_t)p->num_simulations * p->num_mc_shuffles;
    const int64_t chunk = clamp_chunk(opt ? opt->chunk_paths : 0);
    const int mode = (opt && opt->dedup) ? PM_DEDUP : PM_FULL;
    h->block_threads = opt ? opt->block_threads : 0;
    h->int_fast_path = opt ? opt->int_fast_path : 0;
    std::vector<int64_t> first, count;
    plan_visit(0, P, chunk, &first, &count);
    if (cb < 0 || ce > (int64_t)first.size() || cb > ce) return fail(MCTS_EINVAL, "chunk range outside the plan");
    timing_reset(h);
    if ((rc = upload_sched(h, g.L))) return rc;
    int64_t max_n = 0;
    for (int64_t c = cb; c < ce; c++) max_n = std::max(max_n, count[c]);
    if (max_n == 0) return MCTS_OK;
    const int n_lanes = (opt && opt->serial_chunks) || ce - cb < 2 || g.wide ? 1 : 2;  // the wide fallback shares one values buffer
    for (int l = 0; l < n_lanes; l++)
        if ((rc = ensure_out(h->lane[l], max_n))) return rc;
    if ((rc = ensure(&h->d_recs, &h->cap_recs, ce - cb))) return rc;
    if ((rc = fork_lanes(h))) return rc;
    auto enqueue_chunks = [&]() -> int {
        for (int64_t c = cb; c < ce; c++) {
            Lane &L = h->lane[(c - cb) % n_lanes];
            const int64_t f = first[c], n = count[c];
            const int64_t s_lo = f / p->num_mc_shuffles, s_hi = (f + n - 1) / p->num_mc_shuffles;
            int r = launch_gen_base(h, L, g, seed, s_lo, s_hi - s_lo + 1);
            if (r) return r;
            PathArgs a;
            memset(&a, 0, sizeof a);
            a.g = g;
            a.seed = seed;
            a.path0 = f;
            a.n = n;
            a.sim0 = s_lo;
            a.n_sh = p->num_mc_shuffles;
            a.counts = L.d_counts;
            a.sched = h->d_sched;
            a.out = L.d_out;
            a.ld = L.cap_paths;
            if ((r = mark(h, L.stream, h->ev_paths, true))) return r;
            r = g.wide ? launch_wide(h, L, mode, a) : launch_paths(h, L, mode, a);
            if (r) return r;
            if ((r = mark(h, L.stream, h->ev_paths, false))) return r;
            if ((r = launch_reduce(h, L, n, f, h->d_recs + (c - cb)))) return r;
            if ((r = copy_paths_out(h, L, f, n, P, profit, dd))) return r;
            h->timing.trade_steps += n * (int64_t)g.L * MCTS_NSTRAT;
        }
        return join_lanes(h);
    };
    if ((rc = enqueue_chunks())) {
        // nothing may stay in flight on either lane (the caller's output buffers, the next call's events)
        const std::string msg = g_err;
        for (auto &L : h->lane) cudaStreamSynchronize(L.stream);
        cudaGetLastError();
        timing_reset(h);
        return fail(rc, msg);
    }
    CUDA_TRY(cudaMemcpyAsync(out, h->d_recs, sizeof(mcts_chunk_result) * (size_t)(ce - cb), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->timing.d2h_bytes += (int64_t)sizeof(mcts_chunk_result) * (ce - cb);
    h->timing.h2d_bytes += (int64_t)sizeof(PathArgs) * (ce - cb);
    return timing_collect(h);
}

extern "C" {

int mcts_simulate_chunks(mcts_handle *h, const mcts_params *p, uint64_t seed, const mcts_options *opt, int64_t chunk_begin,
                         int64_t chunk_end, mcts_chunk_result *out)
{
    if (!out) return fail(MCTS_EINVAL, "out is NULL");
    return simulate_range(h, p, seed, opt, chunk_begin, chunk_end, out, nullptr, nullptr);
}

static int mcts_simulate_impl(mcts_handle *h, const mcts_params *p, uint64_t seed, const mcts_options *opt, mcts_agg *agg, double *profit, double *dd)
{
    if (!h || !p || !agg) return fail(MCTS_EINVAL, "NULL argument");
    const int64_t P = (int64_t)p->num_simulations * p->num_mc_shuffles;
    if (P <= 0) return fail(MCTS_EEMPTY, "zero paths (np.min of an empty list raises in the reference)");
    const int64_t chunk = opt ? opt->chunk_paths : 0;
    const int64_t nc = mcts_plan_count(P, chunk);
    std::vector<mcts_chunk_result> recs((size_t)nc);
    int rc = simulate_range(h, p, seed, opt, 0, nc, recs.data(), profit, dd);
    if (rc) return rc;
    return mcts_combine(P, chunk, recs.data(), nc, agg);
}

static int mcts_replay_impl(mcts_handle *h, const double *values, int64_t P, int64_t T, const mcts_options *opt, mcts_agg *agg, double *profit, double *dd)
{
    if (!h || !values || !agg) return fail(MCTS_EINVAL, "NULL argument");
    if (P <= 0 || T <= 0) return fail(MCTS_EEMPTY, "zero paths or an empty trade sequence");
    if (T > (int64_t)1 << 20) return fail(MCTS_EUNSUPPORTED, "replay sequences longer than 2^20 trades");
    CUDA_TRY(cudaSetDevice(h->device));
    int64_t chunk = clamp_chunk(opt ? opt->chunk_paths : 0);
    chunk = std::max<int64_t>(128, std::min<int64_t>(chunk, ((int64_t)1 << 27) / T));  // <= 1 GiB of values per chunk
    const int mode = (opt && opt->dedup) ? PM_DEDUP : PM_FULL;
    std::vector<int64_t> first, count;
    plan_visit(0, P, chunk, &first, &count);
    const int64_t nc = (int64_t)first.size();
    timing_reset(h);
    int rc;
    if ((rc = upload_sched(h, T))) return rc;
    int64_t max_n = *std::max_element(count.begin(), count.end());
    Lane &L = h->lane[0];
    if ((rc = ensure_out(L, max_n))) return rc;
    const int n_buf = nc > 1 ? 2 : 1;
    for (int i = 0; i < n_buf; i++) {
        if ((rc = ensure(&h->d_values[i], &h->cap_values[i], max_n * T))) return rc;
        if (!h->values_free[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->values_free[i], cudaEventDisableTiming));
    }
    if (!h->stager.ready()) CUDA_TRY(h->stager.init(h->device));
    if ((rc = ensure(&h->d_recs, &h->cap_recs, nc))) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, sizeof(uint32_t), h->stream));
    CUDA_TRY(cudaEventRecord(h->ev_begin, h->stream));
    for (int64_t c = 0; c < nc; c++) {
        const int64_t f = first[c], n = count[c];
        double *d_vals = h->d_values[c % n_buf];
        // pageable caller memory -> pinned slices -> device, by the stager's worker threads (stager.hpp)
        CUDA_TRY(h->stager.copy(d_vals, values + f * T, (size_t)n * T * 8, c >= n_buf ? h->values_free[c % n_buf] : nullptr, h->stream));
        h->timing.h2d_bytes += n * T * 8;
        PathArgs a;
        memset(&a, 0, sizeof a);
        a.g.L = (int32_t)T;
        a.path0 = f;
        a.n = n;
        a.sched = h->d_sched;
        a.out = L.d_out;
        a.ld = L.cap_paths;
        a.flags = h->d_flags;
        if ((rc = mark(h, L.stream, h->ev_paths, true))) return rc;
        if ((rc = launch_replay_passes(h, h->stream, mode, a, d_vals, T))) return rc;
        if ((rc = mark(h, L.stream, h->ev_paths, false))) return rc;
        CUDA_TRY(cudaEventRecord(h->values_free[c % n_buf], h->stream));
        if ((rc = launch_reduce(h, L, n, f, h->d_recs + c))) return rc;
        if ((rc = copy_paths_out(h, L, f, n, P, profit, dd))) return rc;
        h->timing.trade_steps += n * T * MCTS_NSTRAT;
    }
    CUDA_TRY(cudaEventRecord(h->ev_end, h->stream));
    std::vector<mcts_chunk_result> recs((size_t)nc);
    CUDA_TRY(cudaMemcpyAsync(recs.data(), h->d_recs, sizeof(mcts_chunk_result) * (size_t)nc, cudaMemcpyDeviceToHost, h->stream));
    uint32_t flags = 0;
    CUDA_TRY(cudaMemcpyAsync(&flags, h->d_flags, sizeof flags, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->timing.d2h_bytes += (int64_t)sizeof(mcts_chunk_result) * nc + (int64_t)sizeof flags;
    if ((rc = timing_collect(h))) return rc;
    if (flags & 1u)
        return fail(MCTS_EUNSUPPORTED, "replay: a trade value is NaN or infinite, or an equity / profit sum exceeded DBL_MAX; the reference yields "
                                       "NaN drawdowns there (np.min propagates NaN), which this build does not reproduce");
    return mcts_combine(P, chunk, recs.data(), nc, agg);
}

static int path_range_args(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n, PathArgs *a)
{
    GenSpec g;
    std::string err;
    int rc = build_spec(*p, &g, &err);
    if (rc) return fail(rc, err);
    const int64_t P = (int64_t)p->num_simulations * p->num_mc_shuffles;
    if (first_path < 0 || n <= 0 || first_path + n > P) return fail(MCTS_EINVAL, "path range outside [0, num_simulations*num_mc_shuffles)");
    timing_reset(h);
    if ((rc = upload_sched(h, g.L))) return rc;
    const int64_t s_lo = first_path / p->num_mc_shuffles, s_hi = (first_path + n - 1) / p->num_mc_shuffles;
    CUDA_TRY(cudaEventRecord(h->ev_begin, h->stream));
    if ((rc = launch_gen_base(h, h->lane[0], g, seed, s_lo, s_hi - s_lo + 1))) return rc;
    memset(a, 0, sizeof *a);
    a->g = g;
    a->seed = seed;
    a->path0 = first_path;
    a->n = n;
    a->sim0 = s_lo;
    a->n_sh = p->num_mc_shuffles;
    a->counts = h->lane[0].d_counts;
    a->sched = h->d_sched;
    return MCTS_OK;
}

static int mcts_sequences_impl(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n, double *values)
{
    if (!h || !p || !values) return fail(MCTS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    PathArgs a;
    int rc = path_range_args(h, p, seed, first_path, n, &a);
    if (rc) return rc;
    const int L = a.g.L;
    if ((rc = ensure(&h->d_values[0], &h->cap_values[0], n * L))) return rc;
    if (a.g.wide) {
        k_shuffle_values<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a, h->d_values[0]);
        CUDA_TRY(cudaGetLastError());
        h->timing.launches++;
        CUDA_TRY(cudaEventRecord(h->ev_end, h->stream));
        CUDA_TRY(cudaMemcpyAsync(values, h->d_values[0], (size_t)n * L * 8, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->timing.d2h_bytes += n * L * 8;
        return timing_collect(h);
    }
    size_t smem = (size_t)((L + 7) / 8) * kBlock * 4;
    const bool global = smem > 227 * 1024;
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    if (global) {
        Lane &L0 = h->lane[0];
        a.scratch_stride = (int32_t)(grid * kBlock);
        if ((int64_t)((L + 7) / 8) * a.scratch_stride * 4 > ((int64_t)1 << 31)) return fail(MCTS_EUNSUPPORTED, "mcts_sequences: request fewer paths of this length per call");
        if ((rc = ensure(&L0.d_scratch, &L0.cap_scratch, (int64_t)((L + 7) / 8) * a.scratch_stride))) return rc;
        a.scratch = L0.d_scratch;
        smem = 0;
    }
#define LAUNCH_SEQ(KK)                                                                                                        \
    do {                                                                                                                      \
        if (global) {                                                                                                         \
            k_sequences<KK, true><<<grid, kBlock, 0, h->stream>>>(a, h->d_values[0]);                                            \
        } else {                                                                                                              \
            if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_sequences<KK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
            k_sequences<KK, false><<<grid, kBlock, smem, h->stream>>>(a, h->d_values[0]);                                        \
        }                                                                                                                     \
    } while (0)
    switch (urn_width(a.g.K)) {  // the same generator the path kernels use for this run
        case 2: LAUNCH_SEQ(2); break;
        case 5: LAUNCH_SEQ(5); break;
        case 6: LAUNCH_SEQ(6); break;
        default: LAUNCH_SEQ(16); break;
    }
#undef LAUNCH_SEQ
    CUDA_TRY(cudaGetLastError());
    h->timing.launches++;
    CUDA_TRY(cudaEventRecord(h->ev_end, h->stream));
    CUDA_TRY(cudaMemcpyAsync(values, h->d_values[0], (size_t)n * L * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->timing.d2h_bytes += n * L * 8;
    return timing_collect(h);
}

static int mcts_export_curves_impl(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n, double *curves_dev)
{
    if (!h || !p || !curves_dev) return fail(MCTS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    PathArgs a;
    int rc = path_range_args(h, p, seed, first_path, n, &a);
    if (rc) return rc;
    if (a.g.wide) return fail(MCTS_EUNSUPPORTED, "curve export needs at most 16 distinct trade values (8 per sign)");
    a.curves = curves_dev;
    a.out = nullptr;
    if ((rc = mark(h, h->stream, h->ev_paths, true))) return rc;
    if ((rc = launch_paths(h, h->lane[0], PM_CURVES, a))) return rc;
    if ((rc = mark(h, h->stream, h->ev_paths, false))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_end, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->timing.trade_steps = n * (int64_t)a.g.L * MCTS_NSTRAT;
    return timing_collect(h);
}

void *mcts_device_alloc(mcts_handle *h, int64_t bytes)
{
    if (!h || bytes <= 0) return nullptr;
    cudaSetDevice(h->device);
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        fail(MCTS_ENOMEM, "cudaMalloc failed");
        return nullptr;
    }
    return p;
}

void mcts_device_free(mcts_handle *h, void *ptr)
{
    if (!h || !ptr) return;
    cudaSetDevice(h->device);
    cudaFree(ptr);
}

int mcts_copy_to_host(mcts_handle *h, void *dst_host, const void *src_dev, int64_t bytes)
{
    if (!h || !dst_host || !src_dev || bytes < 0) return fail(MCTS_EINVAL, "bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MCTS_OK;
}

int mcts_measure_step_ceiling(mcts_handle *h, double *strategy_steps_per_s)
{
    if (!h || !strategy_steps_per_s) return fail(MCTS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = h->sm_count * 3 * 4, steps8 = 2000;
    double *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, 64));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {  // last repetition counts
        CUDA_TRY(cudaEventRecord(a, h->stream));
        k_peak_step<<<blocks, 128, 0, h->stream>>>(d_out, steps8, 186.0, 372.0, (unsigned)rep);
        CUDA_TRY(cudaEventRecord(b, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    }
    *strategy_steps_per_s = (double)blocks * 128 * steps8 * 8 * 5 / (ms * 1e-3);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    return MCTS_OK;
}

int mcts_measure_mix_ceiling(mcts_handle *h, double *strategy_steps_per_s, double *ipc)
{
    if (!h || !strategy_steps_per_s) return fail(MCTS_EINVAL, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = h->sm_count * 3 * 4, steps = 40000, per_step = 13;
    double *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, 64));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {  // last repetition counts
        CUDA_TRY(cudaEventRecord(a, h->stream));
        k_peak_mix<<<blocks, 128, 0, h->stream>>>(d_out, steps, 1.0000001, 0x9e3779b9u, 0x7f4a7c15u);
        CUDA_TRY(cudaEventRecord(b, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    }
    const double warp_steps = (double)blocks * 4 * steps;  // one "strategy-step" of a warp = 13 instructions
    *strategy_steps_per_s = warp_steps * 32.0 / (ms * 1e-3);
    if (ipc) {
        int khz = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
        *ipc = warp_steps * per_step / ((double)h->sm_count * 4.0 * ms * 1e-3 * khz * 1e3);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    return MCTS_OK;
}

int mcts_measure_peaks(mcts_handle *h, double *fp64_inst_per_s, double *int_inst_per_s, double *sm_clock_mhz)
{
    if (!h) return fail(MCTS_EINVAL, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const int blocks = h->sm_count * 8, threads = 256, iters = 4096;
    double *d_out = nullptr;
    unsigned long long *d_clk = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, 64));
    CUDA_TRY(cudaMalloc(&d_clk, sizeof(unsigned long long) * 2 * blocks));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    float ms_d = 0, ms_i = 0;
    for (int rep = 0; rep < 3; rep++) {  // last repetition counts
        CUDA_TRY(cudaEventRecord(a, h->stream));
        k_peak_dadd<<<blocks, threads, 0, h->stream>>>(d_out, iters, 1.0000001, d_clk);
        CUDA_TRY(cudaEventRecord(b, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaEventElapsedTime(&ms_d, a, b));
        CUDA_TRY(cudaEventRecord(a, h->stream));
        k_peak_int<<<blocks, threads, 0, h->stream>>>((unsigned *)d_out, iters, 0x9e3779b9u, 0x7f4a7c15u);
        CUDA_TRY(cudaEventRecord(b, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaEventElapsedTime(&ms_i, a, b));
    }
    std::vector<unsigned long long> clk(2 * (size_t)blocks);
    CUDA_TRY(cudaMemcpy(clk.data(), d_clk, sizeof(unsigned long long) * clk.size(), cudaMemcpyDeviceToHost));
    std::vector<double> mhz;
    for (int i = 0; i < blocks; i++)
        if (clk[2 * i + 1] > 0) mhz.push_back((double)clk[2 * i] / (double)clk[2 * i + 1] * 1e3);
    std::sort(mhz.begin(), mhz.end());
    const double ops = (double)blocks * threads * (double)iters * 64.0;
    if (fp64_inst_per_s) *fp64_inst_per_s = ops / (ms_d * 1e-3);
    if (int_inst_per_s) *int_inst_per_s = 2.0 * ops / (ms_i * 1e-3);
    if (sm_clock_mhz) *sm_clock_mhz = mhz.empty() ? 0.0 : mhz[mhz.size() / 2];
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    cudaFree(d_clk);
    h->timing.launches = 6;
    return MCTS_OK;
}


// ---- several devices behind one call ----------------------------------------------------------

struct mcts_group {
    std::vector<mcts_handle *> dev;
    std::vector<mcts_timing> timing;
};

int64_t mcts_auto_chunk(int64_t P, int n_shards)
{
    if (n_shards <= 1) return clamp_chunk(0);
    int64_t c = P / (4 * (int64_t)n_shards);  // >= 4 chunks per shard: two per stream lane, and runs balance to within ~10 %
    c = std::max<int64_t>(c, 16384);  // below this a launch no longer fills one GPU
    return clamp_chunk(std::min<int64_t>(c, (int64_t)1 << 22));
}

static int mcts_plan_shards_impl(int64_t P, int64_t chunk_paths, int n_shards, int64_t *bounds)
{
    if (P <= 0) return fail(MCTS_EEMPTY, "no paths");
    if (n_shards <= 0 || !bounds) return fail(MCTS_EINVAL, "mcts_plan_shards: bad arguments");
    std::vector<int64_t> first, count;
    plan_visit(0, P, clamp_chunk(chunk_paths), &first, &count);
    for (int d = 0; d <= n_shards; d++) bounds[d] = (int64_t)count.size();
    bounds[0] = 0;
    int64_t acc = 0;
    int d = 1;
    for (size_t i = 0; i < count.size() && d < n_shards; i++) {
        // chunk i belongs to the shard whose quantile of the path range holds the chunk's midpoint
        while (d < n_shards && (acc + count[i] / 2) * n_shards >= P * d) bounds[d++] = (int64_t)i;
        acc += count[i];
    }
    return MCTS_OK;
}

static int mcts_group_create_impl(const int *devices, int n, mcts_group **out)
{
    if (!out) return fail(MCTS_EINVAL, "mcts_group_create: out is NULL");
    *out = nullptr;
    std::vector<int> ids;
    if (!devices || n <= 0) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0) {
            cudaGetLastError();
            return fail(MCTS_ENODEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libmcts has no CPU fallback");
        }
        for (int d = 0; d < count; d++) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ids.push_back(d);
        }
        if (ids.empty()) return fail(MCTS_ENODEVICE, "no compute-capability-10.x device visible; libmcts is built for sm_100a only");
    } else {
        ids.assign(devices, devices + n);  // a device may appear more than once: every entry gets its own handle and streams
    }
    mcts_group *g = new (std::nothrow) mcts_group();
    if (!g) return fail(MCTS_ENOMEM, "out of host memory");
    for (int d : ids) {
        mcts_handle *h = nullptr;
        const int rc = mcts_create(d, &h);
        if (rc) {
            mcts_group_destroy(g);
            return rc;
        }
        g->dev.push_back(h);
    }
    g->timing.assign(g->dev.size(), mcts_timing());
    *out = g;
    return MCTS_OK;
}

void mcts_group_destroy(mcts_group *g)
{
    if (!g) return;
    for (mcts_handle *h : g->dev) mcts_destroy(h);
    delete g;
}

int mcts_group_size(const mcts_group *g) { return g ? (int)g->dev.size() : 0; }

int mcts_group_device(const mcts_group *g, int i) { return (g && i >= 0 && i < (int)g->dev.size()) ? g->dev[i]->device : -1; }

int mcts_group_get_timing(const mcts_group *g, mcts_timing *out)
{
    if (!g || !out) return fail(MCTS_EINVAL, "mcts_group_get_timing: NULL argument");
    for (size_t i = 0; i < g->timing.size(); i++) out[i] = g->timing[i];
    return MCTS_OK;
}

static int mcts_group_simulate_impl(mcts_group *g, const mcts_params *p, uint64_t seed, const mcts_options *opt, mcts_agg *agg, double *profit, double *dd)
{
    if (!g || g->dev.empty() || !p || !agg) return fail(MCTS_EINVAL, "NULL argument");
    const int64_t P = (int64_t)p->num_simulations * p->num_mc_shuffles;
    if (P <= 0) return fail(MCTS_EEMPTY, "zero paths (np.min of an empty list raises in the reference)");
    const int n_dev = (int)g->dev.size();
    mcts_options o;
    memset(&o, 0, sizeof o);
    if (opt) o = *opt;
    o.chunk_paths = o.chunk_paths > 0 ? clamp_chunk(o.chunk_paths) : mcts_auto_chunk(P, n_dev);
    const int64_t nc = mcts_plan_count(P, o.chunk_paths);
    std::vector<int64_t> bounds((size_t)n_dev + 1);
    int rc0 = mcts_plan_shards(P, o.chunk_paths, n_dev, bounds.data());
    if (rc0) return rc0;
    std::vector<mcts_chunk_result> recs((size_t)nc);
    std::vector<int> rcs(n_dev, MCTS_OK);
    std::vector<std::string> msgs(n_dev);
    auto work = [&](int d) {
        memset(&g->timing[d], 0, sizeof(mcts_timing));
        if (bounds[d] >= bounds[d + 1]) return;
        rcs[d] = simulate_range(g->dev[d], p, seed, &o, bounds[d], bounds[d + 1], recs.data() + bounds[d], profit, dd);
        if (rcs[d]) msgs[d] = g_err;  // thread-local: hand the message to the calling thread
        else g->timing[d] = g->dev[d]->timing;
    };
    std::vector<std::thread> threads;
    for (int d = 1; d < n_dev; d++) {
        try {
            threads.emplace_back(work, d);
        } catch (...) {  // no thread: evaluate that device's run from this thread after the others
            rcs[d] = 1;
        }
    }
    work(0);
    for (auto &t : threads) t.join();
    for (int d = 1; d < n_dev; d++)
        if (rcs[d] == 1) {
            rcs[d] = MCTS_OK;
            work(d);
        }
    for (int d = 0; d < n_dev; d++)
        if (rcs[d]) {
            char buf[64];
            snprintf(buf, sizeof buf, "device %d: ", g->dev[d]->device);
            return fail(rcs[d], buf + msgs[d]);
        }
    return mcts_combine(P, o.chunk_paths, recs.data(), nc, agg);
}

// ---- the exported entry points of everything above that allocates on the host ---------------------

int mcts_create(int device, mcts_handle **out)
{
    return no_throw([&] { return mcts_create_impl(device, out); });
}

int mcts_plan(int64_t P, int64_t chunk_paths, int64_t *first_path, int64_t *n_paths)
{
    return no_throw([&] { return mcts_plan_impl(P, chunk_paths, first_path, n_paths); });
}

int mcts_combine(int64_t P, int64_t chunk_paths, const mcts_chunk_result *chunks, int64_t n_chunks, mcts_agg *agg)
{
    return no_throw([&] { return mcts_combine_impl(P, chunk_paths, chunks, n_chunks, agg); });
}

int mcts_simulate(mcts_handle *h, const mcts_params *p, uint64_t seed, const mcts_options *opt, mcts_agg *agg, double *profit, double *dd)
{
    return no_throw([&] { return mcts_simulate_impl(h, p, seed, opt, agg, profit, dd); });
}

int mcts_replay(mcts_handle *h, const double *values, int64_t P, int64_t T, const mcts_options *opt, mcts_agg *agg, double *profit, double *dd)
{
    return no_throw([&] { return mcts_replay_impl(h, values, P, T, opt, agg, profit, dd); });
}

int mcts_sequences(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n, double *values)
{
    return no_throw([&] { return mcts_sequences_impl(h, p, seed, first_path, n, values); });
}

int mcts_export_curves(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n, double *curves_dev)
{
    return no_throw([&] { return mcts_export_curves_impl(h, p, seed, first_path, n, curves_dev); });
}

int mcts_plan_shards(int64_t P, int64_t chunk_paths, int n_shards, int64_t *bounds)
{
    return no_throw([&] { return mcts_plan_shards_impl(P, chunk_paths, n_shards, bounds); });
}

int mcts_group_create(const int *devices, int n, mcts_group **out)
{
    return no_throw([&] { return mcts_group_create_impl(devices, n, out); });
}

int mcts_group_simulate(mcts_group *g, const mcts_params *p, uint64_t seed, const mcts_options *opt, mcts_agg *agg, double *profit, double *dd)
{
    return no_throw([&] { return mcts_group_simulate_impl(g, p, seed, opt, agg, profit, dd); });
}

}  // extern "C"
