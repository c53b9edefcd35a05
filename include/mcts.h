/*
 * mcts.h -- C ABI of libmcts.so, the B200 (sm_100a) Monte Carlo path-simulation
 * core that replaces the hot path of jensbrodersen/mc-trade-sizing:
 *
 *     run_all_strategies()            src/trading_models.py:445-530
 *       simulate_trades_dynamic       src/trading_models.py:43-66
 *       simulate_trades_markov        src/trading_models.py:68-85
 *       simulate_trades_markov2       src/trading_models.py:87-104
 *       simulate_trades_regime_switch src/trading_models.py:106-128
 *       np.random.shuffle per path    src/trading_models.py:491-492
 *       strategy_static / _dynamic    src/trading_models.py:135-168
 *       make_condition_func (20 ids)  src/trading_models.py:170-404
 *       calculate_drawdown            src/trading_models.py:130-133
 *       per-strategy aggregates       src/trading_models.py:509-522
 *
 * The reference is pure Python and has no FFI; the binding a maintainer adds is
 * the ctypes stub in INTEGRATION.md (shipped as mc_trade_sizing_b200/_ffi.py).
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative mcts_status; the
 *     message is available from mcts_last_error() (thread-local);
 *   - the caller allocates every output buffer; the library never keeps a
 *     caller pointer after the call returns; all pointers are HOST pointers
 *     unless a parameter says "device";
 *   - there is no CPU fallback: without a usable CUDA device mcts_create fails;
 *   - a handle is bound to one device and must not be used concurrently from
 *     two threads; distinct handles (and distinct processes) are independent.
 */
#ifndef MCTS_H
#define MCTS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCTS_NSTRAT 20
#define MCTS_MAX_REGIMES 16
#define MCTS_ABI_VERSION 2
/* Version of the on-device generator's published stream: results of mcts_simulate are a pure function
 * of (parameters, seed, MCTS_GENERATOR_VERSION) -- independent of chunking, of the number of devices and of the build.  A change
 * of the stream format (3: eight base draws per Philox block + urn shuffle) bumps it; replay results never depend on it. */
#define MCTS_GENERATOR_VERSION 3

typedef enum {
    MCTS_OK = 0,
    MCTS_EINVAL = -1,     /* invalid argument (NumPy would raise ValueError: bad probabilities ...) */
    MCTS_EEMPTY = -2,     /* zero paths or an empty trade sequence (np.min([]) raises in the reference) */
    MCTS_ENOMEM = -3,     /* host or device allocation failed */
    MCTS_ECUDA = -4,      /* CUDA runtime error, message has the details */
    MCTS_ENODEVICE = -5,  /* no usable CUDA device / wrong architecture: there is no CPU fallback */
    MCTS_EUNSUPPORTED = -6 /* valid for the reference but outside this build's limits (see message):
                              > ~260,000 trades per path, > 16 regimes, subnormal trade values in simulate
                              mode, curve export of runs with > 16 distinct values, and NON-FINITE ARITHMETIC:
                              NaN / infinite trade values, or equity / profit sums beyond DBL_MAX (simulate:
                              3 * max|value| * trades >= DBL_MAX is refused up front; replay: detected on the
                              device).  The reference reports NaN drawdowns in those cases (np.min propagates
                              NaN); the kernels' ordered comparisons would drop the NaN, so the call fails
                              instead of returning numbers that differ from the reference's. */
} mcts_status;

/* Generator selection, same precedence as trading_models.py:477-490. */
typedef enum {
    MCTS_MODE_DYNAMIC = 0, /* "No Markov": simulate_trades_dynamic, :43-66 */
    MCTS_MODE_MARKOV1 = 1, /* simulate_trades_markov, :68-85 */
    MCTS_MODE_MARKOV2 = 2, /* simulate_trades_markov2, :87-104 */
    MCTS_MODE_REGIME = 3   /* simulate_trades_regime_switch, :106-128 */
} mcts_mode;

typedef struct {
    int32_t length;
    double hit_rate, avg_win, avg_loss;
} mcts_regime;

/* The argument list of run_all_strategies (:445-450), flattened. */
typedef struct {
    int32_t mode; /* mcts_mode */
    int32_t num_trades;
    int32_t num_simulations;
    int32_t num_mc_shuffles;
    double hit_rate, avg_win, avg_loss;
    double p_win_after_win, p_win_after_loss;
    double p_win_ww, p_win_wl, p_win_lw, p_win_ll;
    int32_t n_regimes; /* < 0: regimes=None, the defaults of :108-112 */
    mcts_regime regimes[MCTS_MAX_REGIMES];
} mcts_params;

/* Per-strategy aggregate over paths (:513-518). sum_* follow np.mean's pairwise
 * summation order over the path index (sim-major, shuffle-minor, :507), so
 * sum_profit / count reproduces np.mean bit for bit. */
typedef struct {
    double sum_profit, sum_dd;
    double min_profit, max_profit;
    double min_dd, max_dd; /* min_dd = most negative, max_dd = least negative (:517-518) */
    int64_t count;
} mcts_agg;

typedef struct {
    int64_t chunk_paths;   /* max paths per device chunk (0 = default 1<<22); chunks are nodes of
                              NumPy's pairwise-summation tree so partial sums combine exactly */
    int32_t dedup;         /* 0 = evaluate all 20 state machines (default); 1 = evaluate the 11
                              distinct behaviours once and fan the results out (identical output) */
    int32_t int_fast_path; /* 0 = fp64 evaluator always; 1 = allow the exact int32 evaluator when
                              every trade value is an integer and sums stay below 2^31 */
    int32_t block_threads; /* 0 = default */
    int32_t serial_chunks; /* 0 = default: consecutive chunks run on two streams so the reduction and the
                              wave tail of one overlap the path kernel of the next; 1 = one stream (clean
                              per-kernel event timing for roofline measurements) */
} mcts_options;

/* Device-side timing of the last call on a handle (CUDA events on the library's stream). */
typedef struct {
    double total_ms;    /* first launch .. last kernel of the call */
    double gen_ms;      /* base multiset generation (per simulation) */
    double paths_ms;    /* the dominant kernel: shuffle + 20 strategies per path */
    double reduce_ms;   /* exact-order pairwise reduction of per-path results */
    int64_t launches;   /* kernels launched by the call */
    int64_t paths_launches;
    int64_t trade_steps; /* paths x trades x 20 */
    int64_t h2d_bytes, d2h_bytes;
} mcts_timing;

/* One node of the pairwise tree over paths, evaluated on one device (multi-GPU exchange record). */
typedef struct {
    int64_t first_path, n_paths;
    double sum_profit[MCTS_NSTRAT], sum_dd[MCTS_NSTRAT];
    double min_profit[MCTS_NSTRAT], max_profit[MCTS_NSTRAT];
    double min_dd[MCTS_NSTRAT], max_dd[MCTS_NSTRAT];
} mcts_chunk_result;

typedef struct mcts_handle mcts_handle;

int mcts_abi_version(void);
int mcts_generator_version(void);
const char *mcts_last_error(void);

/* Bind a handle to CUDA device `device` (>= 0). Fails with MCTS_ENODEVICE if the device is
 * absent or is not compute capability 10.x. */
int mcts_create(int device, mcts_handle **out);
void mcts_destroy(mcts_handle *h);

/* Length of the trade sequence the selected generator yields (not always num_trades:
 * Markov-2 emits >= 2 trades, regime lists can be short; :89-91, :116-126). */
int64_t mcts_sequence_length(const mcts_params *p);

/*
 * run_all_strategies with on-device RNG (replaces :476-522).
 *   seed      Philox4x32-10 key; draws are keyed by (simulation, shuffle, trade), so results
 *             do not depend on chunking or on how paths are split over devices
 *   agg       [20] aggregates in strategy-id order (id 1 first)
 *   profit,dd optional [20][P] per-path results, P = num_simulations*num_mc_shuffles (NULL to skip)
 */
int mcts_simulate(mcts_handle *h, const mcts_params *p, uint64_t seed, const mcts_options *opt,
                  mcts_agg *agg, double *profit, double *dd);

/*
 * Replay: evaluate caller-supplied post-shuffle sequences (replaces :493-507 + :509-518).
 *   values    [P][T] row-major trade results (what np.random.shuffle left in base_results); any finite
 *             doubles, including zeros of either sign and the reference's int64 regime arrays; a NaN or
 *             infinite value, or sums beyond DBL_MAX, fail the call with MCTS_EUNSUPPORTED (see there)
 *   profit,dd optional [20][P]
 */
int mcts_replay(mcts_handle *h, const double *values, int64_t P, int64_t T, const mcts_options *opt,
                mcts_agg *agg, double *profit, double *dd);

/* The post-shuffle sequences the device generator produces for paths [first_path, first_path+n),
 * as doubles [n][L], L = mcts_sequence_length(p).  Test / audit hook: feeding these to the
 * reference evaluator must reproduce mcts_simulate's per-path results bit for bit. */
int mcts_sequences(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n,
                   double *values);

/* --- multi-device: paths shard by pairwise-tree node ------------------------------------- */

/* Number of chunks the path range [0,P) splits into for a given chunk size. */
int64_t mcts_plan_count(int64_t P, int64_t chunk_paths);
/* first_path/n_paths of every chunk, in path order. */
int mcts_plan(int64_t P, int64_t chunk_paths, int64_t *first_path, int64_t *n_paths);
/* Evaluate chunks [chunk_begin, chunk_end) of the plan on this handle's device. */
int mcts_simulate_chunks(mcts_handle *h, const mcts_params *p, uint64_t seed, const mcts_options *opt,
                         int64_t chunk_begin, int64_t chunk_end, mcts_chunk_result *out);
/* Combine the results of ALL chunks of the plan (any device order, sorted here by first_path)
 * in NumPy's tree order.  Pure host arithmetic on ~1 KB records. */
int mcts_combine(int64_t P, int64_t chunk_paths, const mcts_chunk_result *chunks, int64_t n_chunks,
                 mcts_agg *agg);

/* Chunk size used when the caller gives none and the plan is to be split over `n_shards` devices / ranks:
 * every shard gets at least four chunks whenever there are enough paths (>= 16,384 per chunk), never more than
 * the single-device default of 1<<22.  Results do not depend on the chunk size. */
int64_t mcts_auto_chunk(int64_t P, int n_shards);
/* Contiguous runs of chunks with (nearly) equal path counts: shard d evaluates chunks
 * [bounds[d], bounds[d+1]) of the plan for (P, chunk_paths); bounds has n_shards + 1 entries. */
int mcts_plan_shards(int64_t P, int64_t chunk_paths, int n_shards, int64_t *bounds);

/* --- several devices behind one call (the only caller, main() at :588-601, stays unchanged) ------- */

/* A group drives several B200s of one box from ONE process: the chunk plan is cut into contiguous runs
 * of (nearly) equal path counts, one host thread per device evaluates its run on that device's handle
 * (Philox draws are keyed by (simulation, shuffle, trade), so a device regenerates whatever it needs),
 * and the ~1 KB chunk records are folded on the host in tree order: the result is bit-identical to
 * mcts_simulate on one device for any number of devices.  There is no device-to-device exchange at all
 * (paths are independent), so no collective is involved; the torch.distributed form of the same thing
 * (one process per GPU, one all_gather of the records) lives in mc_trade_sizing_b200/distributed.py. */
typedef struct mcts_group mcts_group;

/* devices == NULL or n <= 0: every visible compute-capability-10.x device.  Fails with MCTS_ENODEVICE
 * if there is none.  A device may be listed more than once (every entry gets its own handle). */
int mcts_group_create(const int *devices, int n, mcts_group **out);
void mcts_group_destroy(mcts_group *g);
int mcts_group_size(const mcts_group *g);
/* CUDA device index of member i (or -1). */
int mcts_group_device(const mcts_group *g, int i);
/* Same contract as mcts_simulate.  opt->chunk_paths == 0 means mcts_auto_chunk(P, group size). */
int mcts_group_simulate(mcts_group *g, const mcts_params *p, uint64_t seed, const mcts_options *opt,
                        mcts_agg *agg, double *profit, double *dd);
/* Per-device timing of the last mcts_group_simulate: out[mcts_group_size(g)]. */
int mcts_group_get_timing(const mcts_group *g, mcts_timing *out);

/* --- equity-curve export (extension; the reference only has the temporary of :167) -------- */

/* Equity curves of paths [first_path, first_path+n) for all 20 strategies written to DEVICE memory
 * `curves_dev` as [20][ceil(L/2)][n][2] doubles (two trades per 128-bit store, paths contiguous).
 * Returns the achieved device bandwidth through mcts_get_timing (paths_ms). */
int mcts_export_curves(mcts_handle *h, const mcts_params *p, uint64_t seed, int64_t first_path, int64_t n,
                       double *curves_dev);
void *mcts_device_alloc(mcts_handle *h, int64_t bytes);
void mcts_device_free(mcts_handle *h, void *ptr);
int mcts_copy_to_host(mcts_handle *h, void *dst_host, const void *src_dev, int64_t bytes);

/* --- result-table arithmetic of :513-521 (ratios, inf rules), strategy-id order ----------- */
/* rows[20][9] = avg_profit, avg_drawdown, ratio, min_profit, max_profit, min_dd, max_dd,
 *               avg_per_trade, ratio_max_dd */
int mcts_summary_rows(const mcts_agg *agg, int32_t num_trades, double *rows);

/* --- measurement ---------------------------------------------------------------------------- */
int mcts_get_timing(const mcts_handle *h, mcts_timing *out);
/* The cudaStream_t every kernel of this handle is launched on (for external CUDA-event timing). */
void *mcts_get_stream(const mcts_handle *h);
/* Issue-rate microbenchmarks used as roofline denominators: independent DADD chains
 * and independent IADD3/LOP3 chains over the whole chip. Results in warp-wide... thread-level
 * instructions per second. */
int mcts_measure_peaks(mcts_handle *h, double *fp64_inst_per_s, double *int_inst_per_s, double *sm_clock_mhz);
/* Rate (strategy-steps/s) of a register-only replica of the evaluator's trade step at the hot kernel's
 * occupancy: the ceiling of this instruction mix on the chip (no generator, no memory). */
int mcts_measure_step_ceiling(mcts_handle *h, double *strategy_steps_per_s);

/* Rate (strategy-steps/s) and instructions per cycle and scheduler of a register-only loop with the instruction MIX of the
 * shipped evaluator's trade step (5 fp64 adds + 4 integer multiply-adds + 4 logic operations per step, independent chains)
 * at the hot kernel's occupancy: what the schedulers can issue for this mix -- the practical ceiling of the issue roofline. */
int mcts_measure_mix_ceiling(mcts_handle *h, double *strategy_steps_per_s, double *ipc);

#ifdef __cplusplus
}
#endif
#endif /* MCTS_H */
