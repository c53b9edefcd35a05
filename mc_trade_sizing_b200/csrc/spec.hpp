// spec.hpp -- plain-data descriptions shared by the host planner and the kernels.
//
// A run of run_all_strategies (reference: src/trading_models.py:445-530) is lowered on the host
// to a GenSpec: the generator (:43-128) becomes either a list of iid Bernoulli segments or a
// Markov chain over integer thresholds, and the <= 16 distinct trade values become "categories"
// whose 4-bit codes (bit 0 = "value > 0", the win test of :154) are what the path kernel keeps
// per trade in shared memory.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MCTS_HD __host__ __device__ __forceinline__
#else
#define MCTS_HD inline
#endif

namespace mcts {

constexpr int kNumStrategies = 20;
constexpr int kMaxCat = 32;       // distinct trade values per run (16 regimes x {win, loss})
constexpr int kMaxCodes = 16;     // of which the 4-bit code path handles 16 (<= 8 positive and <= 8 non-positive)
constexpr int kMaxSeg = 20;       // 16 regimes, or the 3 phases + tail fill of :44-65
constexpr int kMaxStackDepth = 12; // pairwise-tree depth over trades: T <= 128 * 2^11
constexpr int kMaxLeaves = 4096;  // <= ~260,000 trades per path

struct Segment {  // one np.random.choice([win, -loss], size=len, p=[p, 1-p]) block (:55-59, :119-123)
    int32_t len;
    int32_t cat_win, cat_loss;
    int32_t pad;
    uint64_t thr;  // win <=> 53-bit uniform integer < thr   (u < cdf[0], searchsorted side='right')
};

struct Tab {  // contribution of one trade at size 1, 2, 3: fl(v*1), fl(v*2), fl(v*3)  (:147)
    double v1, v2, v3, pad;
};

enum GenKind : int32_t { GEN_SEGMENTS = 0, GEN_MARKOV1 = 1, GEN_MARKOV2 = 2 };

struct GenSpec {
    int32_t kind;
    int32_t L;       // trades per path
    int32_t K;       // categories in use, 1..kMaxCat
    int32_t n_seg;
    Segment seg[kMaxSeg];
    uint64_t thr_init;   // P(win) of the first one/two trades (:70, :89)
    uint64_t thr1[2];    // Markov-1: [after loss, after win] (:75-79)
    uint64_t thr2[4];    // Markov-2: index = older*2 + newer (:93-100)
    int32_t cat_win, cat_loss;  // Markov modes
    uint32_t rank_codes;        // nibble r = 4-bit code of category r (r < 8)
    uint32_t rank_codes_hi;     // categories 8..15
    int32_t wide;               // more values than 4-bit codes can name: shuffle fp64 values in global memory instead
    int32_t all_int;            // every value integral and 6*max|v|*L < 2^31
    double val[kMaxCat];
    Tab tab[16];                // indexed by code
    int32_t ival[16];           // indexed by code (valid when all_int)
};

struct LeafSched {  // NumPy pairwise-summation tree over the L trades of one path
    int32_t n_leaves;
    int32_t n_oct;       // L / 8: full 8-trade blocks
    int32_t tail;        // L % 8: trades added one by one after the last leaf's blocks
    int32_t max_depth;
    int32_t end_oct[kMaxLeaves];  // exclusive end (in 8-trade blocks) of leaf i
    int32_t merges[kMaxLeaves];   // tree nodes completed right after leaf i
};

}  // namespace mcts
