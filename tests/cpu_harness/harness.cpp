// harness.cpp -- TEST ONLY.  Compiles the product's host/device headers (eval_core.cuh, gen_core.cuh,
// pairwise.cuh, planner.hpp) with g++ so that the rule tables, the pairwise leaf schedule, the
// urn generator and the chunk planner can be unit-tested on a machine without a GPU.  This is a
// test of the device code's logic, not a CPU fallback: nothing in the product links or loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mcts.h"
#include "../../mc_trade_sizing_b200/csrc/eval_core.cuh"
#include "../../mc_trade_sizing_b200/csrc/gen_core.cuh"
#include "../../mc_trade_sizing_b200/csrc/pairwise.cuh"
#include "../../mc_trade_sizing_b200/csrc/planner.hpp"

using namespace mcts;

template <class IDS, bool DEDUP, class SRC>
static void run_pass(SRC &src, const LeafSched *sch, double *profit, double *dd, int64_t P, int64_t path)
{
    double p[IDS::N], d[IDS::N];
    NoSink sink;
    eval_path<IDS>(src, sink, sch, p, d);
    for (int s = 0; s < IDS::N; s++) {
        const int id = IDS::id(s);
        for (int other = 1; other <= 20; other++) {
            if (other == id || (DEDUP && rule_of_id(other) == rule_of_id(id))) {
                profit[(int64_t)(other - 1) * P + path] = p[s];
                dd[(int64_t)(other - 1) * P + path] = d[s];
            }
        }
    }
}

template <class SRC>
static void run_all(SRC &src, const LeafSched *sch, int dedup, double *profit, double *dd, int64_t P, int64_t path)
{
    if (dedup) {
        run_pass<DedupPass0, true>(src, sch, profit, dd, P, path);
        run_pass<DedupPass1, true>(src, sch, profit, dd, P, path);
        run_pass<DedupPass2, true>(src, sch, profit, dd, P, path);
    } else {
        run_pass<FullPass0, false>(src, sch, profit, dd, P, path);
        run_pass<FullPass1, false>(src, sch, profit, dd, P, path);
        run_pass<FullPass2, false>(src, sch, profit, dd, P, path);
        run_pass<FullPass3, false>(src, sch, profit, dd, P, path);
    }
}

template <class IDS, bool DEDUP>
static void run_pass_int(const uint32_t *codes, const IntTab *tab, int L, double *profit, double *dd, int64_t P, int64_t path)
{
    int32_t p[IDS::N], d[IDS::N];
    eval_path_int<IDS>(codes, 1, tab, L, p, d);
    for (int s = 0; s < IDS::N; s++) {
        const int id = IDS::id(s);
        for (int other = 1; other <= 20; other++) {
            if (other == id || (DEDUP && rule_of_id(other) == rule_of_id(id))) {
                profit[(int64_t)(other - 1) * P + path] = (double)p[s];
                dd[(int64_t)(other - 1) * P + path] = (double)d[s];
            }
        }
    }
}

extern "C" {

// replay evaluator (ValueSource) on values[P][T]
int h_eval_values(const double *values, int64_t P, int64_t T, int dedup, double *profit, double *dd)
{
    LeafSched sch;
    std::string err;
    if (build_leaf_sched(T, &sch, &err)) return -1;
    for (int64_t path = 0; path < P; path++) {
        ValueSource src;
        src.row = values + path * T;
        run_all(src, &sch, dedup, profit, dd, P, path);
    }
    return 0;
}

// full simulate path on the host: spec -> base counts -> urn codes -> CodeSource evaluator.
// seqs (optional) receives the decoded sequences [P][L].
// mode bit 0 = dedup, bit 1 = use the exact int32 evaluator when the run is all-integer
int h_simulate(const mcts_params *p, uint64_t seed, int mode, double *profit, double *dd, double *seqs, int64_t first_path, int64_t n)
{
    const int dedup = mode & 1, want_int = mode & 2;
    GenSpec g;
    std::string err;
    int rc = build_spec(*p, &g, &err);
    if (rc) return rc;
    LeafSched sch;
    if ((rc = build_leaf_sched(g.L, &sch, &err))) return rc;
    const int n_words = (g.L + 7) / 8;
    std::vector<uint32_t> codes(n_words);
    for (int64_t i = 0; i < n; i++) {
        const int64_t path = first_path + i;
        const int64_t sim = path / p->num_mc_shuffles;
        const uint32_t k = (uint32_t)(path % p->num_mc_shuffles);
        uint32_t cnt[kMaxCat];
        gen_base_counts(g, seed, (uint32_t)sim, cnt);
        if (g.wide) {  // fp64 shuffle + replay evaluator, as launch_wide does on the device
            std::vector<double> vals(g.L);
            gen_path_values_fy(g, seed, (uint32_t)sim, k, cnt, vals.data(), 1);
            if (seqs) memcpy(seqs + i * g.L, vals.data(), sizeof(double) * g.L);
            if (profit) {
                ValueSource src;
                src.row = vals.data();
                run_all(src, &sch, dedup, profit, dd, n, i);
            }
            continue;
        }
        switch (urn_width(g.K)) {  // the dispatch of paths_launch.cuh
            case 2: gen_path_codes_urn<2>(g, seed, (uint32_t)sim, k, cnt, codes.data(), 1); break;
            case 5: gen_path_codes_urn<5>(g, seed, (uint32_t)sim, k, cnt, codes.data(), 1); break;
            case 6: gen_path_codes_urn<6>(g, seed, (uint32_t)sim, k, cnt, codes.data(), 1); break;
            default: gen_path_codes_fy(g, seed, (uint32_t)sim, k, cnt, codes.data(), 1); break;
        }
        if (seqs)
            for (int t = 0; t < g.L; t++) seqs[i * g.L + t] = g.tab[(codes[t >> 3] >> (4 * (t & 7))) & 15].v1;
        if (profit && want_int && g.all_int) {
            IntTab tab[16];
            for (int c = 0; c < 16; c++) tab[c] = IntTab{g.ival[c], 2 * g.ival[c]};
            if (dedup) run_pass_int<DedupAll, true>(codes.data(), tab, g.L, profit, dd, n, i);
            else run_pass_int<FullAll, false>(codes.data(), tab, g.L, profit, dd, n, i);
        } else if (profit) {
            CodeSource src;
            src.codes = codes.data();
            src.stride = 1;
            src.tab = g.tab;
            src.word = 0;
            run_all(src, &sch, dedup, profit, dd, n, i);
        }
    }
    return 0;
}

int h_base_counts(const mcts_params *p, uint64_t seed, int64_t sim, uint32_t *counts, double *vals, int *K, int *L)
{
    GenSpec g;
    std::string err;
    int rc = build_spec(*p, &g, &err);
    if (rc) return rc;
    uint32_t c[kMaxCat];
    gen_base_counts(g, seed, (uint32_t)sim, c);
    for (int k = 0; k < kMaxCat; k++) {
        counts[k] = c[k];
        vals[k] = g.val[k];
    }
    *K = g.K;
    *L = g.L;
    return 0;
}

void h_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out)
{
    uint32_t o[4];
    philox4x32_10(k0, k1, c0, c1, c2, c3, o);
    memcpy(out, o, 16);
}

double h_pairwise(const double *a, int64_t n) { return 0.0 + np_pairwise_raw(a, n); }

// emulate the device reduction + host combine of one array of P per-path values
double h_chunked_sum(const double *a, int64_t P, int64_t chunk_paths)
{
    const int64_t chunk = clamp_chunk(chunk_paths);
    std::vector<int64_t> first, count;
    plan_visit(0, P, chunk, &first, &count);
    std::vector<double> partial(first.size());
    for (size_t c = 0; c < first.size(); c++) {
        const int64_t n = count[c];
        const int d = reduce_depth(n);
        std::vector<double> part((size_t)1 << d);
        for (int64_t j = 0; j < ((int64_t)1 << d); j++) {
            int64_t s, len;
            np_node_at_depth(n, d, j, &s, &len);
            part[j] = np_pairwise_raw(a + first[c] + s, len);
        }
        for (int64_t w = (int64_t)1 << d; w > 1; w >>= 1)
            for (int64_t i = 0; i < (w >> 1); i++) part[i] = part[2 * i] + part[2 * i + 1];
        partial[c] = part[0];
    }
    int64_t idx = 0;
    return 0.0 + combine_visit(P, chunk, partial.data(), &idx);
}

int h_build_spec(const mcts_params *p, GenSpec *g)
{
    std::string err;
    return build_spec(*p, g, &err);
}

}  // extern "C"
