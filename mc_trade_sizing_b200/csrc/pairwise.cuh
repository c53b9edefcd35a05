// pairwise.cuh -- NumPy's pairwise summation (DOUBLE_pairwise_sum for contiguous input)
// as an iterative host/device routine, used for the exact-order mean over paths
// (np.mean at src/trading_models.py:513-514).
#pragma once
#include <stdint.h>

#include "spec.hpp"

namespace mcts {

MCTS_HD int64_t np_split_hd(int64_t n)
{
    int64_t n2 = n / 2;
    return n2 - n2 % 8;
}

// leaf of the tree: n <= 128
MCTS_HD double np_leaf_sum(const double *__restrict__ a, int64_t n)
{
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res = res + a[i];
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res = res + a[i];
    return res;
}

// np pairwise sum of a[0..n) WITHOUT the leading `0.0 +` of np.add.reduce; also min / max.
MCTS_HD double np_pairwise_raw(const double *__restrict__ a, int64_t n)
{
    // explicit post-order walk; depth <= 64 - log2(128)
    int64_t nn[60];
    double left[60];
    int8_t st[60];
    int sp = 0;
    int64_t pos = 0;
    nn[0] = n;
    st[0] = 0;
    sp = 1;
    double v = 0.0;
    while (true) {
        const int top = sp - 1;
        if (st[top] == 0) {
            if (nn[top] <= 128) {
                v = np_leaf_sum(a + pos, nn[top]);
                pos += nn[top];
                sp--;
            } else {
                st[top] = 1;
                nn[sp] = np_split_hd(nn[top]);
                st[sp] = 0;
                sp++;
                continue;
            }
        }
        // deliver v to the parents
        while (true) {
            if (sp == 0) return v;
            const int par = sp - 1;
            if (st[par] == 1) {
                left[par] = v;
                st[par] = 2;
                nn[sp] = nn[par] - np_split_hd(nn[par]);
                st[sp] = 0;
                sp++;
                break;
            }
            v = left[par] + v;
            sp--;
        }
    }
}

// Node j (0 <= j < 2^d) at depth d of the tree over n elements: the top d levels are a perfect
// binary tree as long as every node above depth d has more than 128 elements.
MCTS_HD void np_node_at_depth(int64_t n, int d, int64_t j, int64_t *start, int64_t *len)
{
    int64_t s = 0;
    for (int lvl = d - 1; lvl >= 0; lvl--) {
        const int64_t n2 = np_split_hd(n);
        if ((j >> lvl) & 1) {
            s += n2;
            n -= n2;
        } else {
            n = n2;
        }
    }
    *start = s;
    *len = n;
}

}  // namespace mcts
