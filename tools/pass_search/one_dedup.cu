// The dedup flavour (11 distinct behaviours) of the hot kernel, alone.  See search_orders.py.
#include "kernels.cuh"
void launch_one(const mcts::PathArgs &a) { mcts::k_sim_paths<16, mcts::PM_DEDUP, 128, false><<<1, 128>>>(a); }
