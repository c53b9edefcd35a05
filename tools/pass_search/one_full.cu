// One instantiation of the hot kernel (all 20 strategies, Fisher-Yates generator, 128-thread blocks) so that a candidate
// pass composition compiles in ~4 s.  See search_orders.py.
#include "kernels.cuh"
void launch_one(const mcts::PathArgs &a) { mcts::k_sim_paths<16, mcts::PM_FULL, 128, false><<<1, 128>>>(a); }
