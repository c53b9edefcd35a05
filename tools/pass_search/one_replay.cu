// The replay kernel (all 20 ids, caller-supplied values) alone.  See search_insitu.py --kernel replay.
#include "kernels.cuh"
void launch_one(const mcts::PathArgs &a, const double *v) { mcts::k_replay_paths<mcts::PM_FULL><<<1, 128, mcts::kReplaySmem>>>(a, v, 1000); }
