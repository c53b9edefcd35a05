#include "paths_launch.cuh"
MCTS_DEFINE_INT_LAUNCHER(mcts_launch_paths_int_full, mcts::PM_FULL)
MCTS_DEFINE_INT_LAUNCHER(mcts_launch_paths_int_dedup, mcts::PM_DEDUP)
