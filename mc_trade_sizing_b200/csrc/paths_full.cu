#include "paths_launch.cuh"
MCTS_DEFINE_PATH_LAUNCHER(mcts_launch_paths_full, mcts::PM_FULL)
