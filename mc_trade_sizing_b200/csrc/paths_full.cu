// All 20 state machines in fp64.  Peak and drawdown are updated with predicated integer multiply-adds
// on the FMA pipe instead of select pairs on the (saturated) ALU pipe: +4.4 % on the B200, same bits.
#define MCTS_PSEL 3
#include "paths_launch.cuh"
MCTS_DEFINE_PATH_LAUNCHER(mcts_launch_paths_full, mcts::PM_FULL)
