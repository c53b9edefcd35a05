"""mc_trade_sizing_b200 -- B200-native Monte Carlo path-simulation core for mc-trade-sizing.

    from mc_trade_sizing_b200 import run_all_strategies      # drop-in for the reference function
    python -m mc_trade_sizing_b200.trading_models --hit_rate ...   # drop-in for the reference CLI

All compute lives in libmcts.so (CUDA, sm_100a; C ABI in include/mcts.h)."""
from .engine import Engine, make_params  # noqa: F401
from .trading_models import DESCRIPTIONS, find_break_even_hit_rate, run_all_strategies  # noqa: F401
