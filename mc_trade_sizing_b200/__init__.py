"""mc_trade_sizing_b200 -- B200-native Monte Carlo path-simulation core for mc-trade-sizing.

    from mc_trade_sizing_b200 import run_all_strategies      # drop-in for the reference function
    python -m mc_trade_sizing_b200.trading_models --hit_rate ...   # drop-in for the reference CLI

All compute lives in libmcts.so (CUDA, sm_100a; C ABI in include/mcts.h).  Names are resolved on first use so
that importing the package (the CLI does) costs nothing but the interpreter's start-up."""
_LAZY = {"Engine": "engine", "EngineGroup": "engine", "make_params": "engine", "DESCRIPTIONS": "trading_models",
         "find_break_even_hit_rate": "trading_models", "run_all_strategies": "trading_models"}
__all__ = sorted(_LAZY)


def __getattr__(name):
    if name in _LAZY:
        import importlib
        value = getattr(importlib.import_module("." + _LAZY[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
