"""Drop-in for the hot path of the reference's src/trading_models.py.

Same public names, argument meaning, return value and stdout as the reference:

    run_all_strategies(...)        reference src/trading_models.py:445-530
    find_break_even_hit_rate(...)  reference src/trading_models.py:408-442
    main()                         reference src/trading_models.py:532-671 (argparse flags and table)

Everything between the arguments and the 20-row result table runs on the B200 through libmcts.so
(include/mcts.h).  There is no CPU fallback: without the library or a GPU the call raises.

Extra knobs are environment variables so that the unchanged dps.py keeps working:
    MCTS_SEED    Philox key (default 12345; the reference never seeds NumPy, so any seed is a valid run)
    MCTS_DEVICE  CUDA device index (default 0)
    MCTS_DEDUP   1 = evaluate the 11 distinct behaviours once and fan out (identical results)
    MCTS_CHUNK   max paths per device chunk
"""
from __future__ import annotations

import argparse
import json as pyjson
import os

from . import engine as _engine

# strategy descriptions, reference src/trading_models.py:451-472 (they are the `Strategy` column of
# every downstream sink: dps.py:362)
DESCRIPTIONS = {
    1: "Constant position size 1",
    2: "Increase to 2 after win, reset to 1 after loss",
    3: "Increase to 2 after win, reset to 1 after loss or 2 wins",
    4: "Increase to 2 after win, reset to 1 after loss or 3 wins",
    5: "Increase to 2 after win, reset to 1 after loss or 4 wins",
    6: "Increase to 2 after loss, reset to 1 after win",
    7: "Increase to 2 after 2 losses, reset to 1 after win",
    8: "Increase to 2 after 3 losses, reset to 1 after win",
    9: "Pause after 1 win until next loss",
    10: "Pause after 2 wins until next loss",
    11: "Pause after 3 wins until next loss",
    12: "Pause after 4 wins until next loss",
    13: "Increase to 2 after 2 wins, reset to 1 after 2 losses",
    14: "Increase to 2 after 1 win and 1 loss, else 1",
    15: "Pause after 2 consecutive wins until 1 loss, then increase to 2",
    16: "Increase to 2 after 2 losses, pause after 1 win until next loss",
    17: "Increase to 2 after 1 win only if preceded by 1 loss, else 1",
    18: "Increase to 3 after 3 wins, reset to 1 after 1 loss",
    19: "Increase to 2 after 2 wins, to 3 after 2 losses, else 1",
    20: "Increase to 2 after 1 win, to 3 after 2 losses, reset to 1 after win",
}

_engines: dict[int, "_engine.Engine"] = {}


def get_engine(device: int | None = None) -> "_engine.Engine":
    dev = int(os.environ.get("MCTS_DEVICE", "0")) if device is None else int(device)
    if dev not in _engines:
        _engines[dev] = _engine.Engine(dev)
    return _engines[dev]


def rank_rows(rows):
    """The ranking of reference :524-528: rows with Profit/MaxDD >= 0 first, each group sorted
    descending by it; Python's sort is stable so ties keep strategy-id order."""
    positive = [r for r in rows if r[9] >= 0]
    negative = [r for r in rows if r[9] < 0]
    positive.sort(key=lambda x: x[9], reverse=True)
    negative.sort(key=lambda x: x[9], reverse=True)
    return positive + negative


def summary_to_rows(summary):
    """[20,9] array in strategy-id order -> the reference's list of 10-tuples (:522), ranked."""
    import numpy as np
    rows = [(DESCRIPTIONS[i + 1],) + tuple(np.float64(x) for x in summary[i]) for i in range(20)]
    return rank_rows(rows)


def run_all_strategies(
    hit_rate, avg_win, avg_loss, num_trades, num_simulations, num_mc_shuffles,
    use_markov=False, p_win_after_win=0.7, p_win_after_loss=0.5,
    use_markov2=False, p_win_ww=0.8, p_win_wl=0.6, p_win_lw=0.5, p_win_ll=0.3,
    use_regime=False, regimes=None, *, seed=None, device=None, dedup=None, chunk_paths=None
):
    """Same signature and return value as the reference (src/trading_models.py:445-450, :522-530):
    a ranked list of 20 tuples (description, avg_profit, avg_drawdown, ratio, min_profit, max_profit,
    min_dd, max_dd, avg_per_trade, ratio_max_dd).  The keyword-only extras default from MCTS_* env."""
    seed = int(os.environ.get("MCTS_SEED", str(_engine.DEFAULT_SEED))) if seed is None else int(seed)
    dedup = os.environ.get("MCTS_DEDUP", "0") == "1" if dedup is None else bool(dedup)
    chunk_paths = int(os.environ.get("MCTS_CHUNK", "0")) if chunk_paths is None else int(chunk_paths)
    params = _engine.make_params(
        hit_rate, avg_win, avg_loss, num_trades, num_simulations, num_mc_shuffles,
        use_markov=use_markov, p_win_after_win=p_win_after_win, p_win_after_loss=p_win_after_loss,
        use_markov2=use_markov2, p_win_ww=p_win_ww, p_win_wl=p_win_wl, p_win_lw=p_win_lw, p_win_ll=p_win_ll,
        use_regime=use_regime, regimes=regimes)
    eng = get_engine(device)
    out = eng.simulate(params, seed=seed, options=eng.options(chunk_paths=chunk_paths, dedup=dedup))
    return summary_to_rows(out["summary"])


def find_break_even_hit_rate(avg_win, avg_loss, mode):
    """Reference src/trading_models.py:408-442 (printed by main(), scraped by dps.py:154)."""
    if avg_win is None or avg_loss is None or avg_win <= 0 or avg_loss <= 0:
        print("❌ ERROR: Invalid values for avg_win or avg_loss! Calculation not possible.")
        return None
    base_rate = avg_loss / (avg_win + avg_loss)
    if base_rate == float("inf"):
        print("❌ ERROR: Calculation resulted in an infinite value!")
        return None
    if mode == "1st Order Markov":
        base_rate *= (1 + (0.7 - 0.5) / 5)
    elif mode == "2nd Order Markov":
        base_rate *= (1 + (0.8 - 0.5) / 4)
    elif mode == "Regime Switching":
        base_rate *= (1 + (0.6 - 0.5) / 3)
    return base_rate


def build_parser():
    """The 16 flags of reference :533-550, byte-compatible (dps.py:217-258 builds these command lines)."""
    parser = argparse.ArgumentParser(description="Simulate 20 trading strategies with/without Markov correlations, second-order Markov, and regime switching")
    parser.add_argument("--hit_rate", type=float, required=True, help="Hit rate, e.g. 0.7")
    parser.add_argument("--avg_win", type=float, required=True, help="Average win per trade")
    parser.add_argument("--avg_loss", type=float, required=True, help="Average loss per trade")
    parser.add_argument("--num_simulations", type=int, default=200, help="Number of simulations (default: 200)")
    parser.add_argument("--num_trades", type=int, default=400, help="Number of trades per simulation (default: 400)")
    parser.add_argument("--num_mc_shuffles", type=int, default=200, help="Number of shuffles per simulation (default: 200)")
    parser.add_argument("--use_markov", action="store_true", help="Use Markov chain correlations (1st order)")
    parser.add_argument("--p_win_after_win", type=float, default=0.7, help="P(win|win) for 1st order Markov model")
    parser.add_argument("--p_win_after_loss", type=float, default=0.5, help="P(win|loss) for 1st order Markov model")
    parser.add_argument("--use_markov2", action="store_true", help="Use 2nd order Markov model")
    parser.add_argument("--p_win_ww", type=float, default=0.8, help="P(win|win,win) for 2nd order Markov")
    parser.add_argument("--p_win_wl", type=float, default=0.6, help="P(win|win,loss) for 2nd order Markov")
    parser.add_argument("--p_win_lw", type=float, default=0.5, help="P(win|loss,win) for 2nd order Markov")
    parser.add_argument("--p_win_ll", type=float, default=0.3, help="P(win|loss,loss) for 2nd order Markov")
    parser.add_argument("--use_regime", action="store_true", help="Use regime switching model")
    parser.add_argument("--regimes", type=str, default=None, help="Regime list as JSON string")
    return parser


def _model_label(args):
    pct = int(round(args.hit_rate * 100))
    if args.use_regime:
        return f"Hit rate: {pct}%  -  Regime Switching Model"
    if args.use_markov2:
        return f"Hit rate: {pct}%  -  2nd Order Markov"
    if args.use_markov:
        return f"Hit rate: {pct}%  -  1st Order Markov"
    return f"Hit rate: {pct}%  -  No Markov"


def format_row(row):
    """One table row, reference :639-643 (matched by the scraper regex dps.py:356)."""
    description, profit, dd, ratio, min_p, max_p, min_d, max_d, avg_per_trade, ratio_max_dd = row
    return (
        f"{description:<90} {profit:14.2f} {dd:16.2f} {ratio:12.2f} "
        f"{min_p:12.2f} {max_p:12.2f} {min_d:14.2f} {max_d:14.2f} "
        f"{avg_per_trade:12.2f} {ratio_max_dd:18.2f}"
    )


def main(argv=None):
    """stdout-compatible with reference :552-671 (settings block, break-even line, 20-row table,
    rule after the third row, optional colorama 'Top 4' block)."""
    args = build_parser().parse_args(argv)

    print("\n" + "=" * 90)
    print("CURRENT SIMULATION SETTING:")
    print(f"Hit rate: {args.hit_rate:.2%}")
    if args.use_regime:
        print("Mode: Regime Switching")
        if args.regimes:
            print(f"Regimes: {args.regimes}")
    elif args.use_markov2:
        print("Mode: 2nd Order Markov")
        print(f"P(win|WW): {args.p_win_ww}, P(win|WL): {args.p_win_wl}, P(win|LW): {args.p_win_lw}, P(win|LL): {args.p_win_ll}")
    elif args.use_markov:
        print("Mode: 1st Order Markov")
        print(f"P(win|win): {args.p_win_after_win}, P(win|loss): {args.p_win_after_loss}")
    else:
        print("Mode: No Markov")
    print("=" * 90 + "\n")

    print(f"Average win per trade: €{args.avg_win}")
    print(f"Average loss per trade: €{args.avg_loss}")
    print(f"Number of simulations: {args.num_simulations}")
    print(f"Number of trades per simulation: {args.num_trades}")
    print(f"Number of shuffles per simulation: {args.num_mc_shuffles}")

    mode = "No Markov"
    if args.use_regime:
        mode = "Regime Switching"
    elif args.use_markov2:
        mode = "2nd Order Markov"
    elif args.use_markov:
        mode = "1st Order Markov"

    breakeven = find_break_even_hit_rate(args.avg_win, args.avg_loss, mode)
    print(f"Break-even hit rate: {breakeven:.2%}")

    regimes = pyjson.loads(args.regimes) if args.use_regime and args.regimes else None

    summary = run_all_strategies(
        args.hit_rate, args.avg_win, args.avg_loss, args.num_trades,
        args.num_simulations, args.num_mc_shuffles,
        use_markov=args.use_markov, p_win_after_win=args.p_win_after_win, p_win_after_loss=args.p_win_after_loss,
        use_markov2=args.use_markov2, p_win_ww=args.p_win_ww, p_win_wl=args.p_win_wl,
        p_win_lw=args.p_win_lw, p_win_ll=args.p_win_ll,
        use_regime=args.use_regime, regimes=regimes,
    )

    print("\nResults (Monte Carlo, based on input parameters):\n")
    header = (
        f"{'Strategy':<90} {'Avg Profit (€)':>14} {'Avg Drawdown (€)':>16} {'Ratio':>12} "
        f"{'Min (€)':>12} {'Max (€)':>12} {'Min DD (€)':>14} {'Max DD (€)':>14} "
        f"{'Avg/Trade':>12} {'Profit/MaxDD':>18}"
    )
    print(header)
    print("=" * len(header))
    print()
    try:
        from colorama import Fore, Style
        print(Fore.YELLOW + f"*** {_model_label(args)} ***" + Style.RESET_ALL)
    except ImportError:
        print(f"*** {_model_label(args)} ***")
    print()

    for idx, row in enumerate(summary):
        print(format_row(row))
        if idx == 2:
            print("-" * len(header))

    try:
        from colorama import Fore, Style, init
        init(autoreset=True)
        konst_idx = next((i for i, row in enumerate(summary) if row[0].startswith("Constant position size 1")), None)
        print("\n\n\nTop 4 strategies compared to 'Constant position size 1':")
        print("--------------------------------------------------------------")
        for idx in range(4):
            if idx >= len(summary):
                break
            print()
            print(Fore.GREEN + format_row(summary[idx]) + Style.RESET_ALL)
            if konst_idx is not None:
                print(Fore.RED + format_row(summary[konst_idx]) + Style.RESET_ALL)
    except ImportError:
        pass


if __name__ == "__main__":
    main()
