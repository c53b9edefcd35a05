"""Drop-in for the hot path of the reference's src/trading_models.py.

Same public names, argument meaning, return value and stdout as the reference:

    run_all_strategies(...)        reference src/trading_models.py:445-530
    find_break_even_hit_rate(...)  reference src/trading_models.py:408-442
    main()                         reference src/trading_models.py:532-671 (argparse flags and table)

Everything between the arguments and the 20-row result table runs on the B200 through libmcts.so
(include/mcts.h).  There is no CPU fallback: without the library or a GPU the call raises.

Extra knobs are environment variables so that the unchanged dps.py keeps working:
    MCTS_SEED     Philox key.  Unset: a fresh 64-bit key from os.urandom per call -- like the reference, which never
                  seeds NumPy, two runs differ; the key used is kept in `last_seed` and, with MCTS_VERBOSE=1,
                  written to stderr so that a run can be repeated
    MCTS_DEVICE   CUDA device index (default 0)
    MCTS_DEVICES  "all" or a comma-separated list: shard every call over several B200s inside libmcts
                  (mcts_group_*; bit-identical results, no torchrun needed)
    MCTS_DEDUP    1 = evaluate the 11 distinct behaviours once and fan out (identical results)
    MCTS_CHUNK    max paths per device chunk
    MCTS_SERVER   1 = hand the call to the per-user daemon (mc_trade_sizing_b200/server.py) that keeps the CUDA
                  context and the engines alive between CLI invocations; started on first use

This module imports nothing heavy at import time (no NumPy, no ctypes): the CLI's start-up cost is the
interpreter's, which is what makes a daemon-served invocation take tens of milliseconds.
"""
from __future__ import annotations

import argparse
import json as pyjson
import os
import sys

last_seed = None  # Philox key of the most recent run_all_strategies call in this process

# strategy descriptions, reference src/trading_models.py:451-472 (they are the `Strategy` column of
# every downstream sink: dps.py:362)
_DESCRIPTION_TEXT = (
    "Constant position size 1",
    "Increase to 2 after win, reset to 1 after loss",
    "Increase to 2 after win, reset to 1 after loss or 2 wins",
    "Increase to 2 after win, reset to 1 after loss or 3 wins",
    "Increase to 2 after win, reset to 1 after loss or 4 wins",
    "Increase to 2 after loss, reset to 1 after win",
    "Increase to 2 after 2 losses, reset to 1 after win",
    "Increase to 2 after 3 losses, reset to 1 after win",
    "Pause after 1 win until next loss",
    "Pause after 2 wins until next loss",
    "Pause after 3 wins until next loss",
    "Pause after 4 wins until next loss",
    "Increase to 2 after 2 wins, reset to 1 after 2 losses",
    "Increase to 2 after 1 win and 1 loss, else 1",
    "Pause after 2 consecutive wins until 1 loss, then increase to 2",
    "Increase to 2 after 2 losses, pause after 1 win until next loss",
    "Increase to 2 after 1 win only if preceded by 1 loss, else 1",
    "Increase to 3 after 3 wins, reset to 1 after 1 loss",
    "Increase to 2 after 2 wins, to 3 after 2 losses, else 1",
    "Increase to 2 after 1 win, to 3 after 2 losses, reset to 1 after win",
)
DESCRIPTIONS = {i + 1: text for i, text in enumerate(_DESCRIPTION_TEXT)}

_engines: dict = {}


def _device_spec(device):
    """None -> MCTS_DEVICES / MCTS_DEVICE; returns an int (one device) or a tuple of ints / 'all' (a group)."""
    if device is None:
        many = os.environ.get("MCTS_DEVICES", "").strip()
        if many:
            device = many
        else:
            return int(os.environ.get("MCTS_DEVICE", "0"))
    if isinstance(device, str):
        if device.lower() == "all":
            return "all"
        ids = tuple(int(x) for x in device.split(",") if x.strip() != "")
        return ids[0] if len(ids) == 1 else ids
    if isinstance(device, (list, tuple)):
        return int(device[0]) if len(device) == 1 else tuple(int(d) for d in device)
    return int(device)


def get_engine(device=None):
    """The process-wide engine for a device (Engine) or a device list / 'all' (EngineGroup); created on first use."""
    from . import engine as _engine
    key = _device_spec(device)
    if key not in _engines:
        if key == "all":
            _engines[key] = _engine.EngineGroup(None)
        elif isinstance(key, tuple):
            _engines[key] = _engine.EngineGroup(list(key))
        else:
            _engines[key] = _engine.Engine(key)
    return _engines[key]


def resolve_seed(seed=None) -> int:
    """Explicit seed > MCTS_SEED > a fresh key from os.urandom (the reference is unseeded: reruns must differ)."""
    global last_seed
    if seed is None:
        env = os.environ.get("MCTS_SEED")
        seed = int(env) if env not in (None, "") else int.from_bytes(os.urandom(8), "little")
    last_seed = int(seed) & (2 ** 64 - 1)
    if os.environ.get("MCTS_VERBOSE") == "1":
        print("[mcts] seed=%d" % last_seed, file=sys.stderr)
    return last_seed


def rank_rows(rows):
    """The ranking of reference :524-528: rows with Profit/MaxDD >= 0 first, each group sorted
    descending by it; Python's sort is stable so ties keep strategy-id order."""
    positive = [r for r in rows if r[9] >= 0]
    negative = [r for r in rows if r[9] < 0]
    positive.sort(key=lambda x: x[9], reverse=True)
    negative.sort(key=lambda x: x[9], reverse=True)
    return positive + negative


def summary_to_rows(summary):
    """[20,9] values in strategy-id order -> the reference's list of 10-tuples (:522), ranked."""
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
    seed = resolve_seed(seed)
    dedup = os.environ.get("MCTS_DEDUP", "0") == "1" if dedup is None else bool(dedup)
    chunk_paths = int(os.environ.get("MCTS_CHUNK", "0")) if chunk_paths is None else int(chunk_paths)
    if os.environ.get("MCTS_SERVER") == "1":
        from . import server
        return rank_rows(server.client_run(
            [hit_rate, avg_win, avg_loss, num_trades, num_simulations, num_mc_shuffles],
            dict(use_markov=bool(use_markov), p_win_after_win=p_win_after_win, p_win_after_loss=p_win_after_loss,
                 use_markov2=bool(use_markov2), p_win_ww=p_win_ww, p_win_wl=p_win_wl, p_win_lw=p_win_lw, p_win_ll=p_win_ll,
                 use_regime=bool(use_regime), regimes=regimes),
            seed=seed, device=_device_spec(device), dedup=dedup, chunk_paths=chunk_paths))
    from . import engine as _engine
    params = _engine.make_params(
        hit_rate, avg_win, avg_loss, num_trades, num_simulations, num_mc_shuffles,
        use_markov=use_markov, p_win_after_win=p_win_after_win, p_win_after_loss=p_win_after_loss,
        use_markov2=use_markov2, p_win_ww=p_win_ww, p_win_wl=p_win_wl, p_win_lw=p_win_lw, p_win_ll=p_win_ll,
        use_regime=use_regime, regimes=regimes)
    eng = get_engine(device)
    out = eng.simulate(params, seed=seed, options=eng.options(chunk_paths=chunk_paths, dedup=dedup))
    return summary_to_rows(out["summary"])


_BREAK_EVEN_FACTOR = {"1st Order Markov": 1 + (0.7 - 0.5) / 5, "2nd Order Markov": 1 + (0.8 - 0.5) / 4,
                      "Regime Switching": 1 + (0.6 - 0.5) / 3}


def find_break_even_hit_rate(avg_win, avg_loss, mode):
    """Reference src/trading_models.py:408-442 (printed by main(), scraped by dps.py:154): loss / (win + loss)
    times the reference's hard-coded per-mode factor; None (after the reference's message) for invalid input."""
    if avg_win is None or avg_loss is None or avg_win <= 0 or avg_loss <= 0:
        print("❌ ERROR: Invalid values for avg_win or avg_loss! Calculation not possible.")
        return None
    rate = avg_loss / (avg_win + avg_loss)
    if rate == float("inf"):
        print("❌ ERROR: Calculation resulted in an infinite value!")
        return None
    factor = _BREAK_EVEN_FACTOR.get(mode)
    return rate * factor if factor is not None else rate


# The 16 command-line flags of the reference CLI (reference :533-550; dps.py:217-258 builds these
# command lines): (name, type or None for a switch, default or REQUIRED, help text).
REQUIRED = object()
CLI_FLAGS = (
    ("hit_rate", float, REQUIRED, "Hit rate, e.g. 0.7"),
    ("avg_win", float, REQUIRED, "Average win per trade"),
    ("avg_loss", float, REQUIRED, "Average loss per trade"),
    ("num_simulations", int, 200, "Number of simulations (default: 200)"),
    ("num_trades", int, 400, "Number of trades per simulation (default: 400)"),
    ("num_mc_shuffles", int, 200, "Number of shuffles per simulation (default: 200)"),
    ("use_markov", None, False, "Use Markov chain correlations (1st order)"),
    ("p_win_after_win", float, 0.7, "P(win|win) for 1st order Markov model"),
    ("p_win_after_loss", float, 0.5, "P(win|loss) for 1st order Markov model"),
    ("use_markov2", None, False, "Use 2nd order Markov model"),
    ("p_win_ww", float, 0.8, "P(win|win,win) for 2nd order Markov"),
    ("p_win_wl", float, 0.6, "P(win|win,loss) for 2nd order Markov"),
    ("p_win_lw", float, 0.5, "P(win|loss,win) for 2nd order Markov"),
    ("p_win_ll", float, 0.3, "P(win|loss,loss) for 2nd order Markov"),
    ("use_regime", None, False, "Use regime switching model"),
    ("regimes", str, None, "Regime list as JSON string"),
)

# stdout layout of the reference table (:606-643; the scraper regex is dps.py:356)
COLUMNS = (("Strategy", "<90", "<90"), ("Avg Profit (€)", ">14", "14.2f"), ("Avg Drawdown (€)", ">16", "16.2f"), ("Ratio", ">12", "12.2f"),
           ("Min (€)", ">12", "12.2f"), ("Max (€)", ">12", "12.2f"), ("Min DD (€)", ">14", "14.2f"), ("Max DD (€)", ">14", "14.2f"),
           ("Avg/Trade", ">12", "12.2f"), ("Profit/MaxDD", ">18", "18.2f"))
MODE_NAMES = {"regime": "Regime Switching", "markov2": "2nd Order Markov", "markov1": "1st Order Markov", "none": "No Markov"}


def build_parser():
    parser = argparse.ArgumentParser(description="Simulate 20 trading strategies with/without Markov correlations, second-order Markov, and regime switching")
    for name, kind, default, text in CLI_FLAGS:
        if kind is None:
            parser.add_argument("--" + name, action="store_true", help=text)
        elif default is REQUIRED:
            parser.add_argument("--" + name, type=kind, required=True, help=text)
        else:
            parser.add_argument("--" + name, type=kind, default=default, help=text)
    return parser


def _mode_key(args):
    """generator precedence of the reference: regime > Markov-2 > Markov-1 > none (:477-490, :555-567)"""
    return "regime" if args.use_regime else "markov2" if args.use_markov2 else "markov1" if args.use_markov else "none"


def format_row(row):
    """One table row, reference :639-643 (matched by the scraper regex dps.py:356)."""
    return " ".join(format(value, spec) for value, (_, _, spec) in zip(row, COLUMNS))


def settings_block(args):
    """The header the scraper reads (dps.py:145-154): reference :552-584.  A generator: main() prints every line as
    it is produced, so the output up to a failure is the reference's too (for avg_win <= 0 the reference prints the
    settings, then find_break_even_hit_rate's error line, then dies formatting None -- :583-584)."""
    key = _mode_key(args)
    bar = "=" * 90
    yield "\n" + bar
    yield "CURRENT SIMULATION SETTING:"
    yield f"Hit rate: {args.hit_rate:.2%}"
    yield "Mode: " + MODE_NAMES[key]
    if key == "regime" and args.regimes:
        yield f"Regimes: {args.regimes}"
    elif key == "markov2":
        yield f"P(win|WW): {args.p_win_ww}, P(win|WL): {args.p_win_wl}, P(win|LW): {args.p_win_lw}, P(win|LL): {args.p_win_ll}"
    elif key == "markov1":
        yield f"P(win|win): {args.p_win_after_win}, P(win|loss): {args.p_win_after_loss}"
    yield bar + "\n"
    for label, value in (("Average win per trade: €", args.avg_win), ("Average loss per trade: €", args.avg_loss),
                         ("Number of simulations: ", args.num_simulations), ("Number of trades per simulation: ", args.num_trades),
                         ("Number of shuffles per simulation: ", args.num_mc_shuffles)):
        yield f"{label}{value}"
    breakeven = find_break_even_hit_rate(args.avg_win, args.avg_loss, MODE_NAMES[key])  # prints its own error line
    yield f"Break-even hit rate: {breakeven:.2%}"  # TypeError for None, exactly like :584


def main(argv=None):
    """stdout-compatible with reference :552-671 (settings block, break-even line, 20-row table,
    rule after the third row, optional colorama 'Top 4' block); byte-for-byte equality with the
    reference's output for the same rows is tested in tests/test_cli_cpu.py."""
    args = build_parser().parse_args(argv)
    key = _mode_key(args)
    for line in settings_block(args):
        print(line)

    regimes = pyjson.loads(args.regimes) if args.use_regime and args.regimes else None
    flags = {name: getattr(args, name) for name, kind, _, _ in CLI_FLAGS if name.startswith(("use_", "p_win_"))}
    summary = run_all_strategies(args.hit_rate, args.avg_win, args.avg_loss, args.num_trades, args.num_simulations,
                                 args.num_mc_shuffles, regimes=regimes, **flags)

    header = " ".join(format(title, spec) for title, spec, _ in COLUMNS)
    label = "Hit rate: %d%%  -  %s" % (int(round(args.hit_rate * 100)), MODE_NAMES[key] + (" Model" if key == "regime" else ""))
    try:
        from colorama import Fore, Style, init
        paint = lambda colour, text: colour + text + Style.RESET_ALL
    except ImportError:
        Fore = init = None
        paint = lambda colour, text: text
    print("\nResults (Monte Carlo, based on input parameters):\n")
    print(header)
    print("=" * len(header))
    print()
    print(paint(Fore.YELLOW if Fore else None, f"*** {label} ***"))
    print()
    for idx, row in enumerate(summary):
        print(format_row(row))
        if idx == 2:
            print("-" * len(header))

    if Fore is None:
        return
    # the "Top 4" comparison block exists only when colorama is importable (:647-671)
    init(autoreset=True)
    baseline = next((row for row in summary if row[0].startswith(DESCRIPTIONS[1])), None)
    print("\n\n\nTop 4 strategies compared to 'Constant position size 1':")
    print("--------------------------------------------------------------")
    for row in summary[:4]:
        print()
        print(paint(Fore.GREEN, format_row(row)))
        if baseline is not None:
            print(paint(Fore.RED, format_row(baseline)))


if __name__ == "__main__":
    main()
