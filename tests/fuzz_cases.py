"""Random parameter sets for the generator / evaluator fuzz tests (CPU harness and GPU)."""
import numpy as np


def random_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    probs = [0.0, 1.0, 0.05, 0.5, 0.82, 0.999999, 1e-9]
    for _ in range(n):
        mode = rng.integers(0, 4)
        T = int(rng.choice([1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 40, 100, 127, 128, 129, 200, 257, 500, 1000, 1203]))
        ns, nsh = int(rng.integers(1, 6)), int(rng.integers(1, 5))
        bad = rng.random() < 0.1     # now and then a probability outside [0, 1]: np.random.choice raises, rand() < p does not
        pick = lambda: (float(rng.choice([-0.2, 1.3])) if bad and rng.random() < 0.5 else
                        float(rng.choice(probs)) if rng.random() < 0.4 else float(rng.random()))
        val = lambda: float(rng.choice([0.0, 1.0, 100.0, 186.0, 219.0, 0.1, 204.60000000000002, -50.0, 1e-3, 7.25e6])) \
            if rng.random() < 0.6 else float(np.round(rng.normal() * 100, int(rng.integers(0, 4))))
        args = [pick() if (mode != 0 or bad) else float(rng.uniform(0.0, 0.8)), val(), val(), T, ns, nsh]
        kw = {}
        if mode == 1:
            kw = dict(use_markov=True, p_win_after_win=pick(), p_win_after_loss=pick())
        elif mode == 2:
            kw = dict(use_markov2=True, p_win_ww=pick(), p_win_wl=pick(), p_win_lw=pick(), p_win_ll=pick())
        elif mode == 3:
            if rng.random() < 0.25:
                kw = dict(use_regime=True, regimes=None)
            else:
                nr = int(rng.integers(1, 7)) if rng.random() < 0.8 else int(rng.integers(7, 17))
                kw = dict(use_regime=True, regimes=[{"length": int(rng.integers(0, max(2, T))), "hit_rate": pick(),
                                                     "avg_win": val(), "avg_loss": val()} for _ in range(nr)])
        cases.append((tuple(args), kw))
    return cases
