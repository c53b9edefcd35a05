#!/usr/bin/env python
"""Overlay entry point: placed where dps.py expects `src/trading_models.py` (dps.py:219), it hands
over to the B200-native drop-in.  Same flags, same stdout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mc_trade_sizing_b200.trading_models import main  # noqa: E402

if __name__ == "__main__":
    main()
