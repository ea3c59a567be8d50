"""Parameters of the co-event counting stage.

Mirrors the co-count block of the reference's ``config.py`` (lines 41-96): same names, same
defaults, as a frozen dataclass instead of module globals (importing the reference's config has
side effects -- it creates ``artifacts/`` and opens a log file, config.py:13-27 -- which a library
must not have).  ``DIR_DATA`` defaults to ``$OTTO_DIR_DATA`` or ``./data``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Tuple


def _dir_data() -> str:
    return os.environ.get("OTTO_DIR_DATA", os.path.join(os.getcwd(), "data"))


@dataclass(frozen=True)
class CoEventConfig:
    DIR_DATA: str = field(default_factory=_dir_data)
    # config.py:41-49
    MIN_TIME_TO_NEXT: int = -24 * 60 * 60
    MAX_TIME_TO_NEXT: int = 24 * 60 * 60
    MAP_MAX_TIME_TO_NEXT: Dict[str, int] = field(default_factory=lambda: {
        "click_to_click": 12 * 60 * 60,
        "click_to_cart_or_buy": 24 * 60 * 60,
        "cart_to_cart": 24 * 60 * 60,
        "cart_to_buy": 24 * 60 * 60,
        "buy_to_buy": 24 * 60 * 60,
    })
    # config.py:52-53 (row-count triggers of the lossy merge steps)
    OPTIM_ROWS_POLARS_GROUPBY: int = 100_000_000
    MAX_ROWS_POLARS_GROUPBY: int = 300_000_000
    # the literal 100_000_000 of count_co_events.py:131 (a field only so tests can reach that branch)
    ROWS_TRIGGER_MIN_COUNT_IN_PART: int = 100_000_000
    # config.py:56-64
    MIN_COUNT_TO_SAVE: Dict[str, int] = field(default_factory=lambda: {
        "click_to_click": 10,
        "click_to_cart_or_buy": 5,
        "cart_to_cart": 2,
        "cart_to_buy": 2,
        "buy_to_buy": 2,
    })
    MIN_COUNT_IN_PART: Dict[str, int] = field(default_factory=lambda: {
        "click_to_click": 2, "click_to_cart_or_buy": 2})
    MAX_CO_EVENT_PAIRS_TO_SAVE_DISK: int = 300_000_000
    # config.py:67-73
    CO_EVENTS_TO_COUNT: Tuple[str, ...] = (
        "click_to_click", "click_to_cart_or_buy", "cart_to_cart", "cart_to_buy", "buy_to_buy")
    # config.py:81-88  name -> (type of this event, types of the next event)
    MAP_NAME_COUNT_TYPE: Dict[str, Tuple[int, List[int]]] = field(default_factory=lambda: {
        "click_to_click": (0, [0]),
        "click_to_cart_or_buy": (0, [1, 2]),
        "cart_to_cart": (1, [1]),
        "cart_to_buy": (1, [2]),
        "buy_to_buy": (2, [2]),
    })
    # config.py:90-96
    RETRIEVAL_FIRST_N_CO_COUNTS: Dict[str, int] = field(default_factory=lambda: {
        "click_to_click": 10,
        "click_to_cart_or_buy": 10,
        "cart_to_cart": 20,
        "cart_to_buy": 20,
        "buy_to_buy": 20,
    })
    # engine-only knobs (no reference counterpart)
    TOP_K: int = 20                 # north-star default for the segmented top-K
    PAIR_BUDGET: int = 0            # co-event pairs expanded at once; 0 = sized from free HBM
    EXACT_MERGE: bool = False       # True: never apply the row-count-triggered lossy merge steps

    def spec(self, name: str) -> Tuple[int, int, int]:
        """name -> (type_this, next_mask, window_seconds)."""
        this, nxt = self.MAP_NAME_COUNT_TYPE[name]
        mask = 0
        for t in nxt:
            mask |= 1 << int(t)
        return int(this), mask, int(self.MAP_MAX_TIME_TO_NEXT[name])


DEFAULT_CONFIG = CoEventConfig()
