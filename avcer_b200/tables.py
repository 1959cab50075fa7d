"""Zero-copy façade for the long-format audio table.

The reference's audio driver returns a DataFrame with ONE ROW PER (window, covered frame): the window's logits
replicated ~100 times each, plus a `frames` column of strings "000123.jpg" (src/get_prob_audio_8_cl.py:94-126); the fusion
stage then groups those rows by the string, averages, and parses the strings back into frame numbers
(src/run.py:90-97, src/get_pred_av.py:232-278).  After the models run in milliseconds, building and re-parsing
O(windows x 100) Python strings per clip is what the wall clock shows.

`AudioTable` is what the drop-in drivers return instead: it HOLDS the compact form -- per-window logits (on the device)
and each window's frame range [f_lo, f_hi) -- and BEHAVES like the reference's DataFrame: any attribute, item access,
`len()`, iteration or `to_csv` materialises the long table once (vectorised, same columns / dtypes / row order / strings) and
delegates to it.  The accelerated fusion (`avcer_b200.run.audio_frame_rows`) recognises the façade and goes straight from
the window logits to per-frame means on the GPU (avcer_window_to_frame_mean: pandas' float32 Kahan group mean, NaN windows
skipped) without a single string.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd
import torch


def frame_name_column(f_lo: np.ndarray, f_hi: np.ndarray) -> np.ndarray:
    """The `frames` column of the long table, vectorised: "{i:06d}.jpg" for every i of every window's range, window after
    window (get_prob_audio_8_cl.py:94-101)."""
    counts = np.maximum(f_hi - f_lo, 0)
    if counts.sum() == 0:
        return np.zeros(0, dtype=object)
    starts = np.repeat(f_lo, counts)
    within = np.arange(counts.sum()) - np.repeat(np.cumsum(counts) - counts, counts)
    ids = (starts + within).astype(np.int64)
    return np.char.add(np.char.zfill(ids.astype(str), 6), ".jpg").astype(object)


class AudioTable:
    """Compact long-format audio table with a lazily materialised DataFrame behind it (module docstring)."""

    def __init__(self, window_logits: torch.Tensor, f_lo: np.ndarray, f_hi: np.ndarray, columns: List[str]):
        object.__setattr__(self, "window_logits", window_logits)          # [Wn, ncls] float32, device or host
        object.__setattr__(self, "f_lo", np.asarray(f_lo, dtype=np.int64))
        object.__setattr__(self, "f_hi", np.asarray(f_hi, dtype=np.int64))
        object.__setattr__(self, "value_columns", list(columns))
        object.__setattr__(self, "_df", None)

    # ------------------------------------------------------------------ compact access (what the fusion fast path uses)
    @property
    def materialized(self) -> bool:
        return self._df is not None

    def frame_ids(self) -> np.ndarray:
        """Sorted unique frame ids covered by at least one window (the groups of groupby("frames"))."""
        hi = int(self.f_hi.max()) if len(self.f_hi) else 0
        cover = np.zeros(hi + 1, dtype=np.int64)
        np.add.at(cover, np.minimum(self.f_lo, hi), 1)
        np.add.at(cover, np.minimum(np.maximum(self.f_hi, self.f_lo), hi), -1)
        return np.nonzero(np.cumsum(cover)[:hi] > 0)[0]

    # ------------------------------------------------------------------ DataFrame behaviour (lazy)
    def materialize(self) -> pd.DataFrame:
        if self._df is None:
            logits = self.window_logits.detach().cpu().numpy() if isinstance(self.window_logits, torch.Tensor) else np.asarray(self.window_logits)
            counts = np.maximum(self.f_hi - self.f_lo, 0)
            df = pd.DataFrame(np.repeat(logits, counts, axis=0), columns=self.value_columns)
            df["frames"] = frame_name_column(self.f_lo, self.f_hi)
            object.__setattr__(self, "_df", df)
        return self._df

    def __getattr__(self, name):                      # only reached for names this class does not define
        return getattr(self.materialize(), name)

    def __setattr__(self, name, value):
        setattr(self.materialize(), name, value)

    def __getitem__(self, key):
        return self.materialize()[key]

    def __setitem__(self, key, value):
        self.materialize()[key] = value

    def __len__(self) -> int:
        return int(np.maximum(self.f_hi - self.f_lo, 0).sum()) if self._df is None else len(self._df)

    def __iter__(self):
        return iter(self.materialize())

    def __repr__(self) -> str:
        state = "materialised" if self.materialized else "compact"
        return f"AudioTable({len(self.f_lo)} windows x {len(self.value_columns)} classes -> {len(self)} rows, {state})"
