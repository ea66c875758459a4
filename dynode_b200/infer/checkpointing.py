"""Record compartment sizes at key days as deterministic sites (API of reference
src/dynode/infer/checkpointing.py:12-47).  Reads the engine's `Solution.ys[idx][k]` only."""

import datetime
from typing import List, Optional

from . import ppl


def checkpoint_compartment_sizes(config, solution, save_final_timesteps: bool = True,
                                 compartment_save_dates: Optional[List[datetime.date]] = None):
    assert solution.ys is not None, "solution.ys returned None, odes failed."
    names = {name: int(idx) for name, idx in vars(config.idx).items() if not name.startswith("_")}
    if save_final_timesteps:
        for name, idx in names.items():
            ppl.deterministic(f"final_timestep_{name}", solution.ys[idx][-1])
    for date in compartment_save_dates or []:
        sim_day = (date - config.initializer.initialize_date).days
        if 0 <= sim_day < len(solution.ys[0]):
            stamp = date.strftime("%Y_%m_%d")
            for name, idx in names.items():
                ppl.deterministic(f"{stamp}_timestep_{name}", solution.ys[idx][sim_day])
