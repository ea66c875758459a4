"""Simulation-day helper keyed by a per-process environment flag (reference config/simulation_date.py:8-66)."""

import datetime
import os


def _flag() -> str:
    return f"DYNODE_INITIALIZATION_DATE({os.getpid()})"


def get_dynode_init_date_flag():
    raw = os.getenv(_flag())
    return None if raw is None else datetime.datetime.strptime(raw, "%Y-%m-%d").date()


def set_dynode_init_date_flag(init_date: datetime.date) -> None:
    os.environ[_flag()] = init_date.strftime("%Y-%m-%d")


def simulation_day(year: int, month: int, day: int) -> int:
    """Days between the model's initialisation date and the given date (negative if earlier)."""
    init = get_dynode_init_date_flag()
    if init is None:
        raise ValueError("attempting to use SimulationDate helper method without first calling "
                         "set_dynode_init_date_flag() to set env flag.")
    return (datetime.date(year, month, day) - init).days
