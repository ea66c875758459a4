"""A parameter whose value is derived from another (possibly sampled) parameter."""

from typing import Any, Callable, Optional, Union


def _identity(x):
    return x


class DeterministicParameter:
    """Link to `parameter_state[depends_on]` (optionally indexed, optionally transformed)."""

    def __init__(self, depends_on: str, index: Optional[Union[int, tuple, slice]] = None,
                 transform: Callable[[Any], Any] = _identity):
        self.depends_on = depends_on
        self.index = index
        self.transform = transform

    def resolve(self, parameter_state: dict) -> Any:
        """Value of the linked parameter inside `parameter_state` (raises with the scope on failure)."""
        try:
            value = parameter_state[self.depends_on]
            if self.index is not None:
                value = value[self.index]
            return self.transform(value)
        except Exception as e:
            where = self.depends_on if self.index is None else f"{self.depends_on}[{self.index}]"
            raise Exception(
                f"Was unable to find {where} within the following scope, make sure "
                f"DeterministicParameter dependencies are at the top level of the configuration "
                f"object and indexes are correct. Scope: {parameter_state}") from e
