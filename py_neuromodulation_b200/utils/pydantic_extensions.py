"""Pydantic helpers behind the settings models (API of the reference's
``utils/pydantic_extensions.py``: ``NMBaseModel``, ``NMField``, ``NMErrorList``,
``create_validation_error``)."""

from __future__ import annotations

import copy
from pprint import pformat
from typing import Any, Sequence

from pydantic import BaseModel, ConfigDict, Field
from pydantic_core import InitErrorDetails, ValidationError


def create_validation_error(error_message: str, location: Sequence[str | int] = (), title: str = "Validation Error",
                            error_type: str = "value_error") -> ValidationError:
    """Build a pydantic ``ValidationError`` carrying one custom message."""
    details = InitErrorDetails(type=error_type, loc=tuple(location), input=None, ctx={"error": error_message})
    return ValidationError.from_exception_data(title=title, line_errors=[details], input_type="python", hide_input=False)


class NMErrorList:
    """Accumulates validation problems so that they can be raised together."""

    def __init__(self, errors: Sequence[Any] | None = None) -> None:
        self._items: list[Any] = list(errors) if errors is not None else []

    def add_error(self, error_message: str, location: Sequence[str | int] = (), error_type: str = "value_error") -> None:
        self._items.append(InitErrorDetails(type=error_type, loc=tuple(location), input=None, ctx={"error": error_message}))

    def extend(self, other: "NMErrorList") -> None:
        self._items.extend(other._items)

    def create_error(self, title: str = "Validation Error") -> ValidationError:
        cleaned = []
        for e in self._items:
            e = dict(e)
            # errors harvested from a caught ValidationError carry extra keys pydantic refuses on re-entry
            cleaned.append({k: e[k] for k in ("type", "loc", "input", "ctx") if k in e})
            if cleaned[-1]["type"] != "value_error":
                cleaned[-1] = {"type": "value_error", "loc": cleaned[-1].get("loc", ()), "input": cleaned[-1].get("input"),
                               "ctx": {"error": str(e.get("msg", e.get("type")))}}
        return ValidationError.from_exception_data(title=title, line_errors=cleaned)

    def __iter__(self):
        return iter(self._items)

    def __len__(self) -> int:
        return len(self._items)

    def __getitem__(self, idx):
        return copy.deepcopy(self._items[idx])

    def __repr__(self) -> str:
        return repr(self._items)


def NMField(default: Any = ..., *, custom_metadata: dict[str, Any] | None = None, **kwargs: Any) -> Any:
    """``pydantic.Field`` that also remembers GUI metadata (units, widget hints)."""
    extra = dict(kwargs.pop("json_schema_extra", None) or {})
    if custom_metadata:
        extra["custom_metadata"] = dict(custom_metadata)
    if extra:
        kwargs["json_schema_extra"] = extra
    return Field(default, **kwargs)


class NMBaseModel(BaseModel):
    """Base of every settings model: item access, positional construction, re-validation."""

    model_config = ConfigDict(validate_assignment=False, extra="allow")

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        if args:
            names = list(type(self).model_fields.keys())
            if len(args) > len(names):
                raise ValueError(f"Too many positional arguments. Expected at most {len(names)}, got {len(args)}")
            for name, value in zip(names, args):
                if name in kwargs:
                    raise ValueError(f"Got multiple values for field '{name}': positional argument and keyword argument")
                kwargs[name] = value
        super().__init__(**kwargs)

    __init__.__pydantic_base_init__ = True  # type: ignore[attr-defined]

    def __str__(self) -> str:
        return pformat(self.model_dump())

    def validate(self, context: Any | None = None) -> Any:  # type: ignore[override]
        """Re-run validation on the current field values and return the validated COPY."""
        return self.model_validate(self.model_dump(), context=context)

    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def __setitem__(self, key: str, value: Any) -> None:
        setattr(self, key, value)

    @property
    def fields(self):
        return type(self).model_fields

    @classmethod
    def unvalidated(cls, **data: Any) -> Any:
        """Construct without validation (used to keep going after a failed validation)."""
        values = {}
        for name, field in cls.model_fields.items():
            if name in data:
                value = data[name]
                ann = field.annotation
                if isinstance(value, dict) and isinstance(ann, type) and issubclass(ann, NMBaseModel):
                    value = ann.unvalidated(**value)
                values[name] = value
            elif not field.is_required():
                values[name] = copy.deepcopy(field.default)
            else:
                raise TypeError(f"Missing required keyword argument {name!r}")
        obj = cls.__new__(cls)
        object.__setattr__(obj, "__dict__", values)
        object.__setattr__(obj, "__pydantic_private__", None)
        object.__setattr__(obj, "__pydantic_extra__", {})
        object.__setattr__(obj, "__pydantic_fields_set__", set(values))
        return obj
