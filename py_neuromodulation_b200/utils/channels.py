"""Channel table helpers (reference: ``utils/channels.py``).

The channel table is a pandas DataFrame with the columns
``name, rereference, used, target, type, status, new_name``.
"""

from __future__ import annotations

from collections.abc import Iterable

import numpy as np

COLUMNS = ["name", "rereference", "used", "target", "type", "status", "new_name"]
_LFP_TYPES = ("seeg", "dbs", "lfp")


def _as_list(value) -> list[str]:
    if value is None:
        return []
    return [value] if isinstance(value, str) else list(value)


def _default_references(names: list[str], types: list[str]) -> list[str]:
    """ECoG -> common average; LFP/DBS/sEEG -> bipolar to the neighbouring contact of the same side."""
    refs: dict[str, str] = {}
    ecog = [n for n, t in zip(names, types) if "ecog" in t.lower() or "ecog" in n.lower()]
    lfp = [n for n, t in zip(names, types)
           if n not in ecog and any(k in t.lower() or k in n.lower() for k in _LFP_TYPES)]
    if len(ecog) > 1:
        for n in ecog:
            refs[n] = "average"
    for tags in (("_l_", "_left_"), ("_r_", "_right_")):
        side = sorted(n for n in lfp if any(tag in n.lower() for tag in tags))
        if len(side) > 1:
            for i, n in enumerate(side):
                refs[n] = side[i - 1] if i > 0 else side[-1]
    return [refs.get(n, "None") for n in names]


def set_channels(
    ch_names: list[str],
    ch_types: list[str],
    reference: list | str | None = "default",
    bads: list[str] | str | None = None,
    new_names: str | list[str] | None = "default",
    ecog_only: bool = False,
    used_types: Iterable[str] | None = ("ecog", "dbs", "seeg"),
    target_keywords: Iterable[str] | None = ("mov", "squared", "label"),
):
    """Build a channel table (reference ``utils/channels.py:13-203``)."""
    import pandas as pd

    if len(ch_names) != len(ch_types):
        raise ValueError(
            f"Number of `ch_names` and `ch_types` must match.Got: {len(ch_names)} `ch_names` and {len(ch_types)} `ch_types`."
        )
    names = list(ch_names)
    types = list(ch_types)
    used_kinds = [u.lower() for u in _as_list(used_types)]
    used = [int(t.lower() in used_kinds) for t in types]
    keywords = [k.lower() for k in _as_list(target_keywords)]
    target = [int(any(k in n.lower() for k in keywords)) for n in names]
    if ecog_only:
        used = [0 if t in ("seeg", "dbs") else u for u, t in zip(used, types)]

    if isinstance(reference, str):
        if reference.lower() == "default":
            refs = _default_references(names, types)
        elif reference.lower() == "average":
            refs = ["average" if u == 1 else "None" for u in used]
        else:
            raise ValueError(
                "`reference` must be either `default`, `None`, `average` or an iterable of new reference channel "
                f"names. Got: {reference}."
            )
    elif isinstance(reference, list):
        if len(reference) != len(names):
            raise ValueError(
                f"Number of `ch_names` and `reference` must match.Got: {len(names)} `ch_names` and {len(reference)} `references`."
            )
        refs = list(reference)
    elif not reference:
        refs = ["None"] * len(names)
    else:
        raise ValueError(
            f"`reference` must be either `default`, None or an iterable of new reference channel names. Got: {reference}."
        )

    bad = set(_as_list(bads))
    status = ["bad" if n in bad else "good" for n in names]
    used = [0 if s == "bad" else u for u, s in zip(used, status)]

    if not new_names:
        renamed = list(names)
    elif isinstance(new_names, str):
        if new_names.lower() != "default":
            raise ValueError(
                f"`new_names` must be either `default`, None or an iterable of new channel names. Got: {new_names}."
            )
        renamed = []
        for n, r in zip(names, refs):
            if r == "None" or (isinstance(r, float) and np.isnan(r)):
                renamed.append(n)
            elif r == "average":
                renamed.append(n + "_avgref")
            else:
                renamed.append(f"{n}_{r}")
    else:
        if len(new_names) != len(names):
            raise ValueError(
                f"Number of `ch_names` and `new_names` must match. Got: {len(names)} `ch_names` and {len(new_names)} `new_names`."
            )
        renamed = list(names)  # the reference keeps the original names in this branch
    return pd.DataFrame(
        {"name": names, "rereference": refs, "used": used, "target": target, "type": types, "status": status,
         "new_name": renamed},
        columns=COLUMNS,
    )


def get_default_channels_from_data(data, car_rereferencing: bool = True):
    """All channels ECoG, good, used, no targets (reference ``utils/channels.py:257-309``).

    As in the reference the new names always carry the ``_avgref`` suffix, also when
    ``car_rereferencing`` is False.
    """
    import pandas as pd

    n = data.shape[0]
    names = [f"ch{i}" for i in range(n)]
    return pd.DataFrame(
        {
            "name": names,
            "rereference": ["average" if car_rereferencing else "None"] * n,
            "used": np.ones(n, dtype=int),
            "target": np.zeros(n, dtype=int),
            "type": ["ecog"] * n,
            "status": ["good"] * n,
            "new_name": [f"{c}_avgref" for c in names],
        },
        columns=COLUMNS,
    )
