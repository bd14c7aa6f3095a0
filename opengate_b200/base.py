"""User-info machinery of the host-side mirror (the attribute surface of the reference's
`GateObject`, opengate/base.py:362-727, for the objects on the photon-dose path only).

Every façade object declares `user_info_defaults = {name: (default, options)}`; the values live
in `self.user_info` (a `Box`) and are reachable as plain attributes.  Behaviour kept from the
reference: defaults are deep-copied per instance, `allowed_values` are enforced on assignment,
assigning a `deprecated` parameter raises `GateDeprecationError` (base.py:583-589), assigning an
unknown attribute only warns ("Maybe a typo?", base.py:598-601), `setter_hook`s run on assignment.
"""

from __future__ import annotations

import copy

from .exception import fatal, warning


class GateDeprecationError(Exception):
    pass


class Box(dict):
    """dict with attribute access (the reference uses python-box for `position`, `direction`,
    `energy` of sources, opengate/sources/generic.py:24-77)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Box({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def copy(self):
        return Box(dict.copy(self))


def _collect_defaults(cls):
    out = {}
    for c in reversed(cls.mro()):
        out.update(c.__dict__.get("user_info_defaults", {}))
    return out


class GateObject:
    user_info_defaults = {"name": (None, {"doc": "Unique name of the object", "required": True})}

    def __init__(self, *args, simulation=None, **kwargs):
        object.__setattr__(self, "_known", set())
        object.__setattr__(self, "simulation", simulation)
        defaults = _collect_defaults(type(self))
        object.__setattr__(self, "_defaults", defaults)
        ui = Box()
        for k, (v, _opt) in defaults.items():
            ui[k] = copy.deepcopy(v)
        object.__setattr__(self, "user_info", ui)
        if "name" in kwargs:
            ui["name"] = kwargs.pop("name")
        elif args:
            ui["name"] = args[0]
        for k, v in kwargs.items():
            if k not in defaults:
                fatal(f"Unknown user info '{k}' for {type(self).__name__}")
            setattr(self, k, v)
        if ui.get("name") is None:
            fatal(f"No 'name' provided for the {type(self).__name__}")

    @property
    def type_name(self):
        return type(self).__name__

    def __getattr__(self, key):
        # only called when normal lookup fails
        ui = self.__dict__.get("user_info")
        if ui is not None and key in ui:
            return ui[key]
        raise AttributeError(f"{type(self).__name__} has no attribute '{key}'")

    def __setattr__(self, key, value):
        defaults = self.__dict__.get("_defaults", {})
        if key in defaults:
            opt = defaults[key][1]
            if "deprecated" in opt:
                raise GateDeprecationError(opt["deprecated"])
            allowed = opt.get("allowed_values")
            if allowed is not None and not isinstance(value, (list, dict)) and value not in allowed:
                fatal(f"'{value}' is not an allowed value for '{key}' of {self.type_name} '{self.name}'. "
                      f"Allowed values: {tuple(allowed)}")
            hook = opt.get("setter_hook")
            if hook is not None:
                value = hook(self, value)
            self.user_info[key] = value
            return
        if isinstance(getattr(type(self), key, None), property):
            object.__setattr__(self, key, value)
            return
        if not key.startswith("_") and key not in self.__dict__ and not hasattr(type(self), key) \
                and self.__dict__.get("_frozen", False):
            warning(f'For object "{self.name}", attribute "{key}" is not known. Maybe a typo?')
        object.__setattr__(self, key, value)

    def _freeze(self):
        """Called at the end of a subclass __init__: later unknown attributes warn."""
        object.__setattr__(self, "_frozen", True)

    def close(self):
        pass

    def to_dictionary(self):
        return {"user_info": copy.deepcopy(dict(self.user_info)), "object_type": type(self).__name__}

    def __repr__(self):
        return f"{type(self).__name__}({self.user_info.get('name')!r})"
