"""Test-infrastructure stand-in for dataclass-wizard 0.35.1 (uv.lock:395-396 of the reference).

Only the surface the reference's config loading uses is provided
(systems/secondary/config.py:22,127; systems/secondary/__init__.py:243):
``from_dict`` / ``from_yaml_file`` / ``to_dict`` doing recursive dict -> dataclass
construction with str/int -> float coercion by annotation (PyYAML loads ``3000.0e6``
as a string).  It affects config loading only, never step arithmetic.
"""
import dataclasses
import typing


def _coerce(tp, val):
    if val is None:
        return None
    origin = typing.get_origin(tp)
    if tp is typing.Any or tp is None:
        return val
    if origin is typing.Union:
        args = [a for a in typing.get_args(tp) if a is not type(None)]
        for a in args:
            try:
                return _coerce(a, val)
            except Exception:
                continue
        return val
    if origin in (list, typing.List):
        (a,) = typing.get_args(tp) or (typing.Any,)
        return [_coerce(a, v) for v in val]
    if origin in (tuple, typing.Tuple):
        args = typing.get_args(tp)
        if len(args) == 2 and args[1] is Ellipsis:
            return tuple(_coerce(args[0], v) for v in val)
        if args:
            return tuple(_coerce(a, v) for a, v in zip(args, val))
        return tuple(val)
    if origin in (dict, typing.Dict):
        args = typing.get_args(tp)
        if len(args) == 2:
            return {_coerce(args[0], k): _coerce(args[1], v) for k, v in val.items()}
        return dict(val)
    if isinstance(tp, type):
        if dataclasses.is_dataclass(tp):
            if isinstance(val, tp):
                return val
            if isinstance(val, dict):
                return _build(tp, val)
            return val
        if tp is float:
            if isinstance(val, bool):
                return float(val)
            if isinstance(val, (int, float, str)):
                return float(val)
            return val
        if tp is int:
            if isinstance(val, bool):
                return int(val)
            if isinstance(val, str):
                return int(float(val))
            if isinstance(val, float) and val == int(val):
                return int(val)
            return val
        if tp is bool:
            if isinstance(val, str):
                return val.strip().lower() in ("1", "true", "yes", "on")
            return bool(val)
        if tp is str:
            return val if isinstance(val, str) else str(val)
        import enum
        if issubclass(tp, enum.Enum):
            if isinstance(val, tp):
                return val
            try:
                return tp(val)
            except Exception:
                return tp[val]
    return val


def _build(cls, data):
    hints = typing.get_type_hints(cls)
    kwargs = {}
    post = {}
    for f in dataclasses.fields(cls):
        if f.name in data:
            v = _coerce(hints.get(f.name, typing.Any), data[f.name])
            if f.init:
                kwargs[f.name] = v
            else:
                post[f.name] = v
    obj = cls(**kwargs)
    for k, v in post.items():
        setattr(obj, k, v)
    return obj


class _Wizard:
    @classmethod
    def from_dict(cls, data):
        return _build(cls, data)

    @classmethod
    def from_yaml_file(cls, path):
        import yaml
        with open(path) as fh:
            return _build(cls, yaml.safe_load(fh))

    @classmethod
    def from_yaml(cls, text):
        import yaml
        return _build(cls, yaml.safe_load(text))

    @classmethod
    def from_json_file(cls, path):
        import json
        with open(path) as fh:
            return _build(cls, json.load(fh))

    def to_dict(self):
        return dataclasses.asdict(self)

    def to_yaml_file(self, path):
        import yaml
        with open(path, "w") as fh:
            yaml.safe_dump(dataclasses.asdict(self), fh)

    def to_json_file(self, path):
        import json
        with open(path, "w") as fh:
            json.dump(dataclasses.asdict(self), fh)


class YAMLWizard(_Wizard):
    pass


class JSONWizard(_Wizard):
    pass


class TOMLWizard(_Wizard):
    pass
