"""The slice of `tf.app.flags` / `tf.app.run` the reference's experiment drivers use (`DEFINE_string|integer|float|boolean`,
`FLAGS.name`, `tf.app.run()` calling `main(_)`), so the drivers here keep the reference's flag names and defaults
(vdsr/vdsr/experiment_train.py:156-221, espcn/espcn/experiment_test.py:197-211, srcnn/srcnn.py:8-26).  One addition:
`--config file.yaml` overlays flag values from a YAML mapping before the command line is applied (the reference's
`config.yaml` files hold only the Cloud-ML machine description; its `trainingInput` block is accepted and ignored)."""
from __future__ import annotations

import sys


class _Flags:
    def __init__(self):
        object.__setattr__(self, "_defs", {})
        object.__setattr__(self, "_vals", {})

    def __getattr__(self, name):
        vals = object.__getattribute__(self, "_vals")
        if name in vals:
            return vals[name]
        raise AttributeError(f"unknown flag {name!r}")

    def __setattr__(self, name, value):
        self._vals[name] = value

    def __contains__(self, name):
        return name in self._vals


FLAGS = _Flags()


def _define(name, default, help_, kind):
    FLAGS._defs[name] = (kind, help_)
    FLAGS._vals.setdefault(name, default)


def DEFINE_string(name, default, help_=""):
    _define(name, default, help_, str)


def DEFINE_integer(name, default, help_=""):
    _define(name, default, help_, int)


def DEFINE_float(name, default, help_=""):
    _define(name, default, help_, float)


def DEFINE_boolean(name, default, help_=""):
    _define(name, default, help_, bool)


def _convert(kind, text):
    if kind is bool:
        if isinstance(text, bool):
            return text
        return str(text).lower() in ("1", "true", "t", "yes", "y")
    return kind(text)


def parse(argv=None):
    """--name=value | --name value | --flag | --noflag ; --config file.yaml first."""
    argv = list(sys.argv[1:] if argv is None else argv)
    pairs, i = [], 0
    while i < len(argv):
        a = argv[i]
        if not a.startswith("--"):
            raise SystemExit(f"unexpected argument {a!r}")
        body = a[2:]
        if "=" in body:
            k, v = body.split("=", 1)
        elif body in FLAGS._defs and FLAGS._defs[body][0] is bool:
            k, v = body, True
        elif body.startswith("no") and body[2:] in FLAGS._defs and FLAGS._defs[body[2:]][0] is bool:
            k, v = body[2:], False
        else:
            if i + 1 >= len(argv):
                raise SystemExit(f"flag --{body} needs a value")
            k, v = body, argv[i + 1]
            i += 1
        pairs.append((k, v))
        i += 1
    for k, v in pairs:
        if k == "config":
            import yaml
            with open(v) as f:
                for ck, cv in (yaml.safe_load(f) or {}).items():
                    if ck in FLAGS._defs:
                        FLAGS._vals[ck] = _convert(FLAGS._defs[ck][0], cv)
    for k, v in pairs:
        if k == "config":
            continue
        if k not in FLAGS._defs:
            raise SystemExit(f"unknown flag --{k}; known: {sorted(FLAGS._defs)}")
        FLAGS._vals[k] = _convert(FLAGS._defs[k][0], v)
    return FLAGS


def run(main=None, argv=None):
    """`tf.app.run()`: parse the flags, call `main(_)`."""
    parse(argv)
    if main is None:
        main = sys.modules["__main__"].main
    return main(None)
