"""Minimal stand-in for the slice of the TensorFlow-1.x session API the reference's experiment
drivers use (`tf.placeholder`, `tf.Session().run(fetches, feed_dict)`), so those loops read the same
here: numpy NHWC float32 arrays in through `feed_dict`, numpy arrays / python scalars out
(vdsr/vdsr/experiment_train.py:123-153, espcn/espcn/experiment_test.py:164-169)."""
from __future__ import annotations


class Placeholder:
    """`tf.placeholder(shape=..., dtype=tf.float32, name=...)` (vdsr/vdsr/experiment_train.py:17-26)."""

    def __init__(self, name: str, shape=None, variable_of=None):
        self.name = name
        self.shape = shape
        self.variable_of = variable_of  # graph whose variable this feed overrides (e.g. learning_rate)

    def __repr__(self):
        return f"Placeholder({self.name!r}, shape={self.shape})"


def placeholder(shape=None, name: str = "placeholder") -> Placeholder:
    return Placeholder(name, shape)


class Handle:
    """A fetchable value of one model graph (`model['sr_images']`, `model['loss']`, ...)."""

    def __init__(self, graph, key: str):
        self.graph = graph
        self.key = key

    def __repr__(self):
        return f"Handle({self.key!r})"


def pinned_empty(shape, dtype="float32"):
    """A page-locked numpy array (backed by a pinned torch tensor that stays alive with it): feeds / fetch buffers of this kind
    are DMA'd directly, without a staging copy."""
    import numpy as np
    import torch
    t = torch.empty(tuple(shape), dtype=getattr(torch, str(np.dtype(dtype)))).pin_memory()
    a = t.numpy()
    _PINNED[a.__array_interface__["data"][0]] = t
    return a


_PINNED: dict = {}


def as_host_tensor(x, dtype=None):
    """numpy array / CPU tensor -> CPU tensor sharing its memory; arrays from `pinned_empty` come back as their pinned tensor."""
    import numpy as np
    import torch
    if isinstance(x, torch.Tensor):
        return x
    x = np.asarray(x)
    if x.flags["C_CONTIGUOUS"]:
        t = _PINNED.get(x.__array_interface__["data"][0])
        if t is not None and t.numel() == x.size and t.numpy().dtype == x.dtype:
            return t.view(x.shape) if t.shape != x.shape else t
    x = np.ascontiguousarray(x, dtype=dtype) if dtype is not None else np.ascontiguousarray(x)
    return torch.from_numpy(x)


class Session:
    """`with Session() as session: session.run(fetches, feed_dict=...)`."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None, out=None):
        """`out` (extension, optional): {Handle: preallocated host array} -- the fetch is written into that array (page-locked
        arrays avoid a staging copy; see `pinned_empty`) and the same array is returned, like a TF callable with a fetch buffer."""
        feed_dict = feed_dict or {}
        out = out or {}
        flat = []

        def collect(f):
            if isinstance(f, Handle):
                flat.append(f)
            elif isinstance(f, dict):
                for v in f.values():
                    collect(v)
            elif isinstance(f, (list, tuple)):
                for v in f:
                    collect(v)
            elif f is None or isinstance(f, Placeholder):
                pass
            else:
                raise TypeError(f"cannot fetch {f!r}")

        collect(fetches)
        results = {}
        graphs = []
        for h in flat:
            if h.graph not in graphs:
                graphs.append(h.graph)
        for g in graphs:
            keys = {h.key for h in flat if h.graph is g}
            outs = {h.key: a for h, a in out.items() if h.graph is g}
            vals = g.execute(keys, feed_dict, outs) if outs else g.execute(keys, feed_dict)
            for k in keys:
                results[(id(g), k)] = vals.get(k)

        def rebuild(f):
            if isinstance(f, Handle):
                return results[(id(f.graph), f.key)]
            if isinstance(f, dict):
                return {k: rebuild(v) for k, v in f.items()}
            if isinstance(f, (list, tuple)):
                return type(f)(rebuild(v) for v in f)
            if isinstance(f, Placeholder):
                return feed_dict.get(f)
            return None

        return rebuild(fetches)
