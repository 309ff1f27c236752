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


class DeviceFeed:
    """Double-buffered host -> device feed of a training loop's batches into the STATIC input tensors of a graphed step
    (`make_graphed_step`): the copy of batch k+1 from page-locked host memory runs on a side stream while step k computes; at the
    start of step k+1 a device-to-device copy (microseconds) moves it into the static tensors.  What `feed_dict` does in the
    reference's loop (vdsr/vdsr/experiment_train.py:140-160), without stalling the device on PCIe.

        feed = DeviceFeed([sd_static, hd_static])
        feed.put([sd_host, hd_host])          # prefetch the first batch
        for batch in batches:                 # steady state
            feed.take()                       # batch k is in the static tensors (ordered on the current stream)
            feed.put(next_batch)              # batch k+1 starts copying in the background
            step(lr)
    """

    def __init__(self, static_tensors):
        import torch
        self._torch = torch
        self.static = list(static_tensors)
        self.stage = [[torch.empty_like(t) for t in self.static] for _ in range(2)]
        self.stream = torch.cuda.Stream()
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.taken = [None, None]
        self.n_put = self.n_take = 0

    def put(self, host_tensors) -> None:
        assert self.n_put - self.n_take < 2, "DeviceFeed holds at most two batches"
        torch = self._torch
        k = self.n_put & 1
        self.n_put += 1
        with torch.cuda.stream(self.stream):
            if self.taken[k] is not None:
                self.stream.wait_event(self.taken[k])  # the device-to-device copy that last read this staging set is done
            for dst, src in zip(self.stage[k], host_tensors):
                src = src if isinstance(src, torch.Tensor) else torch.from_numpy(src)
                dst.copy_(src, non_blocking=True)
            self.ready[k].record(self.stream)

    def take(self) -> None:
        assert self.n_take < self.n_put, "DeviceFeed.take() without a batch"
        torch = self._torch
        k = self.n_take & 1
        self.n_take += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[k])
        for dst, src in zip(self.static, self.stage[k]):
            dst.copy_(src, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.taken[k] = ev
