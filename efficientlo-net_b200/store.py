"""Variable store: the reference's implicit tf.variable_scope / tf.get_variable lookup, made explicit.

The blocks of pointnet_util.py take a ``scope`` string exactly like the reference's; the weights they
need are looked up in the store that is current (``with use_store(store):``) under the enclosing
``variable_scope`` prefixes, e.g. ``with variable_scope('sa1'): down_conv(..., scope='layer0')`` reads
``sa1/layer0/conv0/weights`` etc. -- the names of the shipped checkpoint.
Packed, BN-folded device copies are cached per (layer list, device).
"""
import contextlib
import threading

import torch

from . import packing


class _Stacks(threading.local):
    """The current-store and scope stacks are per host thread: two threads that drive two engines (or two emulated
    ranks of a row-band group) must not see each other's store -- its scratch buffers would be shared."""

    def __init__(self):
        self.current, self.prefix = [], []


_tls = _Stacks()


def current_stack():
    """The calling thread's stack of stores (innermost last)."""
    return _tls.current


class ParamStore:
    def __init__(self, params, device="cuda"):
        self.P = params
        self.device = torch.device(device)
        self._cache = {}
        # scratch buffers are per name space: every engine that may run concurrently with another one on a
        # different stream sets its own (engine.PWCLOEngine does) so that their forwards do not share them
        self.scratch_ns = 0

    def stream(self, scopes, pad_hidden=None):
        """Packed weights of a layer chain in the format of the MLP engine currently selected (pad_hidden: see
        packing.folded_chain)."""
        from . import _lib
        tc = _lib.mlp_engine() == 1
        key = ("stream_tc" if tc else "stream", pad_hidden) + tuple(scopes)
        if key not in self._cache:
            pack = packing.pack_stream_tc if tc else packing.pack_stream
            self._cache[key] = pack(self.P, scopes, pad_hidden).to(self.device)
        return self._cache[key]

    def plain(self, scopes):
        key = ("plain",) + tuple(scopes)
        if key not in self._cache:
            self._cache[key] = packing.pack_plain(self.P, scopes).to(self.device)
        return self._cache[key]

    def tensor(self, name):
        key = ("raw", name)
        if key not in self._cache:
            self._cache[key] = self.P[name].float().contiguous().to(self.device)
        return self._cache[key]

    def widths(self, scopes):
        return [int(self.P[s + "/weights"].shape[1]) for s in scopes]

    def cin(self, scope):
        return int(self.P[scope + "/weights"].shape[0])

    def scratch(self, name, shape, dtype, fill=None):
        """Persistent scratch buffers (projection cell minima, pose-head partials, counters)."""
        key = ("scratch", self.scratch_ns, name, tuple(shape), dtype)
        if key not in self._cache:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            if fill is not None:
                t.fill_(fill)
            self._cache[key] = t
        return self._cache[key]

    def eye(self, batch):
        """(batch, 4, 4) identity matrices on the device (constant)."""
        key = ("eye", batch)
        if key not in self._cache:
            self._cache[key] = torch.eye(4, device=self.device).expand(batch, 4, 4).contiguous()
        return self._cache[key]

    def default_aug(self, batch):
        """Constant (T (2B,4,4) identity, apply (2B,) = [0]*B + [1]*B): no augmentation matrix given, frame 2 is
        the 'augmented' one (the reference's inference setting, main.py:311-312)."""
        key = ("default_aug", batch)
        if key not in self._cache:
            apply = torch.cat([torch.zeros(batch, dtype=torch.int32), torch.ones(batch, dtype=torch.int32)])
            self._cache[key] = (self.eye(2 * batch), apply.to(self.device))
        return self._cache[key]

    def invalidate(self):
        self._cache.clear()


@contextlib.contextmanager
def use_store(store):
    _tls.current.append(store)
    try:
        yield store
    finally:
        _tls.current.pop()


def current_store(explicit=None):
    if explicit is not None:
        return explicit
    if not _tls.current:
        raise RuntimeError("no parameter store: wrap the call in `with use_store(ParamStore(params)):` "
                           "or pass params=")
    return _tls.current[-1]


@contextlib.contextmanager
def variable_scope(name):
    _tls.prefix.append(name)
    try:
        yield
    finally:
        _tls.prefix.pop()


def scoped(name):
    return "/".join(_tls.prefix + [name])
