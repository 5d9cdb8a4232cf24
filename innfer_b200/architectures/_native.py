"""Shared plumbing between the nn.Module mirrors and their native engines.

A mirror module owns the parameters under the reference's state-dict names; on a CUDA tensor its
``forward`` runs the sm_100a engine, which holds its own repacked copy of the weights.  The copy
is keyed on a fingerprint of the parameters (storage pointer and in-place version counter of each
one) so that ``param.data.copy_()``, optimizer steps or ``load_state_dict`` after the first CUDA
forward rebuild the engine instead of silently using stale weights, and it is left out of
pickling / deep copies (it holds ctypes handles).
"""
import copy

import torch


class NativeEngineMixin:
    """Mix into an ``nn.Module``; set ``_engine_class`` to the name of a class in innfer_b200.engine."""

    _engine_class = "RRDBEngine"

    def _engine_key_extra(self):
        """Module state other than the parameters that the engine bakes in (e.g. train/eval mode of norm layers)."""
        return ()

    def _fingerprint(self):
        fp = tuple((p.data_ptr(), p._version) for p in self.parameters())
        return fp + tuple((b.data_ptr(), b._version) for b in self.buffers()) + tuple(self._engine_key_extra())

    def _engine(self, device, dtype, **variant):
        """``variant``: extra keyword arguments of the engine class's ``from_module`` (e.g. ``unit_io=True`` for the
        generators that run.py wraps in a [-1, 1] normalisation)."""
        from .. import engine as E
        cache = self.__dict__.setdefault("_engines", {})
        key = (str(device), dtype) + tuple(sorted(variant.items()))
        fp = self._fingerprint()
        hit = cache.get(key)
        if hit is not None and hit[0] == fp:
            return hit[1]
        for _, old in cache.values():   # one resident engine per module
            old.close()
        cache.clear()
        eng = getattr(E, self._engine_class).from_module(self, device, fp16=(dtype == torch.float16), **variant)
        cache[key] = (fp, eng)
        return eng

    def invalidate_engine(self):
        """Drop the native engine (it is rebuilt from the current parameters on the next CUDA forward)."""
        for _, old in self.__dict__.get("_engines", {}).values():
            old.close()
        self.__dict__["_engines"] = {}

    def native_engine(self, device, dtype=torch.float16, **variant):
        """The engine serving CUDA tensors of this dtype (built on first use)."""
        return self._engine(torch.device(device), dtype, **variant)

    def chop_forward_native(self, x, patch_size, step):
        """extract_patches_2d -> forward -> recompose_tensor in one native call (CUDA only)."""
        return self._engine(x.device, x.dtype).chop_forward(x, patch_size, step)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engines"] = {}
        return state

    def __deepcopy__(self, memo):
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_engines" else copy.deepcopy(v, memo)
        return new
