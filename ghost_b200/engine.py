"""Thin Python handle on a device plan (ctypes over include/ghost_cwt.h).

PyTorch supplies device buffers and streams only; all arithmetic happens in
libghostcwt.so.  A plan is built from the host planner's per-scale tables and is
independent of the recording length.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ["CwtPlan", "scale_tables", "OUT_KINDS"]

OUT_KINDS = {"complex": _lib.OUT_COMPLEX, "amplitude": _lib.OUT_AMPLITUDE, "power": _lib.OUT_POWER}


def scale_tables(wavelet, norm_radian_freqs, lengths):
    """Per-scale (k_first, n_terms, concatenated X[k]) for the C plan descriptor."""
    k_first, n_terms, terms = [], [], []
    for w, L in zip(np.atleast_1d(norm_radian_freqs), np.atleast_1d(lengths)):
        k0, X = wavelet.spectrum_terms(int(L), float(w))
        k_first.append(k0)
        n_terms.append(len(X))
        terms.append(X)
    return (np.asarray(k_first, dtype=np.int32), np.asarray(n_terms, dtype=np.int32),
            np.ascontiguousarray(np.concatenate(terms), dtype=np.float64))


def _torch():
    import torch
    return torch


class CwtPlan:
    """Device tables for one set of scales."""

    def __init__(self, lengths, k_first, n_terms, terms, *, dtype=np.float32, output="amplitude",
                 device=0, force_generic=False, no_interp=False, band_tol=0.0, guard=True, guard_tol=0.0):
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64 but got {}".format(dtype))
        if output not in OUT_KINDS:
            raise ValueError("output must be 'amplitude', 'power' or 'complex' but got {}".format(output))
        self.lib = _lib.load()
        self.dtype = dtype
        self.output = output
        self.device = int(device)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        self.n_scales = int(self.lengths.size)
        self._k_first = np.ascontiguousarray(k_first, dtype=np.int32)
        self._n_terms = np.ascontiguousarray(n_terms, dtype=np.int32)
        self._terms = np.ascontiguousarray(terms, dtype=np.float64)
        desc = _lib.PlanDesc()
        desc.n_scales = self.n_scales
        desc.lengths = self.lengths.ctypes.data_as(C.POINTER(C.c_int64))
        desc.k_first = self._k_first.ctypes.data_as(C.POINTER(C.c_int32))
        desc.n_terms = self._n_terms.ctypes.data_as(C.POINTER(C.c_int32))
        desc.terms = self._terms.ctypes.data_as(C.POINTER(C.c_double))
        desc.compute_type = _lib.F32 if dtype == np.float32 else _lib.F64
        desc.out_kind = OUT_KINDS[output]
        desc.device = self.device
        desc.flags = (_lib.FLAG_FORCE_GENERIC if force_generic else 0) | (_lib.FLAG_NO_INTERP if no_interp else 0) \
            | (0 if guard else _lib.FLAG_NO_GUARD)
        desc.band_tol = float(band_tol)
        desc.guard_tol = float(guard_tol)
        handle = C.c_void_p()
        _lib.check(self.lib.gcwt_plan_create(C.byref(handle), C.byref(desc)))
        self._h = handle

    # ------------------------------------------------------------------ info
    @property
    def max_length(self):
        return int(self.lengths.max())

    def levels(self):
        out = np.empty(self.n_scales, dtype=np.int32)
        _lib.check(self.lib.gcwt_plan_levels(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def guard_stats(self):
        """Accuracy guard of the fp32 paths (include/ghost_cwt.h, gcwt_guard_stats): dict with the
        (channel, scale) pairs re-computed in fp64 by the last execute, in total, the pairs examined
        so far, and a per-scale flag array of the last execute."""
        last, total, checked = C.c_int64(), C.c_int64(), C.c_int64()
        flags = np.zeros(self.n_scales, dtype=np.uint8)
        _lib.check(self.lib.gcwt_guard_stats(self._h, C.byref(last), C.byref(total), C.byref(checked),
                                             flags.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return {"last": int(last.value), "total": int(total.value), "checked": int(checked.value),
                "scales": flags.astype(bool)}

    def workspace_bytes(self):
        return int(self.lib.gcwt_plan_workspace_bytes(self._h))

    @property
    def torch_out_dtype(self):
        torch = _torch()
        if self.output == "complex":
            return torch.complex64 if self.dtype == np.float32 else torch.complex128
        return torch.float32 if self.dtype == np.float32 else torch.float64

    # ------------------------------------------------------------------ run
    def alloc_out(self, n_channels, n_samples):
        torch = _torch()
        return torch.empty((n_channels, self.n_scales, n_samples), dtype=self.torch_out_dtype,
                           device="cuda:%d" % self.device)

    def channel_means(self, x, n_samples=None):
        """float64 mean per channel of a (C, N) device tensor (reference transforms.py:143)."""
        torch = _torch()
        n = x.shape[1] if n_samples is None else int(n_samples)
        means = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
        in_type = _lib.F32 if x.dtype == torch.float32 else _lib.F64
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.gcwt_channel_means(x.data_ptr(), in_type, x.shape[0], n, x.stride(0),
                                               means.data_ptr(), self.device, st))
        return means

    def execute(self, x, out=None, *, means=None, start=0, stop=None, halo_left=0, halo_right=0,
                out_start=None):
        """Transform samples ``[start, stop)`` of every channel of ``x`` (C, N) into
        ``out[:, :, out_start : out_start + stop - start]`` (``out_start`` defaults to
        ``start``).  ``x`` and ``out`` are CUDA tensors; ``out`` is (C, S, >= needed)."""
        torch = _torch()
        if x.dim() != 2 or not x.is_cuda or x.stride(1) != 1:
            raise ValueError("x must be a (channels, samples) CUDA tensor with unit sample stride")
        if x.dtype not in (torch.float32, torch.float64):
            raise ValueError("x must be float32 or float64")
        n_ch, n_all = x.shape
        stop = n_all if stop is None else int(stop)
        start = int(start)
        if not (0 <= start < stop <= n_all):
            raise ValueError("bad segment [{}, {})".format(start, stop))
        if halo_left > start or halo_right > n_all - stop:
            raise ValueError("halo reaches outside the tensor")
        out_start = start if out_start is None else int(out_start)
        if out is None:
            out = self.alloc_out(n_ch, out_start + stop - start)
        if out.dtype != self.torch_out_dtype or out.dim() != 3 or out.stride(2) != 1 \
                or out.shape[0] != n_ch or out.shape[1] != self.n_scales \
                or out_start < 0 or out.shape[2] < out_start + stop - start:
            raise ValueError("out must be (channels, scales, samples) of dtype %s" % self.torch_out_dtype)
        esz = x.element_size()
        osz = out.element_size()
        mptr = None
        if means is not None:
            if means.dtype != torch.float64 or means.numel() != n_ch or not means.is_cuda:
                raise ValueError("means must be a float64 CUDA tensor with one entry per channel")
            mptr = means.data_ptr()
        st = torch.cuda.current_stream(x.device).cuda_stream
        in_type = _lib.F32 if x.dtype == torch.float32 else _lib.F64
        _lib.check(self.lib.gcwt_execute(
            self._h, x.data_ptr() + start * esz, in_type, n_ch, stop - start, x.stride(0),
            int(halo_left), int(halo_right), mptr,
            out.data_ptr() + out_start * osz, out.stride(1), out.stride(0), st))
        return out

    def execute_tiled(self, x, tile, out=None, consumer=None, means=None):
        """Transform a recording whose result does not fit in device memory.

        The samples are cut into tiles of ``tile`` samples; each tile is transformed with
        real-sample halos on both sides (exact: the filters are FIR) into
        the same reused ``out`` buffer (C, S, tile) and handed to ``consumer(out, start,
        stop)`` -- which must be done with the buffer when it returns (same stream).
        Returns the number of coefficients produced."""
        torch = _torch()
        n_ch, n = x.shape
        tile = int(min(tile, n))
        if out is None:
            out = self.alloc_out(n_ch, tile)
        if means is None:
            means = self.channel_means(x)
        halo = self.max_length - 1
        done = 0
        for a in range(0, n, tile):
            b = min(n, a + tile)
            self.execute(x, out, means=means, start=a, stop=b, halo_left=min(halo, a),
                         halo_right=min(halo, n - b), out_start=0)
            if consumer is not None:
                consumer(out, a, b)
            done += (b - a) * n_ch * self.n_scales
        return done

    def execute_host(self, x, out=None, means=None, epochs=None):
        """Host-buffer entry point (numpy in, numpy out) through gcwt_execute_host: the whole transform
        body of the reference (global mean removal, one zero-padded convolution per epoch, zeros outside
        the epochs; transforms.py:142-143,185,202-204) as one streamed, double-buffered call.

        ``x`` (channels, samples) or (samples,); ``out`` optional (channels, scales, samples) array of the
        plan's output dtype with unit sample stride -- pass a pinned one (e.g. the numpy view of a
        ``torch.empty(..., pin_memory=True)``) to let the device write it by DMA; ``epochs`` optional
        (E, 2) array of [start, stop) sample bounds.  Results larger than device memory are fine."""
        x = np.asarray(x)
        if x.ndim == 1:
            x = x[None, :]
        if x.ndim != 2:
            raise ValueError("x must be (channels, samples)")
        if x.dtype not in (np.float32, np.float64) or (self.dtype == np.float64 and x.dtype != np.float64):
            x = x.astype(np.float64)
        if x.strides[1] != x.itemsize:
            x = np.ascontiguousarray(x)
        n_ch, n = x.shape
        if self.output == "complex":
            odt = np.dtype(np.complex64 if self.dtype == np.float32 else np.complex128)
        else:
            odt = self.dtype
        if out is None:
            out = np.empty((n_ch, self.n_scales, n), dtype=odt)
        if out.dtype != odt or out.shape != (n_ch, self.n_scales, n) or out.strides[2] != out.itemsize \
                or out.strides[0] % out.itemsize or out.strides[1] % out.itemsize:
            raise ValueError("out must be (channels, scales, samples) of dtype %s with unit sample stride" % odt)
        mptr = None
        if means is not None:
            means = np.ascontiguousarray(means, dtype=np.float64)
            if means.size != n_ch:
                raise ValueError("means must hold one value per channel")
            mptr = means.ctypes.data
        eptr, n_ep = None, 0
        if epochs is not None:
            epochs = np.ascontiguousarray(epochs, dtype=np.int64).reshape(-1, 2)
            eptr, n_ep = epochs.ctypes.data, int(epochs.shape[0])
        in_type = _lib.F32 if x.dtype == np.float32 else _lib.F64
        _lib.check(self.lib.gcwt_execute_host(self._h, x.ctypes.data, in_type, n_ch, n, x.strides[0] // x.itemsize,
                                              eptr, n_ep, mptr, out.ctypes.data, out.strides[1] // out.itemsize,
                                              out.strides[0] // out.itemsize))
        return out

    def execute_host_pooled(self, x, pool_width, pool="mean", means=None):
        """As :meth:`execute_host` (one epoch, amplitude or power plans), but every tile is pooled on the
        device: runs of ``pool_width`` consecutive samples are reduced to their mean (``pool='mean'``) or
        maximum (``'max'``) and only those float64 bins cross the link.  Returns (channels, scales,
        ceil(samples / pool_width)) float64 -- what a display of a few thousand columns needs
        (reference plot: transforms.py:356-367,395-396)."""
        if pool not in ("mean", "max"):
            raise ValueError("pool must be 'mean' or 'max' but got {}".format(pool))
        pool_width = int(pool_width)
        if pool_width < 1:
            raise ValueError("pool_width must be a positive integer")
        x = np.asarray(x)
        if x.ndim == 1:
            x = x[None, :]
        if x.dtype not in (np.float32, np.float64) or (self.dtype == np.float64 and x.dtype != np.float64):
            x = x.astype(np.float64)
        if x.strides[1] != x.itemsize:
            x = np.ascontiguousarray(x)
        n_ch, n = x.shape
        n_bins = -(-n // pool_width)
        out = np.empty((n_ch, self.n_scales, n_bins), dtype=np.float64)
        mptr = None
        if means is not None:
            means = np.ascontiguousarray(means, dtype=np.float64)
            mptr = means.ctypes.data
        in_type = _lib.F32 if x.dtype == np.float32 else _lib.F64
        _lib.check(self.lib.gcwt_execute_host_pooled(
            self._h, x.ctypes.data, in_type, n_ch, n, x.strides[0] // x.itemsize, mptr, pool_width,
            _lib.POOL_MEAN if pool == "mean" else _lib.POOL_MAX, out.ctypes.data, n_bins, n_bins * self.n_scales))
        return out

    def pool_rows(self, res, pool_width, pool="mean", square=False):
        """Pool the last axis of a contiguous CUDA tensor (a device-resident result) in runs of ``pool_width``
        samples: float64 CUDA tensor with ceil(n / pool_width) bins per row.  ``square`` pools the squares."""
        torch = _torch()
        if not res.is_cuda or res.is_complex() or res.dtype not in (torch.float32, torch.float64) or res.stride(-1) != 1:
            raise ValueError("pool_rows needs a real float32 / float64 CUDA tensor with unit stride along the last axis")
        if pool not in ("mean", "max"):
            raise ValueError("pool must be 'mean' or 'max' but got {}".format(pool))
        flat = res.reshape(-1, res.shape[-1]) if res.is_contiguous() else None
        if flat is None:
            if res.dim() != 2:
                raise ValueError("pool_rows needs a contiguous tensor (or a 2-D one with a row stride)")
            flat = res
        n_rows, n = flat.shape
        n_bins = -(-n // int(pool_width))
        out = torch.empty((n_rows, n_bins), dtype=torch.float64, device=res.device)
        st = torch.cuda.current_stream(res.device).cuda_stream
        _lib.check(self.lib.gcwt_pool_rows(flat.data_ptr(), _lib.F32 if res.dtype == torch.float32 else _lib.F64, n_rows, n,
                                           flat.stride(0), int(pool_width), _lib.POOL_MEAN if pool == "mean" else _lib.POOL_MAX,
                                           1 if square else 0, out.data_ptr(), n_bins, self.device, st))
        return out.reshape(tuple(res.shape[:-1]) + (n_bins,))

    def host_stats(self):
        """{wall_ms, pinned_destination, tiles, bytes_out} of the last execute_host."""
        v = (C.c_double * 4)()
        _lib.check(self.lib.gcwt_host_stats(self._h, v))
        return {"wall_ms": float(v[0]), "pinned_destination": bool(v[1]), "tiles": int(v[2]), "bytes_out": int(v[3])}

    PROFILE_KINDS = ("mean+pyramid", "fused_full", "fused_banded", "generic", "fused_interp")

    def profile(self, on=True):
        _lib.check(self.lib.gcwt_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, reset=True):
        """{family: (milliseconds, launches)} accumulated since the last reset."""
        ms = (C.c_double * 5)()
        ln = (C.c_int64 * 5)()
        _lib.check(self.lib.gcwt_profile_read(self._h, ms, ln, 1 if reset else 0))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.gcwt_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
