"""Mirror of the slice of NVIDIA pytorch_quantization that quant/*.py uses (QuantDescriptor, TensorQuantizer,
calib.MaxCalibrator): symmetric, narrow-range, round-half-even fake quantisation with a dynamic or calibrated amax.
Reference call sites: quant/quant.py:1-2,14-32; quant/quantize.py:3-7,175-207.  Semantics: SURVEY.md 8a-Q.

The sparse-conv hot path does NOT run `TensorQuantizer.forward`: QConvNd reads the descriptor (bits, axis, _amax)
and dispatches the fused sm_100a kernels.  `forward` is the torch restatement used for weights at wrap time, for
the dense SQ* layers, and when a user calls the quantizer directly."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class QuantDescriptor:
    def __init__(self, num_bits: int = 8, name=None, fake_quant: bool = True, axis=None, amax=None, learn_amax=False,
                 scale_amax=None, calib_method: str = "max", unsigned: bool = False, narrow_range: bool = True):
        if unsigned:
            raise NotImplementedError("unsigned quantisation is commented out at every reference call site (quant/quant.py:24,30)")
        self.num_bits = int(num_bits)
        self.fake_quant = fake_quant
        if axis is not None and not isinstance(axis, (tuple, list)):
            axis = (int(axis),)
        self.axis = None if axis is None else tuple(int(a) for a in axis)
        self.amax = amax
        self.calib_method = calib_method
        self.narrow_range = narrow_range
        self.unsigned = unsigned


class MaxCalibrator:
    """Running max of |x| over the reduction axes ([EXT] calib.MaxCalibrator)."""

    def __init__(self, num_bits, axis, unsigned=False):
        self._axis = axis
        self._calib_amax = None

    def collect(self, x: torch.Tensor):
        a = reduce_amax(x, self._axis)
        self._calib_amax = a if self._calib_amax is None else torch.maximum(self._calib_amax, a)

    def compute_amax(self):
        return self._calib_amax

    def reset(self):
        self._calib_amax = None


class HistogramCalibrator:
    """[EXT] pytorch_quantization calib.HistogramCalibrator as the reference configures it (per-tensor, `_torch_hist = True`:
    quant/quantize.py:138-145; QuantDescriptor(calib_method='histogram'): quant/count_time_n_memory.py:304-365) and
    compute_amax(method=...) as `compute_amax(model, method='entropy' | 'mse' | 'percentile', ...)` forwards it
    (quant/quantize.py:198-207).  The package is not in this image and not vendored by the reference: this restates its PUBLISHED
    algorithm (NVIDIA TensorRT pytorch-quantization 2.1, calib/histogram.py) -- parity UNPINNED, checked only against the oracle's
    independent numpy restatement and against closed-form cases (tests/test_oracle.py).

    collect(x): histogram of |x| with `num_bins` equal bins over [0, max|x|]; a later batch with a larger maximum extends the
    range with bins of the SAME width (old counts keep their bins).  Everything is torch ops on x's device (calibration is one-off
    host-side plumbing, not a hot path).
    compute_amax(method):
      percentile : smallest bin edge below which `percentile` % of the counts lie
      mse        : the candidate amax (bin edges from start_bin, every `stride`) minimising the count-weighted squared error of
                   fake-quantising the bin centres
      entropy    : the TensorRT KL-divergence search: for every candidate i (from start_bin, every `stride`) the reference
                   distribution is bins[:i] with the tail mass folded into its last bin, the candidate is that distribution
                   quantised to 2^(bits-1) levels (each level's mass spread evenly over its non-empty source bins); amax = the edge
                   with the smallest KL(reference || candidate)."""

    def __init__(self, num_bits=8, axis=None, unsigned=False, num_bins=2048, grow_method=None, skip_zeros=False, torch_hist=True):
        if axis is not None:
            raise NotImplementedError("the histogram calibrator is per-tensor ([EXT] raises for an axis, too)")
        self._num_bits, self._unsigned, self._num_bins, self._skip_zeros = int(num_bits), bool(unsigned), int(num_bins), bool(skip_zeros)
        self._torch_hist = True
        self._calib_hist = None
        self._calib_bin_edges = None

    def collect(self, x: torch.Tensor):
        x = x.detach().float().abs().reshape(-1)
        if self._skip_zeros:
            x = x[x != 0]
        if x.numel() == 0:
            return
        x_max = x.max()
        if self._calib_hist is None:
            top = float(x_max) if float(x_max) > 0 else 1.0
            self._calib_hist = torch.histc(x, bins=self._num_bins, min=0, max=top)
            self._calib_bin_edges = torch.linspace(0, top, self._num_bins + 1, device=x.device)
        else:
            edges = self._calib_bin_edges
            if x_max > edges[-1]:
                width = edges[1] - edges[0]
                n_bins = int(torch.ceil(x_max / width).item())
                edges = torch.arange(0, n_bins + 1, device=x.device, dtype=torch.float32) * width
                self._calib_bin_edges = edges
            hist = torch.histc(x, bins=edges.numel() - 1, min=0, max=float(edges[-1]))
            hist[:self._calib_hist.numel()] += self._calib_hist
            self._calib_hist = hist

    def reset(self):
        self._calib_hist = None
        self._calib_bin_edges = None

    def compute_amax(self, method: str = "entropy", *, stride: int = 1, start_bin: int = 128, percentile: float = 99.99):
        if self._calib_hist is None:
            return None
        hist = self._calib_hist.double().cpu()
        edges = self._calib_bin_edges.double().cpu()
        if method == "percentile":
            if not 0 <= percentile <= 100:
                raise ValueError("percentile must be in [0, 100]")
            cdf = torch.cumsum(hist / hist.sum(), 0)
            idx = int(torch.searchsorted(cdf, torch.tensor(percentile / 100.0, dtype=torch.float64)).item())
            return edges[min(idx, edges.numel() - 1)].float()
        if method == "mse":
            centers = ((edges[1:] + edges[:-1]) / 2).float()
            counts = hist.float()
            best, best_amax = None, None
            for i in range(start_bin, hist.numel() + 1, stride):
                amax = centers[i - 1] if i - 1 < centers.numel() else edges[-1].float()
                q = fake_quant(centers, amax, self._num_bits + int(self._unsigned))
                err = float((((q - centers) ** 2) * counts).mean())
                if best is None or err < best:
                    best, best_amax = err, amax
            return best_amax.float()
        if method == "entropy":
            return _entropy_amax(hist, edges, self._num_bits, self._unsigned, stride, start_bin).float()
        raise TypeError(f"unknown calibration method {method}")


def _entropy_amax(hist: torch.Tensor, edges: torch.Tensor, num_bits: int, unsigned: bool, stride: int, start_bin: int) -> torch.Tensor:
    """TensorRT entropy calibration on a histogram of |x| (float64 on the host)."""
    bins = hist.clone()
    bins[0] = bins[1]                                           # the zero bin is dominated by exact zeros (ReLU): TensorRT overwrites it
    total = float(bins.sum())
    nlevels = 1 << (num_bits - 1 + int(unsigned))
    best, best_i = None, None
    starting = max(int(start_bin), nlevels)
    for i in range(starting, bins.numel() + 1, max(int(stride), 1)):
        ref = bins[:i].clone()
        ref[i - 1] += bins[i:].sum()                            # clip: the tail's mass lands in the last kept bin
        # quantise the i bins to nlevels levels: level of source bin j = floor(j * nlevels / i)
        lvl = torch.div(torch.arange(i, dtype=torch.int64) * nlevels, i, rounding_mode="floor")
        src = bins[:i]
        level_mass = torch.zeros(nlevels, dtype=torch.float64).index_add_(0, lvl, src)
        nonzero = (src != 0).double()
        level_cnt = torch.zeros(nlevels, dtype=torch.float64).index_add_(0, lvl, nonzero)
        cand = torch.where(nonzero.bool(), level_mass[lvl] / level_cnt[lvl].clamp(min=1.0), torch.zeros(i, dtype=torch.float64))
        p = ref / ref.sum()
        q = cand / cand.sum() if float(cand.sum()) > 0 else cand
        m = p > 0
        if bool((q[m] == 0).any()):
            continue                                            # KL is infinite: the candidate lost mass the reference has
        kl = float((p[m] * torch.log(p[m] / q[m])).sum())
        if best is None or kl < best:
            best, best_i = kl, i
    if best_i is None:
        best_i = bins.numel()
    _ = total
    return edges[best_i]


def reduce_amax(x: torch.Tensor, axis) -> torch.Tensor:
    a = x.detach().abs()
    if axis is None:
        return a.max() if a.numel() else a.new_zeros(())
    keep = [ax % x.dim() for ax in axis]
    red = [d for d in range(x.dim()) if d not in keep]
    return a.amax(dim=red, keepdim=True) if red else a


def quant_scale(amax: torch.Tensor, bound: float) -> torch.Tensor:
    """scale = bound / amax (0 where amax <= 2^-24) as ONE correctly rounded fp32 division per element, like [EXT]
    pytorch_quantization (`max_bound / amax`, both tensors) and like the device (__fdiv_rn).  Python's `bound / tensor` is
    tensor.reciprocal() * bound -- two roundings, one ulp off in a quarter of the cases -- and CUDA's `tensor / scalar`
    multiplies by the reciprocal; an ulp in a scale flips int8 codes that sit on a rounding boundary."""
    amax = amax.to(torch.float32)
    tiny = amax <= (1.0 / (1 << 24))
    return torch.where(tiny, torch.zeros_like(amax), torch.full_like(amax, float(bound)) / torch.where(tiny, torch.ones_like(amax), amax))


def fake_quant(x: torch.Tensor, amax: torch.Tensor, num_bits: int) -> torch.Tensor:
    bound = float(2 ** (num_bits - 1) - 1)
    scale = quant_scale(amax, bound)
    xf = x.to(torch.float32)
    q = torch.round(xf * scale).clamp_(-bound, bound)
    out = torch.where(scale == 0, torch.zeros_like(q), q / torch.where(scale == 0, torch.ones_like(scale), scale))
    return out.to(x.dtype)


class TensorQuantizer(nn.Module):
    def __init__(self, quant_desc: Optional[QuantDescriptor] = None, disabled=False, if_quant=True, if_calib=False):
        super().__init__()
        quant_desc = quant_desc or QuantDescriptor()
        self._num_bits = quant_desc.num_bits
        self._axis = quant_desc.axis
        self._fake_quant = quant_desc.fake_quant
        self._narrow_range = quant_desc.narrow_range
        self._unsigned = quant_desc.unsigned
        self._disabled = disabled
        self._if_quant = if_quant
        self._if_calib = if_calib
        if quant_desc.amax is not None:
            self.register_buffer("_amax", torch.as_tensor(quant_desc.amax, dtype=torch.float32))
        if quant_desc.calib_method == "max":
            self._calibrator = MaxCalibrator(self._num_bits, self._axis)
        elif quant_desc.calib_method == "histogram":
            self._calibrator = HistogramCalibrator(self._num_bits, self._axis, self._unsigned)
        else:
            raise ValueError(f"unknown calib_method {quant_desc.calib_method!r} (max | histogram)")

    # --- descriptor ---
    @property
    def num_bits(self):
        return self._num_bits

    @property
    def axis(self):
        return self._axis

    @property
    def amax(self):
        return getattr(self, "_amax", None)

    @amax.setter
    def amax(self, value):
        if value is None:
            if hasattr(self, "_amax"):
                delattr(self, "_amax")
            return
        v = torch.as_tensor(value, dtype=torch.float32)
        if hasattr(self, "_amax"):
            self._amax = v.to(self._amax.device)
        else:
            self.register_buffer("_amax", v)

    # --- switches used by collect_stats (quant/quantize.py:177-194) ---
    def enable_calib(self):
        if self._calibrator is None:
            raise RuntimeError("calibrator was not created")
        self._if_calib = True

    def disable_calib(self):
        self._if_calib = False

    def enable_quant(self):
        self._if_quant = True

    def disable_quant(self):
        self._if_quant = False

    def enable(self):
        self._disabled = False

    def disable(self):
        self._disabled = True

    def load_calib_amax(self, *args, strict=True, **kwargs):
        """[EXT] TensorQuantizer.load_calib_amax: *args / **kwargs go to the calibrator's compute_amax (method=, percentile=, stride=,
        start_bin= for the histogram calibrator; quant/quantize.py:198-207 calls it with strict=False for MaxCalibrator)."""
        if self._calibrator is None:
            amax = None
        elif isinstance(self._calibrator, MaxCalibrator):
            amax = self._calibrator.compute_amax()
        else:
            amax = self._calibrator.compute_amax(*args, **kwargs)
        if amax is None:
            if strict:
                raise RuntimeError("calibrator returned None (no data was collected)")
            return
        self.amax = amax.detach().clone()

    # --- torch fake-quant (not the sparse hot path) ---
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._disabled:
            return x
        if self._if_calib:
            self._calibrator.collect(x)
        if not self._if_quant:
            return x
        amax = self.amax if self.amax is not None else reduce_amax(x, self._axis)
        return fake_quant(x, amax.to(x.device), self._num_bits)

    def extra_repr(self):
        return f"{self._num_bits} bit fake per-{'channel axis=' + str(self._axis) if self._axis else 'tensor'} amax={'dynamic' if self.amax is None else 'calibrated'}"
