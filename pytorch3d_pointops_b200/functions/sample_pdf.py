"""sample_pdf -- inverse-CDF sampling of piecewise-constant densities along rays.

Mirrors functions/sample_pdf.py of the reference (sample_pdf :14-66, sample_pdf_python :69-148):
same arguments, defaults, errors and return shape.  `sample_pdf` draws (or, with `det`, lays out)
uniform numbers with torch and lets one kernel turn them into samples in place
(`_C.sample_pdf`, csrc/sample_pdf.cu); `sample_pdf_python` is the searchsorted formulation in
plain torch ops, kept as the secondary, differentiably-structured variant the reference ships.
"""
import torch

from .. import _C


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, det: bool = False,
               eps: float = 1e-5) -> torch.Tensor:
    """bins (..., n_bins+1) edges, weights (..., n_bins) non-negative -> samples (..., n_samples).
    det=True: evenly spaced quantiles instead of random ones."""
    if torch.is_grad_enabled() and (bins.requires_grad or weights.requires_grad):
        raise NotImplementedError("sample_pdf differentiability.")
    if weights.min() <= -eps:
        raise ValueError("Negative weights provided.")
    batch_shape = bins.shape[:-1]
    n_bins = weights.shape[-1]
    if n_bins + 1 != bins.shape[-1] or weights.shape[:-1] != batch_shape:
        raise ValueError("Inconsistent shapes of bins and weights: " + f"{bins.shape}{weights.shape}")
    out_shape = batch_shape + (n_samples,)
    if det:
        u = torch.linspace(0.0, 1.0, n_samples, device=bins.device, dtype=torch.float32)
        out = u.expand(out_shape).contiguous()
    else:
        out = torch.rand(out_shape, dtype=torch.float32, device=bins.device)
    _C.sample_pdf(bins.reshape(-1, n_bins + 1), weights.reshape(-1, n_bins), out.view(-1, n_samples), eps)
    return out


def sample_pdf_python(bins: torch.Tensor, weights: torch.Tensor, N_samples: int, det: bool = False,
                      eps: float = 1e-5) -> torch.Tensor:
    """Same sampling with torch ops only: O(n_bins + n_samples log n_bins) per row."""
    weights = weights + eps  # keeps empty bins from producing NaNs
    if weights.min() <= 0:
        raise ValueError("Negative weights provided.")
    pdf = weights / weights.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    shape = list(cdf.shape[:-1]) + [N_samples]
    if det:
        u = torch.linspace(0.0, 1.0, N_samples, device=cdf.device, dtype=cdf.dtype).expand(shape).contiguous()
    else:
        u = torch.rand(shape, device=cdf.device, dtype=cdf.dtype)
    hi_idx = torch.searchsorted(cdf, u, right=True)
    lo_idx = (hi_idx - 1).clamp(0)
    hi_idx = hi_idx.clamp(max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, -1, lo_idx), torch.gather(cdf, -1, hi_idx)
    bin_lo, bin_hi = torch.gather(bins, -1, lo_idx), torch.gather(bins, -1, hi_idx)
    width = cdf_hi - cdf_lo
    width = torch.where(width < eps, torch.ones_like(width), width)
    return bin_lo + (u - cdf_lo) / width * (bin_hi - bin_lo)
