"""List / padded / packed conversions for ragged batches (reference: structures/utils.py).

Same function names, arguments and results as the reference.  On CUDA float32 data the
list->padded path runs as ONE ragged-copy kernel over the concatenated list
(`_C.packed_to_padded`) instead of a Python loop of N slice assignments (utils.py:73-79), and
the index helpers are vectorised (no per-cloud `arange` loop, utils.py:232-240).
"""
from typing import List, Sequence, Tuple, Union

import torch


def _first_idx_from_sizes(sizes: torch.Tensor) -> torch.Tensor:
    first = torch.zeros_like(sizes)
    if sizes.numel() > 1:
        first[1:] = torch.cumsum(sizes[:-1], dim=0)
    return first


def _cuda_ragged_ok(x: Sequence[torch.Tensor], pad_value: float) -> bool:
    return (
        len(x) > 0
        and pad_value == 0.0
        and all(y.is_cuda and y.dtype == torch.float32 and y.dim() == 2 for y in x)
        and len({y.shape[1] for y in x}) == 1
        and not any(y.requires_grad for y in x)
    )


def list_to_padded(
    x: Union[List[torch.Tensor], Tuple[torch.Tensor]],
    pad_size: Union[Sequence[int], None] = None,
    pad_value: float = 0.0,
    equisized: bool = False,
) -> torch.Tensor:
    """List of N tensors (Si_0, ..., Si_D) -> one tensor (N, pad_size...) (or the per-dim maxima
    when pad_size is None), filled with pad_value outside each item (reference :19-79)."""
    if equisized:
        return torch.stack(x, 0)
    if not all(torch.is_tensor(y) for y in x):
        raise ValueError("All items have to be instances of a torch.Tensor.")
    ndim = max(y.ndim for y in x)
    x = [y.new_zeros([0] * ndim) if (y.ndim == 1 and y.nelement() == 0) else y for y in x]
    if any(y.ndim != x[0].ndim for y in x):
        raise ValueError("All items have to have the same number of dimensions!")
    if pad_size is None:
        pad_dims = [max(y.shape[d] for y in x if len(y) > 0) for d in range(x[0].ndim)]
    else:
        if any(len(pad_size) != y.ndim for y in x):
            raise ValueError("Pad size must contain target size for all dimensions.")
        pad_dims = list(pad_size)

    if _cuda_ragged_ok(x, pad_value) and x[0].shape[1] == pad_dims[1] and all(
        y.shape[0] <= pad_dims[0] for y in x
    ):
        from .. import _C  # CUDA only; importing lazily keeps this module usable on CPU

        sizes = torch.tensor([y.shape[0] for y in x], dtype=torch.int64)
        first = _first_idx_from_sizes(sizes).to(x[0].device, non_blocking=True)
        return _C.packed_to_padded(torch.cat(list(x), dim=0), first, int(pad_dims[0]))

    out = x[0].new_full((len(x), *pad_dims), pad_value)
    for i, y in enumerate(x):
        if len(y) > 0:
            out[(i, *(slice(0, s) for s in y.shape))] = y
    return out


def padded_to_list(
    x: torch.Tensor,
    split_size: Union[Sequence[int], Sequence[Sequence[int]], None] = None,
):
    """Padded (N, S_1, ..., S_D) -> list of N views, optionally cropped to split_size[i]
    (an int crops dim 0, a sequence crops every dim) (reference :82-116)."""
    items = list(x.unbind(0))
    if split_size is None:
        return items
    if x.shape[0] != len(split_size):
        raise ValueError("Split size must be of same length as inputs first dimension")
    for i, s in enumerate(split_size):
        items[i] = items[i][:s] if isinstance(s, int) else items[i][tuple(slice(0, e) for e in s)]
    return items


def list_to_packed(x: List[torch.Tensor]):
    """List of N tensors (Mi, K, ...) -> (x_packed (sum Mi, K, ...), num_items (N,),
    item_packed_first_idx (N,), item_packed_to_list_idx (sum Mi,)) (reference :119-154)."""
    if not x:
        raise ValueError("Input list is empty")
    device = x[0].device
    sizes = [xi.shape[0] for xi in x]
    num_items = torch.tensor(sizes, dtype=torch.int64, device=device)
    first = _first_idx_from_sizes(num_items)
    to_list = torch.repeat_interleave(
        torch.arange(len(sizes), dtype=torch.int64, device=device), num_items,
        output_size=sum(sizes))
    return torch.cat(x, dim=0), num_items, first, to_list


def packed_to_list(x: torch.Tensor, split_size: Union[list, int]):
    """Packed (sum Mi, K, ...) -> tuple of tensors (Mi, K, ...) (reference :157-170)."""
    return x.split(split_size, dim=0)


def padded_to_packed(
    x: torch.Tensor,
    split_size: Union[list, tuple, None] = None,
    pad_value: Union[float, int, None] = None,
):
    """Padded (N, M, K) -> packed (sum Mi, K) using split_size, or dropping rows equal to
    pad_value, or (N*M, K) when neither is given (reference :173-242)."""
    if x.ndim != 3:
        raise ValueError("Supports only 3-dimensional input tensors")
    N, M, D = x.shape
    if split_size is not None and pad_value is not None:
        raise ValueError("Only one of split_size or pad_value should be provided.")
    flat = x.reshape(-1, D)
    if pad_value is None and split_size is None:
        return flat
    if pad_value is not None:
        return flat[flat.ne(pad_value).any(-1)]
    if x.shape[0] != len(split_size):
        raise ValueError("Split size must be of same length as inputs first dimension")
    if not all(isinstance(i, int) for i in split_size):
        raise ValueError(
            "Support only 1-dimensional unbinded tensor. \
                Split size for more dimensions provided"
        )
    if x.is_cuda and x.dtype == torch.float32 and not x.requires_grad:
        from .. import _C

        sizes = torch.tensor(list(split_size), dtype=torch.int64)
        first = _first_idx_from_sizes(sizes).to(x.device, non_blocking=True)
        return _C.padded_to_packed(x.contiguous(), first, int(sum(split_size)))
    return flat[padded_to_packed_index(torch.tensor(list(split_size), dtype=torch.int64, device=x.device), M)]


def padded_to_packed_index(num_items: torch.Tensor, max_size: int) -> torch.Tensor:
    """Flat indices into a (N*max_size) padded layout of every valid item, in packed order:
    item i of cloud b -> b*max_size + i.  Vectorised (no per-cloud loop)."""
    total = int(num_items.sum()) if num_items.numel() else 0
    cloud = torch.repeat_interleave(
        torch.arange(num_items.numel(), dtype=torch.int64, device=num_items.device), num_items,
        output_size=total)
    first = _first_idx_from_sizes(num_items)
    within = torch.arange(total, dtype=torch.int64, device=num_items.device) - first[cloud]
    return cloud * max_size + within
