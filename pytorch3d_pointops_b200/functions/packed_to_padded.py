"""packed_to_padded / padded_to_packed -- reference: functions/packed_to_padded.py."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _C


def _check(inputs, first_idxs, ndim, size_arg):
    if inputs.dim() != ndim:
        raise ValueError("input can only be %d-dimensional." % ndim)
    if first_idxs.dim() != 1:
        raise ValueError("first_idxs can only be 1-dimensional.")
    if inputs.dtype != torch.float32:
        raise ValueError("input has to be of type torch.float32.")
    if first_idxs.dtype != torch.int64:
        raise ValueError("first_idxs has to be of type torch.int64.")
    if not isinstance(size_arg, int):
        raise ValueError("max_size has to be int.")


class _PackedToPadded(Function):
    """(F,D) -> (N,max_size,D); backward is the inverse copy (reference :15-62)."""

    @staticmethod
    def forward(ctx, inputs, first_idxs, max_size):
        _check(inputs, first_idxs, 2, max_size)
        ctx.save_for_backward(first_idxs)
        ctx.num_inputs = int(inputs.shape[0])
        return _C.packed_to_padded(inputs.contiguous(), first_idxs.contiguous(), max_size)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (first_idxs,) = ctx.saved_tensors
        return _C.padded_to_packed(grad_output.contiguous(), first_idxs, ctx.num_inputs), None, None


class _PaddedToPacked(Function):
    """(N,max_size,D) -> (F,D); backward is the inverse copy (reference :106-151)."""

    @staticmethod
    def forward(ctx, inputs, first_idxs, num_inputs):
        _check(inputs, first_idxs, 3, num_inputs)
        ctx.save_for_backward(first_idxs)
        ctx.max_size = inputs.shape[1]
        return _C.padded_to_packed(inputs.contiguous(), first_idxs.contiguous(), num_inputs)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (first_idxs,) = ctx.saved_tensors
        return _C.packed_to_padded(grad_output.contiguous(), first_idxs, ctx.max_size), None, None


def packed_to_padded(inputs: torch.Tensor, first_idxs: torch.LongTensor, max_size: int) -> torch.Tensor:
    """Packed (F,) / (F, ...) -> padded (N, max_size) / (N, max_size, ...), zero padded; cloud i
    starts at inputs[first_idxs[i]] (reference :65-103)."""
    shape = inputs.shape
    flat = inputs.unsqueeze(1) if inputs.dim() == 1 else inputs.reshape(shape[0], -1)
    padded = _PackedToPadded.apply(flat, first_idxs, max_size)
    if inputs.dim() == 1:
        return padded.squeeze(2)
    if inputs.dim() == 2:
        return padded
    return padded.view(*padded.shape[:2], *shape[1:])


def padded_to_packed(
    inputs: torch.Tensor,
    first_idxs: torch.LongTensor,
    num_inputs: int,
    max_size_dim: int = 1,
) -> torch.Tensor:
    """Padded (N, ..., max_size, ...) -> packed (F,) / (F, ...) with the ragged dimension at
    `max_size_dim` (reference :154-198)."""
    n_dims = inputs.dim()
    inputs = inputs.movedim(max_size_dim, 1)
    shape = inputs.shape
    flat = inputs.unsqueeze(2) if n_dims == 2 else inputs.reshape(*shape[:2], -1)
    packed = _PaddedToPacked.apply(flat, first_idxs, num_inputs)
    if n_dims == 2:
        return packed.squeeze(1)
    return packed.view(-1, *shape[2:])
