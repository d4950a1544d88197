"""ORACLE SHIM: roma.internal batch-dimension helpers (roma 1.2.7)."""
import torch


def flatten_batch_dims(tensor, end_dim):
    batch_shape = tensor.shape[: end_dim + 1]
    flattened = tensor.flatten(end_dim=end_dim) if len(batch_shape) > 0 else tensor.unsqueeze(0)
    return flattened, batch_shape


def unflatten_batch_dims(tensor, batch_shape):
    return tensor.reshape(batch_shape + tensor.shape[1:]) if len(batch_shape) > 0 else tensor.squeeze(0)
