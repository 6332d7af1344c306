"""Index helpers with the names of reference src/utils/tensor.py / src/utils/sparse.py that
the data containers use."""
import numpy as np
import torch

__all__ = ['tensor_idx', 'is_arange', 'sizes_to_pointers', 'indices_to_pointers']


def tensor_idx(idx, device=None):
    """int / slice / numpy / bool-mask / tensor index -> 1-D LongTensor on `device`
    (reference src/utils/tensor.py:33-70; None stays None; like the reference, a Python list
    is rejected)."""
    if not isinstance(idx, (int, slice, np.ndarray, torch.Tensor, type(None))):
        raise ValueError(
            f"Expected an int, slice, list, np.ndarray, or torch.Tensor index, but received a "
            f"{type(idx)} instead.")
    if idx is None:
        return None
    if device is None:
        device = idx.device if hasattr(idx, 'device') else 'cpu'
    if isinstance(idx, int):
        idx = torch.tensor([idx], device=device, dtype=torch.long)
    elif isinstance(idx, slice):
        idx = torch.arange(idx.start, idx.stop, device=device)
    elif isinstance(idx, np.ndarray):
        idx = torch.from_numpy(idx).to(device)
    else:
        idx = idx.to(device)
    if idx.dtype == torch.bool:
        idx = torch.where(idx)[0]
    return idx.long()


def is_arange(a, n):
    """a == arange(n)  (reference src/utils/tensor.py:73-79)"""
    if a.dim() != 1 or a.shape[0] != n:
        return False
    return n == 0 or bool((a == torch.arange(n, device=a.device)).all())


def sizes_to_pointers(sizes):
    """reference src/utils/sparse.py:44-50"""
    assert sizes.dim() == 1
    zero = torch.zeros(1, device=sizes.device, dtype=torch.long)
    return torch.cat((zero, sizes.long())).cumsum(dim=0)


def indices_to_pointers(indices, num_groups=None):
    """Dense group ids -> (CSR pointers, order) (reference src/utils/sparse.py:23-41).  The
    reference sorts with torch.sort (order inside a group unspecified); here the order is the
    STABLE one.  CUDA tensors go through spt_group_index; `num_groups` saves the max() sync."""
    assert indices.dim() == 1, "Only 1D indices are accepted."
    assert indices.shape[0] >= 1, "At least one group index is required."
    if num_groups is None:
        num_groups = int(indices.max()) + 1
    if indices.is_cuda:
        from .. import ops
        seg = ops.segment_index(indices, num_groups)
        return seg.ptr.long(), seg.perm.long()
    order = torch.sort(indices, stable=True).indices
    pointers = torch.zeros(num_groups + 1, dtype=torch.long)
    pointers[1:] = torch.bincount(indices, minlength=num_groups).cumsum(0)
    return pointers, order
