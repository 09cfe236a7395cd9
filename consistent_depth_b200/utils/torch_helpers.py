"""Mirror of the reference's utils/torch_helpers.py:7-23 (`_device`, recursive `to_device`)."""
import torch

_device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def to_device(data):
    """Recursively move tensors of nested dict/list structures to `_device`, in place
    (same contract as utils/torch_helpers.py:10-23: dict values / list items are re-bound)."""
    if isinstance(data, torch.Tensor):
        return data.to(_device, non_blocking=True)
    if isinstance(data, dict):
        for k, v in data.items():
            data[k] = to_device(v)
        return data
    if isinstance(data, (int, float, str)) or data is None:
        return data
    for i, v in enumerate(data):
        data[i] = to_device(v)
    return data
