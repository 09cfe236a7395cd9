"""`_device` and `to_device` with the contract of the reference's utils/torch_helpers.py:7-23: tensors anywhere inside
nested dicts / lists are moved to `_device` (asynchronously), containers are updated IN PLACE and returned, plain
scalars / strings / None pass through."""
import torch

_device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
_PASS_THROUGH = (int, float, str, type(None))


def to_device(data):
    if torch.is_tensor(data):
        return data.to(_device, non_blocking=True)
    if isinstance(data, _PASS_THROUGH):
        return data
    keys = data.keys() if isinstance(data, dict) else range(len(data))
    for k in list(keys):
        data[k] = to_device(data[k])          # re-bind the slot: callers rely on the container being mutated
    return data
