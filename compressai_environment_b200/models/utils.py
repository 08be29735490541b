"""Helpers of compressai/models/utils.py that sit on the hot path: ``conv`` / ``deconv`` factories
(:128-146) and ``update_registered_buffers`` (:90-125, dynamic CDF buffer sizes at load time)."""
import torch
import torch.nn as nn


def find_named_buffer(module, query):
    """The buffer registered under the (possibly dotted) name ``query`` below ``module``, or None."""
    for name, buf in module.named_buffers():
        if name == query:
            return buf
    return None


_POLICIES = ("resize_if_empty", "resize", "register")


def _update_registered_buffer(module, buffer_name, state_dict_key, state_dict, policy="resize_if_empty",
                              dtype=torch.int):
    """Make ONE buffer of ``module`` shape-compatible with ``state_dict[state_dict_key]`` before loading."""
    if policy not in _POLICIES:
        raise ValueError(f'Invalid policy "{policy}"')
    want = state_dict[state_dict_key].size()
    have = find_named_buffer(module, buffer_name)
    if policy == "register":
        if have is not None:
            raise RuntimeError(f'buffer "{buffer_name}" was already registered')
        module.register_buffer(buffer_name, torch.zeros(want, dtype=dtype))
        return
    if have is None:
        raise RuntimeError(f'buffer "{buffer_name}" was not registered')
    if policy == "resize" or have.numel() == 0:
        have.resize_(want)


def update_registered_buffers(module, module_name, buffer_names, state_dict, policy="resize_if_empty",
                              dtype=torch.int):
    """The CDF buffers (``_quantized_cdf`` / ``_offset`` / ``_cdf_length`` / ``scale_table``) are created empty and
    sized by ``update()``, so a checkpoint taken after ``update()`` would not fit a fresh module: give every listed
    buffer of ``module`` the size found under ``{module_name}.{buffer}`` in ``state_dict`` first (contract of the
    reference helper, compressai/models/utils.py:90-125; policies "resize_if_empty" | "resize" | "register")."""
    known = {name for name, _ in module.named_buffers()}
    unknown = [b for b in buffer_names if b not in known]
    if unknown:
        raise ValueError(f'Invalid buffer name "{unknown[0]}"')
    for b in buffer_names:
        _update_registered_buffer(module, b, f"{module_name}.{b}", state_dict, policy, dtype)


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    from ..transforms import Conv2d

    return Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    from ..transforms import ConvTranspose2d

    return ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                           output_padding=stride - 1, padding=kernel_size // 2)
