"""Host helpers on the decoder path (/root/reference/utils.py:10-14, 29-35)."""
import torch


def get_mask_from_lengths(lengths: torch.Tensor) -> torch.Tensor:
    """bool [B, max(lengths)], True on valid positions.  Same result as the reference
    (utils.py:10-14) but on ``lengths.device`` instead of a hard-coded CUDA tensor."""
    max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device, dtype=torch.long)
    return ids < lengths.unsqueeze(1)


def to_gpu(x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    if torch.cuda.is_available():
        x = x.cuda(non_blocking=True)
    return x
