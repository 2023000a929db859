"""Parity metrics (oracle protocol item 4).  Test infrastructure."""
import torch


def rel_l2(candidate: torch.Tensor, truth: torch.Tensor) -> float:
    """||candidate - truth||_2 / ||truth||_2 computed in float64."""
    c = candidate.detach().to("cpu", torch.float64).reshape(-1)
    t = truth.detach().to("cpu", torch.float64).reshape(-1)
    if c.shape != t.shape:
        raise ValueError(f"shape mismatch: {tuple(candidate.shape)} vs {tuple(truth.shape)}")
    denom = t.norm().item()
    if denom == 0.0:
        return float((c - t).norm().item())
    return float((c - t).norm().item() / denom)


def centred_rel_l2(candidate: torch.Tensor, truth: torch.Tensor) -> float:
    """rel-L2 after removing the per-channel mean over every other axis.

    Random-init networks produce outputs dominated by input-independent
    per-channel offsets; removing the channel mean of
    the *truth* from both sides makes the metric sensitive to the part of the
    output that actually depends on the input.  The channel axis is the last.
    """
    c = candidate.detach().to("cpu", torch.float64)
    t = truth.detach().to("cpu", torch.float64)
    if c.shape != t.shape:
        raise ValueError(f"shape mismatch: {tuple(candidate.shape)} vs {tuple(truth.shape)}")
    mean = t.reshape(-1, t.shape[-1]).mean(dim=0)
    return rel_l2(c - mean, t - mean)
