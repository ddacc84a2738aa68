"""bird_view/utils/train_utils.py:33-40 -- ``one_hot`` (host side, as in the reference)."""
import torch


def one_hot(x, num_digits=4, start=1):
    """Float class ids ``start..start+num_digits-1`` -> one-hot [N,num_digits] fp32 on the CPU;
    out-of-range ids clamp to the first / last class."""
    n = x.size()[0]
    idx = torch.clamp(x.long()[:, None] - start, 0, num_digits - 1)
    y = torch.zeros(n, num_digits, dtype=torch.float32)
    y.scatter_(1, idx.cpu(), 1)
    return y
