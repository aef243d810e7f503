"""TEST INFRASTRUCTURE ONLY: sort-based torch formulation of the batch replay screen -- the semantic reference the engine's
hash-table kernels (act_flag_replays[_dev]) are compared against, and what the CPU-only gloo test uses in their place.
NullifierDb semantics of /root/reference src/tests.rs:28-50 applied in slice order."""
import torch


def flag_replays(status, nullifiers, seen=None):
    n = status.numel()
    if n == 0:
        return status.clone()
    keys = nullifiers.view(n, 32).contiguous().view(torch.int64).view(n, 4)
    ok = status == 0
    if seen is not None and seen.numel():
        k = seen.numel() // 32
        allk = torch.cat([seen.view(k, 32).contiguous().view(torch.int64).view(k, 4), keys])
        first_pos = torch.cat([torch.full((k,), -1, dtype=torch.int64, device=status.device),
                               torch.where(ok, torch.arange(n, device=status.device), torch.full((n,), n, device=status.device))])
    else:
        allk = keys
        first_pos = torch.where(ok, torch.arange(n, device=status.device), torch.full((n,), n, device=status.device))
    _, inv = torch.unique(allk, dim=0, return_inverse=True)
    groups = int(inv.max().item()) + 1
    first = torch.full((groups,), n, dtype=torch.int64, device=status.device).scatter_reduce(0, inv, first_pos, reduce="amin")
    mine = inv[-n:]
    dup = ok & (first[mine] != torch.arange(n, device=status.device))
    out = status.clone()
    out[dup] = 3
    return out
