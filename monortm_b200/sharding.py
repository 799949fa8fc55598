"""Multi-GPU partitioning of the hot path: one process per GPU, no data-path collective.

Frequencies (configs 3, 5) and profiles (config 4) are independent through MODM, CALCTMR and RTM
(src/modm.f90:253, src/RTMmono.f90:177,286, src/monortm.f90:357), so each rank computes a contiguous
block and the only communication is the final gather of the spectra.  The spectral quantities the
reference derives from the WHOLE run -- v1=wn(1), v2=wn(nwn) for the continuum grid and the gridded
interpolation origin (modm.f90:180-185,218) and the line-load range (lnfl_mod.f90:116,161) -- stay
global: every rank passes (v1, v2, iw0) through mrtm_opts.
"""
import numpy as np


def block_partition(n, rank, world):
    """Contiguous block [start, start+count) of n items for `rank` of `world` (remainder spread over
    the first ranks)."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return start, count


def freq_shard(wn, rank, world):
    """Returns (wn_local, (v1_global, v2_global, iw0)) for a frequency-sharded run."""
    wn = np.asarray(wn, dtype=np.float64)
    start, count = block_partition(len(wn), rank, world)
    return wn[start:start + count], (float(wn[0]), float(wn[-1]), start)


def profile_shard(prof, rank, world):
    """Slice a profile dict (trailing profile dimension) for a profile-sharded ensemble run."""
    start, count = block_partition(int(prof["nprof"]), rank, world)
    out = dict(prof)
    for k in ("p", "t", "tz", "clw", "wbrodl"):
        out[k] = np.asfortranarray(prof[k][:, start:start + count])
    out["wkl"] = np.asfortranarray(prof["wkl"][:, :, start:start + count])
    out["nprof"] = count
    return out, start


def gather_blocks(local, counts, dist=None, group=None):
    """All-gather variable-sized leading-dimension blocks with torch.distributed (NCCL on GPUs,
    gloo on CPU).  `local` is a torch tensor (n_local, ...); returns the concatenation over ranks."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    nmax = max(counts)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    buf = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(buf, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(buf, counts)], dim=0)
