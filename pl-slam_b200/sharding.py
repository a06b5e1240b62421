"""Frame sharding across ranks: extraction/matching/pose-LM are independent per frame, so a sequence is
split into contiguous blocks with a 1-frame halo (frame k is matched against k-1, which the owning rank recomputes
instead of receiving ~60 KB of features); the only exchange is an all-gather of fixed-size pose records so that the rank
running the sequential Tracking logic sees them in frame order."""
import numpy as np

RECORD_FLOATS = 20   # 16 pose + frame id + inliers + n_keypoints + n_keylines


def shard_frames(n_frames, world, rank):
    """Contiguous block of this rank: (first, count, halo) — halo = 1 if the block needs the previous frame."""
    base, rem = divmod(n_frames, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count, (1 if first > 0 and count > 0 else 0)


def make_records(first, poses, inliers, n_kp, n_kl):
    """[count][RECORD_FLOATS] float32 records for frames first .. first+count-1."""
    count = len(poses)
    rec = np.zeros((count, RECORD_FLOATS), np.float32)
    rec[:, :16] = np.asarray(poses, np.float32).reshape(count, 16)
    rec[:, 16] = first + np.arange(count)
    rec[:, 17] = inliers; rec[:, 18] = n_kp; rec[:, 19] = n_kl
    return rec


def gather_records(local, n_frames, world, dist=None, device="cpu"):
    """All-gather the per-frame records of every rank and return them ordered by frame id ([n_frames][RECORD_FLOATS]).
    Blocks may differ by one frame, so each rank pads to the maximum block size; padded rows carry frame id -1."""
    import torch
    max_count = -(-n_frames // world)
    pad = torch.full((max_count, RECORD_FLOATS), -1.0, dtype=torch.float32, device=device)
    if len(local):
        pad[:len(local)] = torch.as_tensor(local, device=device)
    if dist is None or world == 1:
        allr = pad
    else:
        allr = torch.empty((world * max_count, RECORD_FLOATS), dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(allr, pad)
    allr = allr.cpu().numpy()
    allr = allr[allr[:, 16] >= 0]
    order = np.argsort(allr[:, 16], kind="stable")
    return allr[order]
