"""Frame-range sharding over GPUs: frames are independent units under the
parity contract, so rank r of N simply owns a contiguous range - no collective on the data path."""


def frame_range(total_frames: int, rank: int, world: int):
    """Contiguous [begin, end) of rank's frames; sizes differ by at most one, earlier ranks get the extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(max(total_frames, 0), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)
