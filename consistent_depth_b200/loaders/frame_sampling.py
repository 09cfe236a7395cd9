"""Frame-pair sampling — the 'hierarchical2' scheme of the reference's utils/frame_sampling.py:78-122
(pairs at distance 2^l with start stride 2^max(0,l-1)), and `to_one_way` (:140-147)."""
import math


def sample_hierarchical2(num_frames, two_way=True, min_dist=1, max_dist=None):
    if max_dist is None:
        max_dist = num_frames - 1
    min_level = int(math.ceil(math.log2(min_dist)))
    max_level = int(math.floor(math.log2(max_dist)))
    signs = (-1, 1) if two_way else (1,)
    pairs = set()
    for level in range(min_level, max_level + 1):
        dist = 1 << level
        step = 1 << max(0, level - 1)
        for start in range(0, num_frames, step):
            for sign in signs:
                end = start + sign * dist
                if 0 <= end < num_frames:
                    pairs.add((start, end))
    return pairs


def to_one_way(pairs):
    return {(a, b) if a <= b else (b, a) for a, b in pairs}


def hierarchical2_one_way(num_frames):
    """Sorted one-way pair list (138 pairs for 50 frames, 286 for 100, 584 for 200)."""
    return sorted(to_one_way(sample_hierarchical2(num_frames, True)))


def to_in_range(pairs, frame_range=None):
    """Keep the pairs whose two frames lie in [frame_range[0], frame_range[1]) (utils/frame_sampling.py:149-156);
    no range given: all pairs (how loaders/video_dataset.py:120 calls it)."""
    if frame_range is None:
        return pairs
    lo, hi = frame_range[0], frame_range[1]
    return [p for p in pairs if all(lo <= i < hi for i in p)]
